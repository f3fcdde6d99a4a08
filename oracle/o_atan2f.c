/*
 * o_atan2f.c -- oracle restatement of std.math.atan2 for f32.  TEST INFRASTRUCTURE ONLY.
 *
 * Call site: Polygon/Triangulation.zig:403 (twice per candidate triangle).
 * The implementation lives in the Zig standard library (pinned only by
 * minimum_zig_version 0.14.0-dev.2577, build.zig.zon:18), which is not under
 * /root/reference.  Zig's std/math/atan2.zig and atan.zig are ports of musl's
 * e_atan2f.c / s_atanf.c (FreeBSD msun); this file restates that published
 * algorithm.  No reference test pins it -> PARITY UNPINNED.
 *
 * Why that is tolerable: for finite input the comparison at :403 is always
 * true (both vectors point from the lowest list entry to entries sorted above
 * it, so both angles lie in [0,pi]) except when one angle rounds to exactly
 * (float)pi while the other is exactly 0 -- which needs |dy/dx| < ~3e-8 with
 * dx < 0 on one side and an underflowing or zero dy on the other.  Any atan2f
 * that is exact at 0 and rounds to (float)pi in that range gives the same
 * control flow.
 */
#include <string.h>
#include "mr_oracle.h"

static uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static float u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

float mr_o_atanf(float x) {
    /* hi/lo split of atan(0.5), atan(1), atan(1.5), atan(inf); odd polynomial coefficients */
    static const uint32_t hi[4] = {0x3eed6338u, 0x3f490fdau, 0x3f7b985eu, 0x3fc90fdau};
    static const uint32_t lo[4] = {0x31ac3769u, 0x33222168u, 0x33140fb4u, 0x33a22168u};
    static const float aT[5] = {3.3333328366e-01f, -1.9999158382e-01f, 1.4253635705e-01f,
                                -1.0648017377e-01f, 6.1687607318e-02f};
    uint32_t ix = f2u(x);
    uint32_t sign = ix >> 31;
    int id;
    float z, w, s1, s2;
    ix &= 0x7fffffffu;
    if (ix >= 0x4c800000u) { /* |x| >= 2^26 */
        if (ix > 0x7f800000u) return x; /* NaN */
        z = u2f(hi[3]) + 7.5231638453e-37f; /* 0x1p-120 */
        return sign ? -z : z;
    }
    if (ix < 0x3ee00000u) {     /* |x| < 0.4375 */
        if (ix < 0x39800000u) { /* |x| < 2^-12 */
            return x;
        }
        id = -1;
    } else {
        x = u2f(ix); /* fabsf */
        if (ix < 0x3f980000u) {     /* |x| < 1.1875 */
            if (ix < 0x3f300000u) { /* 7/16 <= |x| < 11/16 */
                id = 0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else { /* 11/16 <= |x| < 19/16 */
                id = 1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000u) { /* |x| < 2.4375 */
                id = 2;
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else { /* 2.4375 <= |x| < 2^26 */
                id = 3;
                x = -1.0f / x;
            }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (aT[0] + w * (aT[2] + w * aT[4]));
    s2 = w * (aT[1] + w * aT[3]);
    if (id < 0) return x - x * (s1 + s2);
    z = u2f(hi[id]) - ((x * (s1 + s2) - u2f(lo[id])) - x);
    return sign ? -z : z;
}

float mr_o_atan2f(float y, float x) {
    const float pi = u2f(0x40490fdbu);    /* 3.1415927410e+00 */
    const float pi_lo = u2f(0xb3bbbd2eu); /* -8.7422776573e-08 */
    uint32_t ix = f2u(x), iy = f2u(y), m;
    float z;
    if ((ix & 0x7fffffffu) > 0x7f800000u || (iy & 0x7fffffffu) > 0x7f800000u) return x + y;
    if (ix == 0x3f800000u) return mr_o_atanf(y); /* x == 1.0 */
    m = ((iy >> 31) & 1u) | ((ix >> 30) & 2u);   /* 2*sign(x) + sign(y) */
    ix &= 0x7fffffffu;
    iy &= 0x7fffffffu;
    if (iy == 0u) { /* y == 0 */
        switch (m) {
            case 0:
            case 1: return y;
            case 2: return pi;
            default: return -pi;
        }
    }
    if (ix == 0u) return (m & 1u) ? -pi / 2 : pi / 2; /* x == 0 */
    if (ix == 0x7f800000u) {                          /* x == inf */
        if (iy == 0x7f800000u) {
            switch (m) {
                case 0: return pi / 4;
                case 1: return -pi / 4;
                case 2: return 3 * pi / 4;
                default: return -3 * pi / 4;
            }
        } else {
            switch (m) {
                case 0: return 0.0f;
                case 1: return -0.0f;
                case 2: return pi;
                default: return -pi;
            }
        }
    }
    /* |y/x| > 2^26 */
    if (ix + (26u << 23) < iy || iy == 0x7f800000u) return (m & 1u) ? -pi / 2 : pi / 2;
    /* z = atan(|y/x|) with correct underflow */
    if ((m & 2u) && iy + (26u << 23) < ix)
        z = 0.0f; /* |y/x| < 2^-26, x < 0 */
    else
        z = mr_o_atanf(u2f(f2u(y / x) & 0x7fffffffu));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}
