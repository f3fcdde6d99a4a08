/*
 * o_synth.c -- host restatement of the synthetic-workload generators.
 * TEST INFRASTRUCTURE ONLY (see mr_oracle.h).  Not reference code: the reference has no
 * generators (unirand.zig is an index permuter).  The definitions are repeated
 * here, independently of myrenderer_b200/csrc/synth.cu, so tests can check that the
 * device generators produce these exact bytes.
 */
#include <math.h>
#include "mr_oracle.h"

static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void mr_o_synth_heightmap_u16(uint64_t seed, uint32_t n, uint32_t row0, uint32_t rows, uint16_t* out) {
    uint64_t i, first = (uint64_t)row0 * n, count = (uint64_t)rows * n;
    for (i = 0; i < count; ++i) out[i] = (uint16_t)(mix64(seed ^ (first + i)) >> 48);
}

void mr_o_synth_polygon_sizes(uint64_t seed, uint64_t poly_index0, uint32_t npoly, uint32_t nmin,
                              uint32_t nmax, int dist, uint64_t* first_point_out) {
    uint64_t acc = 0;
    uint32_t i;
    first_point_out[0] = 0;
    for (i = 0; i < npoly; ++i) {
        uint64_t h = mix64(seed ^ mix64(poly_index0 + i));
        uint32_t n;
        if (dist == MR_SIZES_LOGUNIFORM) {
            double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
            n = (uint32_t)floor((double)nmin * pow((double)nmax / (double)nmin, u));
            if (n > nmax) n = nmax;
            if (n < nmin) n = nmin;
        } else {
            n = nmin + (uint32_t)(h % (uint64_t)(nmax - nmin + 1u));
        }
        acc += n;
        first_point_out[i + 1] = acc;
    }
}

/* sin/cos of 2*pi*t for t in [0,1): fixed polynomial, every operation a single IEEE
 * double operation in a fixed order, so the device version (explicit __dmul_rn/__dadd_rn)
 * returns the same bits. */
static void sincos_turn(double t, double* s_out, double* c_out) {
    const double two_pi = 6.283185307179586476925286766559;
    double j = floor(t * 4.0 + 0.5); /* nearest quarter turn */
    double f = t - j * 0.25;         /* [-1/8, 1/8] */
    double x = f * two_pi;
    double x2 = x * x;
    /* Taylor to x^17 / x^16: |x| <= pi/4 -> truncation < 1e-17 */
    double s = -1.0 / 355687428096000.0;
    double c = 1.0 / 20922789888000.0;
    int q;
    s = s * x2 + 1.0 / 1307674368000.0;
    s = s * x2 + -1.0 / 6227020800.0;
    s = s * x2 + 1.0 / 39916800.0;
    s = s * x2 + -1.0 / 362880.0;
    s = s * x2 + 1.0 / 5040.0;
    s = s * x2 + -1.0 / 120.0;
    s = s * x2 + 1.0 / 6.0;
    s = s * x2;
    s = x - x * s;
    c = c * x2 + -1.0 / 87178291200.0;
    c = c * x2 + 1.0 / 479001600.0;
    c = c * x2 + -1.0 / 3628800.0;
    c = c * x2 + 1.0 / 40320.0;
    c = c * x2 + -1.0 / 720.0;
    c = c * x2 + 1.0 / 24.0;
    c = c * x2 + -0.5;
    c = c * x2 + 1.0;
    q = (int)j & 3;
    switch (q) {
        case 0: *s_out = s; *c_out = c; break;
        case 1: *s_out = c; *c_out = -s; break;
        case 2: *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out = s; break;
    }
}

/* One vertex of a synthetic polygon (definitions in include/myrenderer_b200.h, MR_FAMILY_*). */
static void synth_vertex(int family, uint64_t key, uint32_t n, uint32_t k, float* x_out, float* y_out) {
    uint64_t h = mix64(key + k);
    double u1 = (double)(h >> 40) * (1.0 / 16777216.0);
    double u2 = (double)((h >> 16) & 0xFFFFFFu) * (1.0 / 16777216.0);
    if (family == MR_FAMILY_ZIPPER) {
        /* two y-monotone chains on strictly interleaved levels: vertices 0..ceil(n/2)-1 walk down the
         * right chain on the even levels, the rest walk up the left chain on the odd levels */
        uint32_t na = (n + 1u) / 2u;
        uint32_t level = k < na ? 2u * k : 2u * (n - 1u - k) + 1u;
        double y = 10.0 + 180.0 * (((double)level + (0.1 + 0.8 * u1)) / (double)n);
        double x = k < na ? 105.0 + 85.0 * u2 : 10.0 + 85.0 * u2;
        *x_out = (float)x;
        *y_out = (float)y;
    } else {
        double t = ((double)k + (0.8 * u1 - 0.4)) / (double)n;
        double s, c;
        if (t < 0.0) t = t + 1.0;
        sincos_turn(t, &s, &c);
        if (family == MR_FAMILY_ELLIPSE) {
            /* convex: points of an ellipse (semi-axes 40..90) rotated by a per-polygon phase */
            uint64_t g = mix64(key ^ 0x5bd1e9955bd1e995ull);
            double a = 40.0 + 50.0 * ((double)(g >> 40) * (1.0 / 16777216.0));
            double b = 40.0 + 50.0 * ((double)((g >> 16) & 0xFFFFFFu) * (1.0 / 16777216.0));
            double ph = (double)(mix64(g) >> 40) * (1.0 / 16777216.0);
            double sp, cp, ex, ey;
            sincos_turn(ph, &sp, &cp);
            ex = a * c;
            ey = b * s;
            *x_out = (float)(100.0 + (cp * ex - sp * ey));
            *y_out = (float)(100.0 + (sp * ex + cp * ey));
        } else { /* MR_FAMILY_STAR */
            double radius = 20.0 + 70.0 * u2;
            *x_out = (float)(100.0 + radius * c);
            *y_out = (float)(100.0 + radius * s);
        }
    }
}

void mr_o_synth_polygons_family(int family, uint64_t seed, uint64_t poly_index0, const uint64_t* first_point,
                                uint32_t npoly, float* xy_out) {
    uint32_t i;
    for (i = 0; i < npoly; ++i) {
        uint64_t p0 = first_point[i] - first_point[0];
        uint32_t n = (uint32_t)(first_point[i + 1] - first_point[i]);
        uint64_t key = mix64(seed ^ mix64((poly_index0 + i) ^ 0xA5A5A5A5A5A5A5A5ull));
        uint32_t k;
        for (k = 0; k < n; ++k) synth_vertex(family, key, n, k, &xy_out[2u * (p0 + k)], &xy_out[2u * (p0 + k) + 1u]);
    }
}

void mr_o_synth_polygons(uint64_t seed, uint64_t poly_index0, const uint64_t* first_point,
                         uint32_t npoly, float* xy_out) {
    mr_o_synth_polygons_family(MR_FAMILY_STAR, seed, poly_index0, first_point, npoly, xy_out);
}
