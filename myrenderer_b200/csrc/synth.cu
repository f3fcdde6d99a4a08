// synth.cu -- device generators for the benchmark workloads: hash-noise heightmaps
// and star-shaped simple polygons.  Not reference code (the reference has no generators;
// unirand.zig is an index permuter).  Integer/f64 arithmetic only, every f64 operation separately
// rounded, so the bytes equal the host definition used by the tests.
#include <cmath>
#include "common.cuh"

namespace {

__global__ void synth_heightmap_u16_k(uint64_t seed, uint64_t first, uint64_t count, uint16_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (; i < count; i += step) out[i] = (uint16_t)(mr_mix64(seed ^ (first + i)) >> 48);
}

// sin/cos of 2*pi*t, t in [0,1): fixed polynomial, single IEEE double operations in a fixed order
__device__ __forceinline__ void sincos_turn(double t, double* s_out, double* c_out) {
    const double two_pi = 6.283185307179586476925286766559;
    const double j = floor(__dadd_rn(__dmul_rn(t, 4.0), 0.5));
    const double f = __dsub_rn(t, __dmul_rn(j, 0.25));
    const double x = __dmul_rn(f, two_pi);
    const double x2 = __dmul_rn(x, x);
    double s = -1.0 / 355687428096000.0;
    double c = 1.0 / 20922789888000.0;
    s = __dadd_rn(__dmul_rn(s, x2), 1.0 / 1307674368000.0);
    s = __dadd_rn(__dmul_rn(s, x2), -1.0 / 6227020800.0);
    s = __dadd_rn(__dmul_rn(s, x2), 1.0 / 39916800.0);
    s = __dadd_rn(__dmul_rn(s, x2), -1.0 / 362880.0);
    s = __dadd_rn(__dmul_rn(s, x2), 1.0 / 5040.0);
    s = __dadd_rn(__dmul_rn(s, x2), -1.0 / 120.0);
    s = __dadd_rn(__dmul_rn(s, x2), 1.0 / 6.0);
    s = __dmul_rn(s, x2);
    s = __dsub_rn(x, __dmul_rn(x, s));
    c = __dadd_rn(__dmul_rn(c, x2), -1.0 / 87178291200.0);
    c = __dadd_rn(__dmul_rn(c, x2), 1.0 / 479001600.0);
    c = __dadd_rn(__dmul_rn(c, x2), -1.0 / 3628800.0);
    c = __dadd_rn(__dmul_rn(c, x2), 1.0 / 40320.0);
    c = __dadd_rn(__dmul_rn(c, x2), -1.0 / 720.0);
    c = __dadd_rn(__dmul_rn(c, x2), 1.0 / 24.0);
    c = __dadd_rn(__dmul_rn(c, x2), -0.5);
    c = __dadd_rn(__dmul_rn(c, x2), 1.0);
    switch ((int)j & 3) {
        case 0: *s_out = s; *c_out = c; break;
        case 1: *s_out = c; *c_out = -s; break;
        case 2: *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out = s; break;
    }
}

// one vertex of a synthetic polygon (families: include/myrenderer_b200.h); same operations, same order as
// oracle/o_synth.c, every f64 operation separately rounded
__device__ __forceinline__ float2 synth_vertex(int family, uint64_t key, uint32_t n, uint32_t k) {
    const double inv24 = 1.0 / 16777216.0;
    const uint64_t h = mr_mix64(key + k);
    const double u1 = __dmul_rn((double)(h >> 40), inv24);
    const double u2 = __dmul_rn((double)((h >> 16) & 0xFFFFFFull), inv24);
    if (family == MR_FAMILY_ZIPPER) {
        const uint32_t na = (n + 1u) / 2u;
        const uint32_t level = k < na ? 2u * k : 2u * (n - 1u - k) + 1u;
        const double y = __dadd_rn(10.0, __dmul_rn(180.0, __ddiv_rn(__dadd_rn((double)level, __dadd_rn(0.1, __dmul_rn(0.8, u1))), (double)n)));
        const double x = __dadd_rn(k < na ? 105.0 : 10.0, __dmul_rn(85.0, u2));
        return make_float2((float)x, (float)y);
    }
    double t = __ddiv_rn(__dadd_rn((double)k, __dsub_rn(__dmul_rn(0.8, u1), 0.4)), (double)n);
    if (t < 0.0) t = __dadd_rn(t, 1.0);
    double s, c;
    sincos_turn(t, &s, &c);
    if (family == MR_FAMILY_ELLIPSE) {
        const uint64_t g = mr_mix64(key ^ 0x5bd1e9955bd1e995ull);
        const double a = __dadd_rn(40.0, __dmul_rn(50.0, __dmul_rn((double)(g >> 40), inv24)));
        const double b = __dadd_rn(40.0, __dmul_rn(50.0, __dmul_rn((double)((g >> 16) & 0xFFFFFFull), inv24)));
        const double ph = __dmul_rn((double)(mr_mix64(g) >> 40), inv24);
        double sp, cp;
        sincos_turn(ph, &sp, &cp);
        const double ex = __dmul_rn(a, c), ey = __dmul_rn(b, s);
        return make_float2((float)__dadd_rn(100.0, __dsub_rn(__dmul_rn(cp, ex), __dmul_rn(sp, ey))),
                           (float)__dadd_rn(100.0, __dadd_rn(__dmul_rn(sp, ex), __dmul_rn(cp, ey))));
    }
    const double radius = __dadd_rn(20.0, __dmul_rn(70.0, u2));
    return make_float2((float)__dadd_rn(100.0, __dmul_rn(radius, c)), (float)__dadd_rn(100.0, __dmul_rn(radius, s)));
}

// one warp per polygon, lanes over vertices
__global__ void synth_polygons_k(int family, uint64_t seed, uint64_t poly_index0, const uint64_t* __restrict__ first_point,
                                 uint32_t npoly, float* __restrict__ xy_out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t base = first_point[0];
    for (uint32_t i = warp; i < npoly; i += nwarps) {
        const uint64_t p0 = first_point[i] - base;
        const uint32_t n = (uint32_t)(first_point[i + 1] - first_point[i]);
        const uint64_t key = mr_mix64(seed ^ mr_mix64((poly_index0 + i) ^ 0xA5A5A5A5A5A5A5A5ull));
        for (uint32_t k = lane; k < n; k += 32) {
            const float2 v = synth_vertex(family, key, n, k);
            xy_out[2 * (p0 + k)] = v.x;
            xy_out[2 * (p0 + k) + 1] = v.y;
        }
    }
}

}  // namespace

int mr_synth_heightmap_u16_impl(mr_context* ctx, uint64_t seed, uint32_t n, uint32_t row0, uint32_t rows,
                                uint16_t* out_dev) {
    const uint64_t count = (uint64_t)rows * n;
    if (count == 0) return MR_OK;
    uint64_t blocks = (count + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    synth_heightmap_u16_k<<<(unsigned)blocks, 256, 0, ctx->stream>>>(seed, (uint64_t)row0 * n, count, out_dev);
    MR_LAUNCH_CHECK(ctx, "synth_heightmap_u16_k");
    return MR_OK;
}

int mr_synth_polygons_impl(mr_context* ctx, int family, uint64_t seed, uint64_t poly_index0, const uint64_t* first_point_dev,
                           uint32_t npoly, float* xy_dev) {
    if (npoly == 0) return MR_OK;
    uint64_t blocks = ((uint64_t)npoly + 7) / 8;
    const uint64_t cap = (uint64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    synth_polygons_k<<<(unsigned)blocks, 256, 0, ctx->stream>>>(family, seed, poly_index0, first_point_dev, npoly, xy_dev);
    MR_LAUNCH_CHECK(ctx, "synth_polygons_k");
    return MR_OK;
}

// host helper (sizes are needed on the host to allocate buffers)
extern "C" int mr_synth_polygon_sizes(uint64_t seed, uint64_t poly_index0, uint32_t npoly, uint32_t nmin,
                                      uint32_t nmax, int dist, uint64_t* first_point_out) {
    if (!first_point_out || nmin < 1 || nmax < nmin) return MR_E_BADARG;
    uint64_t acc = 0;
    first_point_out[0] = 0;
    for (uint32_t i = 0; i < npoly; ++i) {
        const uint64_t h = mr_mix64(seed ^ mr_mix64(poly_index0 + i));
        uint32_t n;
        if (dist == MR_SIZES_LOGUNIFORM) {
            const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
            n = (uint32_t)std::floor((double)nmin * std::pow((double)nmax / (double)nmin, u));
            if (n > nmax) n = nmax;
            if (n < nmin) n = nmin;
        } else {
            n = nmin + (uint32_t)(h % (uint64_t)(nmax - nmin + 1u));
        }
        acc += n;
        first_point_out[i + 1] = acc;
    }
    return MR_OK;
}
