// terrain.cu -- heightmap -> grid mesh kernels for sm_100a.
//
// Replaces the body of Terrain.create_terrain (Terrain/Terrain.zig:88-129) plus the per-frame WGSL
// vertex formula (Terrain/Terrain.zig:21-50) with a one-off mesh build:
//   terrain_vertices_k   streaming stencil: height tile (+1 halo) staged in shared memory as f32
//                        (u16 -> f32 of Terrain.zig:120 fused into the load), one thread per
//                        column walking down the tile, one 32-byte STG.256 per vertex
//                        (position slot + normal slot), 1 KB contiguous per warp store.
//   terrain_indices_k    closed-form, write-only: one 16-byte store per thread, aligned per row.
//   heightmap_normalize_k  standalone Terrain.zig:120.
// Float rules: every product/difference/quotient is its own IEEE round-to-nearest operation
// (__fmul_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn are never contracted), so positions are bit-identical
// to the reference formula evaluated without FMA and normals follow the spec in the header.
#include "common.cuh"

namespace {

constexpr int TV_THREADS = 256;  // columns per tile
constexpr int TV_ROWS = 8;       // rows per tile

struct TerrainArgs {
    const void* height;
    uint32_t n;
    uint32_t height_row0, height_rows;  // rows of the heightmap present at `height`
    uint32_t row_begin, row_end;
    unsigned char* vtx_out;  // already offset so that row `row_begin` of the band is addressable
    uint32_t vtx_row0;
    float grid_step, origin_scale, height_scale;
    uint32_t stride, pos_off, nrm_off;  // nrm_off == 0xFFFFFFFF -> no normal attribute
};

__device__ __forceinline__ float load_height_u16(const uint16_t* p) {
    // Terrain.zig:120: 1.0 - f32(u16) / 65535.0
    return __fsub_rn(1.0f, __fdiv_rn((float)__ldg(p), 65535.0f));
}

template <bool U16>
__device__ __forceinline__ float load_height(const void* base, size_t idx) {
    if (U16) return load_height_u16(reinterpret_cast<const uint16_t*>(base) + idx);
    return __ldg(reinterpret_cast<const float*>(base) + idx);
}

__device__ __forceinline__ void store_vertex32(unsigned char* p, float a0, float a1, float a2, float b0,
                                               float b1, float b2) {
    // one 256-bit store: [a0 a1 a2 0 | b0 b1 b2 0]
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a0), "f"(a1), "f"(a2),
                 "f"(0.0f), "f"(b0), "f"(b1), "f"(b2), "f"(0.0f)
                 : "memory");
}

// FAST32: stride 32, {pos,normal} at {0,16} or {16,0}: one STG.256 per vertex.
template <bool U16, bool FAST32>
__global__ void __launch_bounds__(TV_THREADS) terrain_vertices_k(const TerrainArgs a) {
    __shared__ float tile[TV_ROWS + 2][TV_THREADS + 2];
    const uint32_t n = a.n;
    const uint32_t c0 = blockIdx.x * TV_THREADS;
    const uint32_t r0 = a.row_begin + blockIdx.y * TV_ROWS;
    const uint32_t tid = threadIdx.x;

    // ---- stage the (TV_ROWS+2) x (TV_THREADS+2) height tile, clamped at the borders ----
    // Rows are clamped to the terrain (central differences are one-sided at the border) and to the
    // rows the caller provided (a band-local buffer holds only the band plus its halo; rows beyond
    // it belong to tile rows past row_end, which are never emitted).
    const int64_t r_lo = a.height_row0;
    const int64_t r_hi = min((int64_t)n, (int64_t)a.height_row0 + a.height_rows) - 1;
    {
        const uint32_t c = min(c0 + tid, n - 1);
#pragma unroll
        for (int rr = 0; rr < TV_ROWS + 2; ++rr) {
            int64_t r = (int64_t)r0 + rr - 1;
            r = r < r_lo ? r_lo : (r > r_hi ? r_hi : r);
            tile[rr][tid + 1] = load_height<U16>(a.height, (size_t)((uint32_t)r - a.height_row0) * n + c);
        }
        if (tid < 2 * (TV_ROWS + 2)) {
            const int rr = tid >> 1;
            const int side = tid & 1;
            int64_t r = (int64_t)r0 + rr - 1;
            r = r < r_lo ? r_lo : (r > r_hi ? r_hi : r);
            int64_t cc = side ? (int64_t)c0 + TV_THREADS : (int64_t)c0 - 1;
            cc = cc < 0 ? 0 : (cc > (int64_t)n - 1 ? (int64_t)n - 1 : cc);
            tile[rr][side ? TV_THREADS + 1 : 0] =
                load_height<U16>(a.height, (size_t)((uint32_t)r - a.height_row0) * n + (uint32_t)cc);
        }
    }
    __syncthreads();

    const uint32_t c = c0 + tid;
    if (c >= n) return;
    const float org = __fmul_rn(a.origin_scale, (float)n);
    const float z = __fsub_rn(__fmul_rn(a.grid_step, (float)c), org);
    const uint32_t cm = c > 0 ? c - 1 : 0, cp = c + 1 < n ? c + 1 : n - 1;
    const float den_c = __fmul_rn(a.grid_step, (float)(cp - cm));
    const uint32_t rows = min((uint32_t)TV_ROWS, a.row_end - r0);
    unsigned char* out = a.vtx_out + ((size_t)(r0 - a.vtx_row0) * n + c) * a.stride;
    const size_t row_pitch = (size_t)n * a.stride;

    float up = tile[0][tid + 1];
    float mid = tile[1][tid + 1];
#pragma unroll
    for (int rr = 0; rr < TV_ROWS; ++rr) {
        const float down = tile[rr + 2][tid + 1];
        if ((uint32_t)rr < rows) {
            const uint32_t r = r0 + rr;
            const float left = tile[rr + 1][tid];
            const float right = tile[rr + 1][tid + 2];
            const float x = __fsub_rn(__fmul_rn(a.grid_step, (float)r), org);
            const float y = __fmul_rn(a.height_scale, mid);
            float nx = 0.0f, ny = 1.0f, nz = 0.0f;
            if (a.nrm_off != 0xFFFFFFFFu) {
                const uint32_t rm = r > 0 ? r - 1 : 0, rp = r + 1 < n ? r + 1 : n - 1;
                float gx = 0.0f, gz = 0.0f;
                if (rp != rm)
                    gx = __fdiv_rn(__fmul_rn(a.height_scale, __fsub_rn(down, up)),
                                   __fmul_rn(a.grid_step, (float)(rp - rm)));
                if (cp != cm) gz = __fdiv_rn(__fmul_rn(a.height_scale, __fsub_rn(right, left)), den_c);
                const float len =
                    __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(gx, gx), 1.0f), __fmul_rn(gz, gz)));
                nx = __fdiv_rn(-gx, len);
                ny = __fdiv_rn(1.0f, len);
                nz = __fdiv_rn(-gz, len);
            }
            unsigned char* v = out + (size_t)rr * row_pitch;
            if (FAST32) {
                if (a.pos_off == 0)
                    store_vertex32(v, x, y, z, nx, ny, nz);
                else
                    store_vertex32(v, nx, ny, nz, x, y, z);
            } else {
                // generic layout: zero the vertex, then the attributes (4-byte granularity)
                float* f = reinterpret_cast<float*>(v);
                for (uint32_t k = 0; k < a.stride / 4; ++k) f[k] = 0.0f;
                float* p = reinterpret_cast<float*>(v + a.pos_off);
                p[0] = x;
                p[1] = y;
                p[2] = z;
                if (a.nrm_off != 0xFFFFFFFFu) {
                    float* q = reinterpret_cast<float*>(v + a.nrm_off);
                    q[0] = nx;
                    q[1] = ny;
                    q[2] = nz;
                }
            }
        }
        up = mid;
        mid = down;
    }
}

// ---- index buffer ------------------------------------------------------------------------
// Row q of quads occupies L = 6*(n-1) consecutive u32.  Each thread owns one 16-byte aligned
// group of 4 words of the output; the group is mapped back to (quad, corner) with a division
// by the constant 6 only.
constexpr int TI_THREADS = 256;
constexpr int TI_GROUPS = 4;  // 16-byte groups per thread

struct IndexArgs {
    uint32_t* idx_out;
    uint32_t n;
    uint32_t qrow_begin, qrow_end, idx_qrow0;
};

__device__ __forceinline__ uint32_t quad_corner_index(uint32_t i00, uint32_t n, uint32_t k) {
    // corner order of Terrain.zig:28-35 / lookups :38-45
    //   k: 0 (r+1,c)  1 (r,c)  2 (r+1,c+1)  3 (r+1,c+1)  4 (r,c)  5 (r,c+1)
    const uint32_t add_n = (0x0Du >> k) & 1u;  // k in {0,2,3}
    const uint32_t add_1 = (0x2Cu >> k) & 1u;  // k in {2,3,5}
    return i00 + (add_n ? n : 0u) + add_1;
}

__global__ void __launch_bounds__(TI_THREADS) terrain_indices_k(const IndexArgs a) {
    const uint32_t q = a.qrow_begin + blockIdx.y;
    const uint32_t L = 6u * (a.n - 1u);
    uint32_t* row = a.idx_out + (size_t)(q - a.idx_qrow0) * L;
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(row) >> 2) & 3u);  // words past 16B alignment
    const uint32_t row_i0 = q * a.n;
#pragma unroll
    for (int g = 0; g < TI_GROUPS; ++g) {
        const uint32_t grp = (blockIdx.x * TI_GROUPS + g) * TI_THREADS + threadIdx.x;
        const int64_t w0 = (int64_t)grp * 4 - mis;  // first word of the group, relative to the row
        if (w0 >= (int64_t)L) break;
        uint32_t v[4];
        const uint32_t wfirst = w0 < 0 ? 0u : (uint32_t)w0;
        uint32_t quad = wfirst / 6u;
        uint32_t k = wfirst - quad * 6u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t w = w0 + i;
            if (w >= (int64_t)wfirst) {
                v[i] = quad_corner_index(row_i0 + quad, a.n, k);
                if (++k == 6u) {
                    k = 0;
                    ++quad;
                }
            } else {
                v[i] = 0;
            }
        }
        if (w0 >= 0 && w0 + 4 <= (int64_t)L) {
            *reinterpret_cast<uint4*>(row + w0) = make_uint4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t w = w0 + i;
                if (w >= 0 && w < (int64_t)L) row[w] = v[i];
            }
        }
    }
}

__global__ void __launch_bounds__(256) heightmap_normalize_k(const uint16_t* __restrict__ in,
                                                             float* __restrict__ out, uint64_t count) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (; i < count; i += step) out[i] = load_height_u16(in + i);
}

}  // namespace

int mr_terrain_build_impl(mr_context* ctx, const mr_terrain_job* j) {
    const uint32_t n = j->n;
    if (j->vtx_out && j->row_end > j->row_begin) {
        TerrainArgs a;
        a.height = j->height;
        a.n = n;
        a.height_row0 = j->height_row0;
        a.height_rows = j->height_rows;
        a.row_begin = j->row_begin;
        a.row_end = j->row_end;
        a.vtx_out = static_cast<unsigned char*>(j->vtx_out);
        a.vtx_row0 = j->vtx_row0;
        a.grid_step = j->params.grid_step;
        a.origin_scale = j->params.origin_scale;
        a.height_scale = j->params.height_scale;
        a.stride = j->layout.stride;
        a.pos_off = j->layout.attr[0].offset;
        a.nrm_off = j->layout.nattr > 1 ? j->layout.attr[1].offset : 0xFFFFFFFFu;
        const bool fast32 = a.stride == 32 && j->layout.nattr == 2 &&
                            ((a.pos_off == 0 && a.nrm_off == 16) || (a.pos_off == 16 && a.nrm_off == 0)) &&
                            (reinterpret_cast<uintptr_t>(a.vtx_out) % 32 == 0);
        const uint32_t rows = j->row_end - j->row_begin;
        dim3 grid((n + TV_THREADS - 1) / TV_THREADS, (rows + TV_ROWS - 1) / TV_ROWS);
        // grid.y limit is 65535: 8 rows per tile covers n up to 524k rows per launch
        if (grid.y > 65535u) return mr_fail(ctx, MR_E_BADARG, "terrain band too tall for one launch");
        const bool u16 = j->height_fmt == MR_HEIGHT_U16;
        if (u16 && fast32)
            terrain_vertices_k<true, true><<<grid, TV_THREADS, 0, ctx->stream>>>(a);
        else if (u16)
            terrain_vertices_k<true, false><<<grid, TV_THREADS, 0, ctx->stream>>>(a);
        else if (fast32)
            terrain_vertices_k<false, true><<<grid, TV_THREADS, 0, ctx->stream>>>(a);
        else
            terrain_vertices_k<false, false><<<grid, TV_THREADS, 0, ctx->stream>>>(a);
        MR_LAUNCH_CHECK(ctx, "terrain_vertices_k");
    }
    if (j->idx_out && n > 1 && j->qrow_end > j->qrow_begin) {
        IndexArgs a;
        a.idx_out = j->idx_out;
        a.n = n;
        a.qrow_begin = j->qrow_begin;
        a.qrow_end = j->qrow_end;
        a.idx_qrow0 = j->idx_qrow0;
        const uint32_t L = 6u * (n - 1u);
        const uint32_t groups = (L + 3u) / 4u + 1u;  // +1: a misaligned row spills into one more group
        const uint32_t per_block = TI_THREADS * TI_GROUPS;
        uint32_t rows_left = j->qrow_end - j->qrow_begin;
        // grid.y is limited to 65535 rows per launch
        while (rows_left) {
            const uint32_t rows = rows_left > 65535u ? 65535u : rows_left;
            dim3 grid((groups + per_block - 1) / per_block, rows);
            terrain_indices_k<<<grid, TI_THREADS, 0, ctx->stream>>>(a);
            MR_LAUNCH_CHECK(ctx, "terrain_indices_k");
            a.qrow_begin += rows;
            rows_left -= rows;
        }
    }
    return MR_OK;
}

int mr_heightmap_normalize_impl(mr_context* ctx, const uint16_t* in, uint64_t count, float* out) {
    if (count == 0) return MR_OK;
    uint64_t blocks = (count + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    heightmap_normalize_k<<<(unsigned)blocks, 256, 0, ctx->stream>>>(in, out, count);
    MR_LAUNCH_CHECK(ctx, "heightmap_normalize_k");
    return MR_OK;
}
