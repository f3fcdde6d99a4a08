// api.cu -- the extern "C" surface of libmyrenderer_b200: context, memory helpers, argument
// validation, host<->device staging, and the host-side helpers (layout presets, partitions,
// unirand host restatement).  Kernels live in terrain.cu / triangulate.cu / synth.cu.
#include <algorithm>
#include <new>
#include "common.cuh"
#include "unirand.cuh"

int mr_heightmap_normalize_impl(mr_context* ctx, const uint16_t* in, uint64_t count, float* out);
int mr_terrain_tile_bounds_impl(mr_context* ctx, const void* height_dev, uint32_t height_fmt, uint32_t n, uint32_t tile_rows,
                                uint32_t tile_cols, const mr_terrain_params* p, float* bbox_dev);
int mr_terrain_cull_impl(mr_context* ctx, const float* bbox_dev, uint32_t n, uint32_t tile_rows, uint32_t tile_cols,
                         const float xform[16], uint32_t* visible_dev, uint32_t* ids_dev, unsigned long long* first_index_dev,
                         uint32_t* idx_dev, unsigned long long* counts_dev);
int mr_selftest_fastdiv_impl(mr_context* ctx, float b, int force_fast, unsigned long long* mismatches_dev);
int mr_polygon_offsets_impl(mr_context* ctx, const uint64_t* first_point_dev, uint32_t npoly, uint64_t* first_tri_dev);
int mr_triangulate_tier_counts_impl(mr_context* ctx, uint32_t out[8]);
int mr_unirand_seed_batch_impl(mr_context* ctx, const uint64_t* first_point_dev, uint32_t npoly, uint64_t seed,
                               uint64_t poly_index0, uint32_t* out_dev);
int mr_synth_heightmap_u16_impl(mr_context* ctx, uint64_t seed, uint32_t n, uint32_t row0, uint32_t rows,
                                uint16_t* out_dev);
int mr_synth_polygons_impl(mr_context* ctx, int family, uint64_t seed, uint64_t poly_index0, const uint64_t* first_point_dev,
                           uint32_t npoly, float* xy_dev);

namespace {
// unirand.zig:24: host view of the table defined once in unirand.cuh (mr_unirand_seed_host)
const uint32_t k_primes_host[MR_NPRIMES] = {MR_PRIME_LIST};

bool layout_ok(const mr_layout* L, uint32_t need_attrs, uint32_t ncomp0) {
    if (!L || L->nattr < need_attrs || L->nattr > MR_MAX_ATTR) return false;
    if (L->stride < 8 || L->stride > 256 || (L->stride & 3u)) return false;
    for (uint32_t i = 0; i < L->nattr; ++i) {
        if (L->attr[i].offset & 3u) return false;
        if (L->attr[i].ncomp < 2 || L->attr[i].ncomp > 4) return false;
        if (L->attr[i].offset + 4u * L->attr[i].ncomp > L->stride) return false;
    }
    if (L->attr[0].ncomp < ncomp0) return false;
    return true;
}
}  // namespace

extern "C" {

int mr_abi_version(void) { return MR_ABI_VERSION; }

int mr_build_flags(void) {
#ifdef MR_CHECKED
    return 1;
#else
    return 0;
#endif
}

int mr_device_count(int* count_out) {
    if (!count_out) return MR_E_BADARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count_out = 0;
        return MR_E_CUDA;
    }
    *count_out = n;
    return MR_OK;
}

int mr_context_create(int device, mr_context** ctx_out) {
    if (!ctx_out) return MR_E_BADARG;
    *ctx_out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return MR_E_CUDA;  // no CPU fallback
    }
    if (device < 0 || device >= n) return MR_E_BADARG;
    if (cudaSetDevice(device) != cudaSuccess) return MR_E_CUDA;
    mr_context* ctx = new (std::nothrow) mr_context();
    if (!ctx) return MR_E_NOMEM;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return MR_E_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return MR_E_CUDA;
    }
    ctx->own_stream = true;
    *ctx_out = ctx;
    return MR_OK;
}

int mr_context_destroy(mr_context* ctx) {
    if (!ctx) return MR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < MR_NUM_SCRATCH; ++i)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->pinned_mailbox) cudaFreeHost(ctx->pinned_mailbox);
    if (ctx->small_pinned) cudaFreeHost(ctx->small_pinned);
    if (ctx->small_dev) cudaFree(ctx->small_dev);
    mr_polygon_plan_free(ctx);
    for (int i = 0; i < MR_NUM_AUX; ++i) {
        if (ctx->aux[i]) cudaStreamDestroy(ctx->aux[i]);
        if (ctx->join_ev[i]) cudaEventDestroy(ctx->join_ev[i]);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MR_OK;
}

int mr_context_trim(mr_context* ctx) {
    if (!ctx) return MR_E_BADARG;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < MR_NUM_SCRATCH; ++i) {
        if (ctx->scratch[i]) MR_CUDA(ctx, cudaFree(ctx->scratch[i]));
        ctx->scratch[i] = nullptr;
        ctx->scratch_bytes[i] = 0;
    }
    ctx->last_header_dev = nullptr;
    return MR_OK;
}

int mr_context_scratch_bytes(const mr_context* ctx, uint64_t* bytes_out) {
    if (!ctx || !bytes_out) return MR_E_BADARG;
    uint64_t t = 0;
    for (int i = 0; i < MR_NUM_SCRATCH; ++i) t += ctx->scratch_bytes[i];
    *bytes_out = t;
    return MR_OK;
}

int mr_context_set_stream(mr_context* ctx, void* cuda_stream) {
    if (!ctx) return MR_E_BADARG;
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return MR_OK;
}

int mr_context_stream(mr_context* ctx, void** cuda_stream_out) {
    if (!ctx || !cuda_stream_out) return MR_E_BADARG;
    *cuda_stream_out = ctx->stream;
    return MR_OK;
}

int mr_sync(mr_context* ctx) {
    if (!ctx) return MR_E_BADARG;
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MR_OK;
}

const char* mr_last_error(const mr_context* ctx) { return ctx ? ctx->err : "null context"; }
uint64_t mr_launch_count(const mr_context* ctx) { return ctx ? ctx->launches : 0; }

int mr_device_alloc(mr_context* ctx, size_t bytes, void** dev_out) {
    if (!ctx || !dev_out) return MR_E_BADARG;
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc(dev_out, bytes ? bytes : 16);
    if (e != cudaSuccess) return mr_fail(ctx, MR_E_NOMEM, "cudaMalloc", e);
    return MR_OK;
}
int mr_device_free(mr_context* ctx, void* dev) {
    if (!ctx) return MR_E_BADARG;
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MR_CUDA(ctx, cudaFree(dev));
    return MR_OK;
}
int mr_pinned_alloc(mr_context* ctx, size_t bytes, void** host_out) {
    if (!ctx || !host_out) return MR_E_BADARG;
    cudaError_t e = cudaMallocHost(host_out, bytes ? bytes : 16);
    if (e != cudaSuccess) return mr_fail(ctx, MR_E_NOMEM, "cudaMallocHost", e);
    return MR_OK;
}
int mr_pinned_free(mr_context* ctx, void* host) {
    if (!ctx) return MR_E_BADARG;
    MR_CUDA(ctx, cudaFreeHost(host));
    return MR_OK;
}
int mr_copy(mr_context* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return MR_E_BADARG;
    MR_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return MR_OK;
}
int mr_fill_zero(mr_context* ctx, void* dev, size_t bytes) {
    if (!ctx || (!dev && bytes)) return MR_E_BADARG;
    MR_CUDA(ctx, cudaMemsetAsync(dev, 0, bytes, ctx->stream));
    return MR_OK;
}

int mr_layout_preset(int which, mr_layout* out) {
    if (!out) return MR_E_BADARG;
    memset(out, 0, sizeof(*out));
    out->stride = 32;
    out->nattr = 2;
    out->attr[0].location = 0;
    out->attr[1].location = 1;
    switch (which) {
        case MR_LAYOUT_GPUVERTEX_DECL:
            out->attr[0] = {0, 2, 0};
            out->attr[1] = {16, 3, 1};
            return MR_OK;
        case MR_LAYOUT_GPUVERTEX_ZIGAUTO:
            out->attr[0] = {16, 2, 0};
            out->attr[1] = {0, 3, 1};
            return MR_OK;
        case MR_LAYOUT_TERRAINVERTEX:
            out->attr[0] = {0, 3, 0};
            out->attr[1] = {16, 3, 1};
            return MR_OK;
        default: return MR_E_BADARG;
    }
}

int mr_terrain_params_default(mr_terrain_params* out) {
    if (!out) return MR_E_BADARG;
    out->grid_step = 0.2f;     // Terrain.zig:36
    out->origin_scale = 0.1f;  // Terrain.zig:36
    out->height_scale = 5.0f;  // Terrain.zig:48
    return MR_OK;
}

int mr_terrain_describe(uint32_t n, const mr_terrain_params* params, float bbox_min[3], float bbox_max[3],
                        uint64_t* vertex_count, uint64_t* index_count) {
    mr_terrain_params p;
    if (params) p = *params; else mr_terrain_params_default(&p);
    if (n == 0) return MR_E_BADARG;
    const float bound = (float)n * p.origin_scale;  // Terrain.zig:103-104
    if (bbox_min) {
        bbox_min[0] = -bound;
        bbox_min[1] = 0.0f;
        bbox_min[2] = -bound;
    }
    if (bbox_max) {  // Terrain.zig:109: (bound, 5.0, bound); 5.0 is height_scale * 1.0
        bbox_max[0] = bound;
        bbox_max[1] = p.height_scale;
        bbox_max[2] = bound;
    }
    if (vertex_count) *vertex_count = (uint64_t)n * n;
    if (index_count) *index_count = 6ull * (uint64_t)(n - 1) * (uint64_t)(n - 1);
    return MR_OK;
}

int mr_terrain_build(mr_context* ctx, const mr_terrain_job* job) {
    if (!ctx || !job) return MR_E_BADARG;
    const mr_terrain_job& j = *job;
    if (j.n == 0 || j.n > 65535u || !j.height) return mr_fail(ctx, MR_E_BADARG, "terrain: n must be 1..65535 and height non-null");
    if (j.height_fmt != MR_HEIGHT_U16 && j.height_fmt != MR_HEIGHT_F32) return mr_fail(ctx, MR_E_BADARG, "terrain: bad height_fmt");
    if (j.row_begin > j.row_end || j.row_end > j.n) return mr_fail(ctx, MR_E_BADARG, "terrain: bad row range");
    if (j.qrow_begin > j.qrow_end || j.qrow_end > j.n - 1u) return mr_fail(ctx, MR_E_BADARG, "terrain: bad quad-row range");
    const bool want_v = j.vtx_out && j.row_end > j.row_begin;
    const bool want_i = j.idx_out && j.n > 1 && j.qrow_end > j.qrow_begin;
    if (j.vtx_out && !layout_ok(&j.layout, 1, 3)) return mr_fail(ctx, MR_E_BADARG, "terrain: bad vertex layout");
    if (j.vtx_out && j.layout.nattr > 1 && j.layout.attr[1].ncomp < 3) return mr_fail(ctx, MR_E_BADARG, "terrain: normal needs 3 components");
    if (j.vtx_out && j.vtx_row0 > j.row_begin) return mr_fail(ctx, MR_E_BADARG, "terrain: vtx_row0 > row_begin");
    if (j.idx_out && j.idx_qrow0 > j.qrow_begin) return mr_fail(ctx, MR_E_BADARG, "terrain: idx_qrow0 > qrow_begin");
    if (j.vtx_out && (reinterpret_cast<uintptr_t>(j.vtx_out) & 3u)) return mr_fail(ctx, MR_E_BADARG, "terrain: vtx_out must be 4-byte aligned");
    if (j.idx_out && (reinterpret_cast<uintptr_t>(j.idx_out) & 7u)) return mr_fail(ctx, MR_E_BADARG, "terrain: idx_out must be 8-byte aligned");
    if (want_v) {
        const uint32_t need_lo = j.row_begin > 0 ? j.row_begin - 1 : 0;
        const uint32_t need_hi = std::min(j.row_end + 1, j.n);
        if (j.height_row0 > need_lo || (uint64_t)j.height_row0 + j.height_rows < need_hi)
            return mr_fail(ctx, MR_E_BADARG, "terrain: height rows do not cover the band plus halo");
    }
    if (!want_v && !want_i) return MR_OK;  // empty band: nothing to produce
    MR_CUDA(ctx, cudaSetDevice(ctx->device));

    mr_terrain_job d = j;
    if (!want_v) d.vtx_out = nullptr;
    if (!want_i) d.idx_out = nullptr;
    bool vs = false, is = false;
    void* vdev = nullptr;
    void* idev = nullptr;
    const size_t row_bytes = (size_t)j.n * j.layout.stride;
    const size_t qrow_bytes = 6u * (size_t)(j.n - 1u) * 4u;
    int rc;
    // The index buffer does not depend on the heightmap.  When it has to travel to a host buffer, it is built and copied
    // on a side stream first, so that its device-to-host copy overlaps the heightmap's host-to-device copy (PCIe is full
    // duplex) and the vertex kernel, instead of both copies queueing behind both kernels.
    cudaStream_t idx_stream = ctx->stream;
    if (want_i) {
        // staged output holds quad rows [qrow_begin, qrow_end) only (device outputs keep the caller's idx_qrow0 origin)
        rc = mr_stage_out(ctx, 5, j.idx_out, (size_t)(j.qrow_end - j.qrow_begin) * qrow_bytes, &idev, &is);
        if (rc) return rc;
        if (is) {
            d.idx_out = static_cast<uint32_t*>(idev);
            d.idx_qrow0 = j.qrow_begin;
            if (want_v && !mr_is_device_ptr(j.vtx_out)) {
                if (mr_aux_streams(ctx)) return mr_fail(ctx, MR_E_CUDA, "side streams");
                MR_CUDA(ctx, cudaEventRecord(ctx->fork_ev, ctx->stream));  // order after earlier work on the context
                MR_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[0], ctx->fork_ev, 0));
                idx_stream = ctx->aux[0];
            }
        }
    }
    if (idx_stream != ctx->stream) {
        mr_terrain_job di = d;
        di.vtx_out = nullptr;
        rc = mr_terrain_build_impl(ctx, &di, idx_stream);
        if (rc) return rc;
        unsigned char* dst = reinterpret_cast<unsigned char*>(j.idx_out) + (size_t)(j.qrow_begin - j.idx_qrow0) * qrow_bytes;
        MR_CUDA(ctx, cudaMemcpyAsync(dst, idev, (size_t)(j.qrow_end - j.qrow_begin) * qrow_bytes, cudaMemcpyDeviceToHost, idx_stream));
        MR_CUDA(ctx, cudaEventRecord(ctx->join_ev[0], idx_stream));
        d.idx_out = nullptr;
    }
    if (want_v) {
        const size_t texel = j.height_fmt == MR_HEIGHT_U16 ? 2 : 4;
        const void* hdev = nullptr;
        rc = mr_stage_in(ctx, 0, j.height, (size_t)j.height_rows * j.n * texel, &hdev);
        if (rc) return rc;
        d.height = hdev;
        rc = mr_stage_out(ctx, 4, j.vtx_out, (size_t)(j.row_end - j.row_begin) * row_bytes, &vdev, &vs);
        if (rc) return rc;
        if (vs) {
            d.vtx_out = vdev;
            d.vtx_row0 = j.row_begin;
        }
    }
    rc = mr_terrain_build_impl(ctx, &d, ctx->stream);
    if (rc) return rc;
    if (vs) {
        rc = mr_copy_back(ctx, static_cast<unsigned char*>(j.vtx_out) + (size_t)(j.row_begin - j.vtx_row0) * row_bytes, vdev,
                          (size_t)(j.row_end - j.row_begin) * row_bytes);
        if (rc) return rc;
    }
    if (is && idx_stream == ctx->stream) {
        rc = mr_copy_back(ctx, reinterpret_cast<unsigned char*>(j.idx_out) + (size_t)(j.qrow_begin - j.idx_qrow0) * qrow_bytes, idev,
                          (size_t)(j.qrow_end - j.qrow_begin) * qrow_bytes);
        if (rc) return rc;
    }
    if (idx_stream != ctx->stream) MR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->join_ev[0], 0));
    return mr_finish_host_io(ctx);
}

int mr_terrain_build_full(mr_context* ctx, const void* height, uint32_t height_fmt, uint32_t n, const mr_layout* layout,
                          const mr_terrain_params* params, void* vtx_out, uint32_t* idx_out) {
    if (!ctx) return MR_E_BADARG;
    mr_terrain_job j;
    memset(&j, 0, sizeof(j));
    j.n = n;
    j.height_fmt = height_fmt;
    j.height = height;
    j.height_row0 = 0;
    j.height_rows = n;
    j.row_begin = 0;
    j.row_end = n;
    j.vtx_out = vtx_out;
    j.vtx_row0 = 0;
    j.qrow_begin = 0;
    j.qrow_end = n ? n - 1 : 0;
    j.idx_out = idx_out;
    j.idx_qrow0 = 0;
    if (layout) j.layout = *layout; else mr_layout_preset(MR_LAYOUT_TERRAINVERTEX, &j.layout);
    if (params) j.params = *params; else mr_terrain_params_default(&j.params);
    return mr_terrain_build(ctx, &j);
}

int mr_terrain_tile_count(uint32_t n, uint32_t tile_rows, uint32_t tile_cols, uint32_t* tiles_r_out, uint32_t* tiles_c_out) {
    if (n < 2 || tile_rows == 0 || tile_cols == 0) return MR_E_BADARG;
    if (tiles_r_out) *tiles_r_out = (n - 1u + tile_rows - 1u) / tile_rows;
    if (tiles_c_out) *tiles_c_out = (n - 1u + tile_cols - 1u) / tile_cols;
    return MR_OK;
}

int mr_terrain_tile_bounds(mr_context* ctx, const void* height, uint32_t height_fmt, uint32_t n, uint32_t tile_rows,
                           uint32_t tile_cols, const mr_terrain_params* params, float* bbox_out) {
    if (!ctx || !height || !bbox_out) return MR_E_BADARG;
    uint32_t tr = 0, tc = 0;
    if (n > 65535u || mr_terrain_tile_count(n, tile_rows, tile_cols, &tr, &tc) != MR_OK) return mr_fail(ctx, MR_E_BADARG, "tiles: bad n or tile size");
    if (height_fmt != MR_HEIGHT_U16 && height_fmt != MR_HEIGHT_F32) return mr_fail(ctx, MR_E_BADARG, "tiles: bad height_fmt");
    if (reinterpret_cast<uintptr_t>(bbox_out) & 15u) return mr_fail(ctx, MR_E_BADARG, "tiles: bbox_out must be 16-byte aligned");
    mr_terrain_params p;
    if (params) p = *params; else mr_terrain_params_default(&p);
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* hdev = nullptr;
    int rc = mr_stage_in(ctx, 0, height, (size_t)n * n * (height_fmt == MR_HEIGHT_U16 ? 2 : 4), &hdev);
    if (rc) return rc;
    void* bdev = nullptr;
    bool st = false;
    const size_t bytes = (size_t)tr * tc * 32;
    rc = mr_stage_out(ctx, 4, bbox_out, bytes, &bdev, &st);
    if (rc) return rc;
    rc = mr_terrain_tile_bounds_impl(ctx, hdev, height_fmt, n, tile_rows, tile_cols, &p, static_cast<float*>(bdev));
    if (rc) return rc;
    if (st) {
        rc = mr_copy_back(ctx, bbox_out, bdev, bytes);
        if (rc) return rc;
    }
    return mr_finish_host_io(ctx);
}

int mr_terrain_cull(mr_context* ctx, const float* bbox, uint32_t n, uint32_t tile_rows, uint32_t tile_cols, const float xform[16],
                    uint32_t* visible_out, uint32_t* visible_ids_out, uint32_t* idx_out, uint64_t* counts_out) {
    if (!ctx || !bbox || !xform) return MR_E_BADARG;
    uint32_t tr = 0, tc = 0;
    if (n > 65535u || mr_terrain_tile_count(n, tile_rows, tile_cols, &tr, &tc) != MR_OK) return mr_fail(ctx, MR_E_BADARG, "cull: bad n or tile size");
    if (idx_out && (reinterpret_cast<uintptr_t>(idx_out) & 7u)) return mr_fail(ctx, MR_E_BADARG, "cull: idx_out must be 8-byte aligned");
    const size_t ntiles = (size_t)tr * tc;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* bdev = nullptr;
    int rc = mr_stage_in(ctx, 0, bbox, ntiles * 32, &bdev);
    if (rc) return rc;
    void *vdev = nullptr, *idev = nullptr, *xdev = nullptr, *cdev = nullptr, *work = nullptr;
    bool sv = false, si = false, sx = false, sc = false;
    if (visible_out) { rc = mr_stage_out(ctx, 1, visible_out, ntiles * 4, &vdev, &sv); if (rc) return rc; }
    // visible ids and per-slot first index are needed internally even when the caller does not ask for the ids
    rc = mr_scratch(ctx, 10, ntiles * 12 + 16, &work);
    if (rc) return rc;
    unsigned long long* first_index = static_cast<unsigned long long*>(work);
    uint32_t* ids_scratch = reinterpret_cast<uint32_t*>(first_index + ntiles);
    if (visible_ids_out) { rc = mr_stage_out(ctx, 2, visible_ids_out, ntiles * 4, &idev, &si); if (rc) return rc; }
    else idev = ids_scratch;
    const size_t max_idx_bytes = 24ull * (size_t)(n - 1u) * (n - 1u);
    if (idx_out) { rc = mr_stage_out(ctx, 5, idx_out, max_idx_bytes, &xdev, &sx); if (rc) return rc; }
    if (counts_out) { rc = mr_stage_out(ctx, 3, counts_out, 16, &cdev, &sc); if (rc) return rc; }
    else { rc = mr_scratch(ctx, 3, 16, &cdev); if (rc) return rc; }
    rc = mr_terrain_cull_impl(ctx, static_cast<const float*>(bdev), n, tile_rows, tile_cols, xform, static_cast<uint32_t*>(vdev),
                              static_cast<uint32_t*>(idev), first_index, static_cast<uint32_t*>(xdev),
                              static_cast<unsigned long long*>(cdev));
    if (rc) return rc;
    if (sv) { rc = mr_copy_back(ctx, visible_out, vdev, ntiles * 4); if (rc) return rc; }
    uint64_t counts[2] = {0, 0};
    if (si || sx || sc) {  // host outputs: only the filled prefixes travel
        MR_CUDA(ctx, cudaMemcpyAsync(counts, cdev, 16, cudaMemcpyDeviceToHost, ctx->stream));
        MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (sc) memcpy(counts_out, counts, 16);
        if (si && counts[0]) { rc = mr_copy_back(ctx, visible_ids_out, idev, (size_t)counts[0] * 4); if (rc) return rc; }
        if (sx && counts[1]) { rc = mr_copy_back(ctx, idx_out, xdev, (size_t)counts[1] * 4); if (rc) return rc; }
    }
    return mr_finish_host_io(ctx);
}

int mr_selftest_fastdiv(mr_context* ctx, float divisor, int force_fast, uint64_t* mismatches_out) {
    if (!ctx || !mismatches_out) return MR_E_BADARG;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    void* d = nullptr;
    int rc = mr_scratch(ctx, 10, 8, &d);
    if (rc) return rc;
    MR_CUDA(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
    rc = mr_selftest_fastdiv_impl(ctx, divisor, force_fast, static_cast<unsigned long long*>(d));
    if (rc) return rc;
    MR_CUDA(ctx, cudaMemcpyAsync(mismatches_out, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->host_io = false;
    return MR_OK;
}

int mr_heightmap_normalize(mr_context* ctx, const uint16_t* in, uint64_t count, float* out) {
    if (!ctx || (count && (!in || !out))) return MR_E_BADARG;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* din = nullptr;
    int rc = mr_stage_in(ctx, 0, in, (size_t)count * 2, &din);
    if (rc) return rc;
    void* dout = nullptr;
    bool st = false;
    rc = mr_stage_out(ctx, 4, out, (size_t)count * 4, &dout, &st);
    if (rc) return rc;
    rc = mr_heightmap_normalize_impl(ctx, static_cast<const uint16_t*>(din), count, static_cast<float*>(dout));
    if (rc) return rc;
    if (st) {
        rc = mr_copy_back(ctx, out, dout, (size_t)count * 4);
        if (rc) return rc;
    }
    return mr_finish_host_io(ctx);
}

int mr_polygon_offsets(mr_context* ctx, const uint64_t* first_point, uint32_t npoly, uint64_t* first_tri_out) {
    if (!ctx || !first_point || !first_tri_out) return MR_E_BADARG;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* dfp = nullptr;
    int rc = mr_stage_in(ctx, 1, first_point, (size_t)(npoly + 1) * 8, &dfp);
    if (rc) return rc;
    void* dft = nullptr;
    bool st = false;
    rc = mr_stage_out(ctx, 2, first_tri_out, (size_t)(npoly + 1) * 8, &dft, &st);
    if (rc) return rc;
    rc = mr_polygon_offsets_impl(ctx, static_cast<const uint64_t*>(dfp), npoly, static_cast<uint64_t*>(dft));
    if (rc) return rc;
    if (st) {
        rc = mr_copy_back(ctx, first_tri_out, dft, (size_t)(npoly + 1) * 8);
        if (rc) return rc;
    }
    return mr_finish_host_io(ctx);
}

int mr_polygon_draw_range(uint64_t first_tri_i, uint64_t first_tri_next, uint64_t tri_base, mr_draw_range* out) {
    if (!out || first_tri_next < first_tri_i || first_tri_i < tri_base) return MR_E_BADARG;
    const uint64_t prims = first_tri_next - first_tri_i, off = first_tri_i - tri_base;
    if (prims * 3 > 0xFFFFFFFFull || off * 3 > 0xFFFFFFFFull) return MR_E_BADARG;  // VertexBuffer fields are u32
    out->vertex_count = (uint32_t)(prims * 3);  // VertexBuffer.zig:21
    out->instance_count = 1;
    out->first_vertex = (uint32_t)(off * 3);  // VertexBuffer.zig:22
    out->first_instance = 0;
    return MR_OK;
}

int mr_triangulate_batch(mr_context* ctx, const mr_polygon_job* job) {
    if (!ctx || !job) return MR_E_BADARG;
    const mr_polygon_job& j = *job;
    if (j.npoly == 0) return MR_OK;
    if (!j.xy || !j.first_point || !j.first_tri || !j.vtx_out) return mr_fail(ctx, MR_E_BADARG, "polygons: null pointer");
    if (!layout_ok(&j.layout, 1, 2)) return mr_fail(ctx, MR_E_BADARG, "polygons: bad vertex layout");
    if (j.layout.nattr > 1 && j.layout.attr[1].ncomp < 3) return mr_fail(ctx, MR_E_BADARG, "polygons: colour needs 3 components");
    if (reinterpret_cast<uintptr_t>(j.vtx_out) & 3u) return mr_fail(ctx, MR_E_BADARG, "polygons: vtx_out must be 4-byte aligned");
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    {   // the reference's own call shape (Polygon.create_polygon: one small polygon, everything in host memory)
        int rc_small = MR_OK;
        if (mr_triangulate_small(ctx, &j, &rc_small)) return rc_small;
    }

    mr_polygon_job d = j;
    const bool fp_dev = mr_is_device_ptr(j.first_point);
    const bool ft_dev = mr_is_device_ptr(j.first_tri);
    // sizes of the data arrays: known from first_point/first_tri when those live on the host
    uint64_t npts = 0, ntri = 0;
    bool need_sizes = !mr_is_device_ptr(j.xy) || !mr_is_device_ptr(j.vtx_out);
    if (need_sizes) {
        uint64_t fp_ends[2], ft_ends[2];
        if (fp_dev) {
            MR_CUDA(ctx, cudaMemcpyAsync(&fp_ends[0], j.first_point, 8, cudaMemcpyDeviceToHost, ctx->stream));
            MR_CUDA(ctx, cudaMemcpyAsync(&fp_ends[1], j.first_point + j.npoly, 8, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            fp_ends[0] = j.first_point[0];
            fp_ends[1] = j.first_point[j.npoly];
        }
        if (ft_dev) {
            MR_CUDA(ctx, cudaMemcpyAsync(&ft_ends[0], j.first_tri, 8, cudaMemcpyDeviceToHost, ctx->stream));
            MR_CUDA(ctx, cudaMemcpyAsync(&ft_ends[1], j.first_tri + j.npoly, 8, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            ft_ends[0] = j.first_tri[0];
            ft_ends[1] = j.first_tri[j.npoly];
        }
        if (fp_dev || ft_dev) MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (fp_ends[0] < j.point_base || ft_ends[0] < j.tri_base) return mr_fail(ctx, MR_E_BADARG, "polygons: base beyond first offset");
        npts = fp_ends[1] - j.point_base;   // xy covers [point_base, first_point[npoly])
        ntri = ft_ends[1] - j.tri_base;
    }
    const void* p = nullptr;
    int rc;
    rc = mr_stage_in(ctx, 0, j.xy, (size_t)npts * 8, &p);
    if (rc) return rc;
    d.xy = static_cast<const float*>(p);
    rc = mr_stage_in(ctx, 1, j.first_point, (size_t)(j.npoly + 1) * 8, &p);
    if (rc) return rc;
    d.first_point = static_cast<const uint64_t*>(p);
    rc = mr_stage_in(ctx, 2, j.first_tri, (size_t)(j.npoly + 1) * 8, &p);
    if (rc) return rc;
    d.first_tri = static_cast<const uint64_t*>(p);
    if (j.offset_prime) {
        rc = mr_stage_in(ctx, 3, j.offset_prime, (size_t)j.npoly * 8, &p);
        if (rc) return rc;
        d.offset_prime = static_cast<const uint32_t*>(p);
    }
    bool sv = false, sb = false, ss = false, sn = false;
    void* q = nullptr;
    const size_t vbytes = (size_t)ntri * 3u * j.layout.stride;
    // The vertex range is written over the whole kernel run: into pinned host memory the kernels store
    // directly (zero-copy over PCIe, overlapping the transfer with the triangulation; measured 10.3 vs
    // 12.3 ms for the 100k batch); pageable host memory is staged and copied back.
    if (void* alias = mr_pinned_device_alias(j.vtx_out)) {
        q = alias;
        ctx->host_io = true;  // the kernels store into the caller's host buffer: the call returns after they finish
    } else {
        rc = mr_stage_out(ctx, 4, j.vtx_out, vbytes, &q, &sv);
        if (rc) return rc;
    }
    d.vtx_out = q;
    if (j.bbox_out) {
        rc = mr_stage_out(ctx, 5, j.bbox_out, (size_t)j.npoly * 16, &q, &sb);
        if (rc) return rc;
        d.bbox_out = static_cast<float*>(q);
    }
    if (j.status_out) {
        rc = mr_stage_out(ctx, 6, j.status_out, (size_t)j.npoly * 4, &q, &ss);
        if (rc) return rc;
        d.status_out = static_cast<uint32_t*>(q);
    }
    if (j.ntri_out) {
        rc = mr_stage_out(ctx, 7, j.ntri_out, (size_t)j.npoly * 4, &q, &sn);
        if (rc) return rc;
        d.ntri_out = static_cast<uint32_t*>(q);
    }
    rc = mr_triangulate_impl(ctx, &d, fp_dev ? nullptr : j.first_point);
    if (rc) return rc;
    if (sv) {
        // the polygons of this job own [first_tri[0], first_tri[npoly]) only
        uint64_t ft0 = ft_dev ? 0 : j.first_tri[0];
        if (ft_dev) {
            MR_CUDA(ctx, cudaMemcpyAsync(&ft0, j.first_tri, 8, cudaMemcpyDeviceToHost, ctx->stream));
            MR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        const size_t off = (size_t)(ft0 - j.tri_base) * 3u * j.layout.stride;
        rc = mr_copy_back(ctx, static_cast<unsigned char*>(j.vtx_out) + off, static_cast<unsigned char*>(d.vtx_out) + off, vbytes - off);
        if (rc) return rc;
    }
    if (sb) { rc = mr_copy_back(ctx, j.bbox_out, d.bbox_out, (size_t)j.npoly * 16); if (rc) return rc; }
    if (ss) { rc = mr_copy_back(ctx, j.status_out, d.status_out, (size_t)j.npoly * 4); if (rc) return rc; }
    if (sn) { rc = mr_copy_back(ctx, j.ntri_out, d.ntri_out, (size_t)j.npoly * 4); if (rc) return rc; }
    return mr_finish_host_io(ctx);
}

int mr_triangulate_tier_counts(mr_context* ctx, uint32_t out[8]) {
    if (!ctx || !out) return MR_E_BADARG;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    return mr_triangulate_tier_counts_impl(ctx, out);
}

uint64_t mr_rng_state0(uint64_t seed, uint64_t index) { return mr_rng_state0_hd(seed, index); }

uint32_t mr_rng_u32(uint64_t* state) {
    // sequential form of the same stream: draw k of state0 == k-th call starting from state0
    const uint32_t v = mr_rng_draw(*state, 0u);
    *state += MR_GOLDEN;
    return v;
}

// unirand.zig:26-50, host restatement (same draws as the device port)
int mr_unirand_seed_host(uint32_t top, uint64_t seed, uint64_t index, uint32_t* offset_out, uint32_t* prime_out) {
    if (!offset_out || !prime_out) return MR_E_BADARG;
    if (top == 1u) {
        *offset_out = 0;
        *prime_out = 1;
        return MR_OK;
    }
    uint64_t st = mr_rng_state0_hd(seed, index);
    *offset_out = mr_rng_u32(&st) % (uint32_t)(top - 1u) + 1u;
    uint32_t best = 1;
    for (int i = 0; i < MR_NPRIMES; ++i) {
        const uint32_t p = k_primes_host[i];
        if (p < top && top % p != 0u) {
            if (mr_rng_u32(&st) % 3u > 0u) best = p;
        }
    }
    *prime_out = best;
    return MR_OK;
}

int mr_unirand_seed_batch(mr_context* ctx, const uint64_t* first_point, uint32_t npoly, uint64_t seed,
                          uint64_t poly_index0, uint32_t* offset_prime_out) {
    if (!ctx || !first_point || !offset_prime_out) return MR_E_BADARG;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* dfp = nullptr;
    int rc = mr_stage_in(ctx, 1, first_point, (size_t)(npoly + 1) * 8, &dfp);
    if (rc) return rc;
    void* dout = nullptr;
    bool st = false;
    rc = mr_stage_out(ctx, 3, offset_prime_out, (size_t)npoly * 8, &dout, &st);
    if (rc) return rc;
    rc = mr_unirand_seed_batch_impl(ctx, static_cast<const uint64_t*>(dfp), npoly, seed, poly_index0,
                                    static_cast<uint32_t*>(dout));
    if (rc) return rc;
    if (st) {
        rc = mr_copy_back(ctx, offset_prime_out, dout, (size_t)npoly * 8);
        if (rc) return rc;
    }
    return mr_finish_host_io(ctx);
}

int mr_synth_heightmap_u16(mr_context* ctx, uint64_t seed, uint32_t n, uint32_t row0, uint32_t rows, uint16_t* out) {
    if (!ctx || !out) return MR_E_BADARG;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    void* dout = nullptr;
    bool st = false;
    const size_t bytes = (size_t)rows * n * 2;
    int rc = mr_stage_out(ctx, 4, out, bytes, &dout, &st);
    if (rc) return rc;
    rc = mr_synth_heightmap_u16_impl(ctx, seed, n, row0, rows, static_cast<uint16_t*>(dout));
    if (rc) return rc;
    if (st) {
        rc = mr_copy_back(ctx, out, dout, bytes);
        if (rc) return rc;
    }
    return mr_finish_host_io(ctx);
}

int mr_synth_polygons(mr_context* ctx, uint64_t seed, uint64_t poly_index0, const uint64_t* first_point,
                      uint32_t npoly, float* xy_out) {
    return mr_synth_polygons_family(ctx, MR_FAMILY_STAR, seed, poly_index0, first_point, npoly, xy_out);
}

int mr_synth_polygons_family(mr_context* ctx, int family, uint64_t seed, uint64_t poly_index0, const uint64_t* first_point,
                             uint32_t npoly, float* xy_out) {
    if (!ctx || !first_point || !xy_out) return MR_E_BADARG;
    if (family != MR_FAMILY_STAR && family != MR_FAMILY_ELLIPSE && family != MR_FAMILY_ZIPPER)
        return mr_fail(ctx, MR_E_BADARG, "synth_polygons: unknown family");
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* dfp = nullptr;
    int rc = mr_stage_in(ctx, 1, first_point, (size_t)(npoly + 1) * 8, &dfp);
    if (rc) return rc;
    void* dout = nullptr;
    bool st = false;
    size_t bytes = 0;
    if (!mr_is_device_ptr(xy_out)) {
        if (mr_is_device_ptr(first_point)) return mr_fail(ctx, MR_E_BADARG, "synth_polygons: host xy_out needs host first_point");
        bytes = (size_t)(first_point[npoly] - first_point[0]) * 8;
    }
    rc = mr_stage_out(ctx, 0, xy_out, bytes, &dout, &st);
    if (rc) return rc;
    rc = mr_synth_polygons_impl(ctx, family, seed, poly_index0, static_cast<const uint64_t*>(dfp), npoly, static_cast<float*>(dout));
    if (rc) return rc;
    if (st) {
        rc = mr_copy_back(ctx, xy_out, dout, bytes);
        if (rc) return rc;
    }
    return mr_finish_host_io(ctx);
}

int mr_ipc_export(mr_context* ctx, void* dev, unsigned char handle_out[MR_IPC_HANDLE_BYTES]) {
    if (!ctx || !dev || !handle_out) return MR_E_BADARG;
    static_assert(sizeof(cudaIpcMemHandle_t) <= MR_IPC_HANDLE_BYTES, "ipc handle size");
    cudaIpcMemHandle_t h;
    MR_CUDA(ctx, cudaIpcGetMemHandle(&h, dev));
    memset(handle_out, 0, MR_IPC_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    return MR_OK;
}
int mr_ipc_open(mr_context* ctx, const unsigned char handle[MR_IPC_HANDLE_BYTES], void** dev_out) {
    if (!ctx || !handle || !dev_out) return MR_E_BADARG;
    MR_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    MR_CUDA(ctx, cudaIpcOpenMemHandle(dev_out, h, cudaIpcMemLazyEnablePeerAccess));
    return MR_OK;
}
int mr_ipc_close(mr_context* ctx, void* dev) {
    if (!ctx || !dev) return MR_E_BADARG;
    MR_CUDA(ctx, cudaIpcCloseMemHandle(dev));
    return MR_OK;
}

// Contiguous split of polygons over ranks, balanced by the cost model w(n) = n*log2(n) + n
// (descent work of the trapezoidation) -- SURVEY 8-e.
int mr_polygon_partition(const uint64_t* first_point, uint32_t npoly, uint32_t nranks, uint32_t* range_out) {
    if (!first_point || !range_out || nranks == 0) return MR_E_BADARG;
    std::vector<double> acc((size_t)npoly + 1, 0.0);
    for (uint32_t i = 0; i < npoly; ++i) {
        const double n = (double)(first_point[i + 1] - first_point[i]);
        const double w = n > 1.0 ? n * std::log2(n) + n : 1.0;
        acc[i + 1] = acc[i] + w;
    }
    const double total = acc[npoly];
    range_out[0] = 0;
    uint32_t cur = 0;
    for (uint32_t r = 1; r < nranks; ++r) {
        const double target = total * (double)r / (double)nranks;
        while (cur < npoly && acc[cur + 1] <= target) ++cur;
        range_out[r] = cur;
    }
    range_out[nranks] = npoly;
    return MR_OK;
}

int mr_terrain_partition(uint32_t n, uint32_t nranks, uint32_t* rows_out, uint32_t* qrows_out) {
    if (n == 0 || nranks == 0) return MR_E_BADARG;
    for (uint32_t r = 0; r <= nranks; ++r) {
        if (rows_out) rows_out[r] = (uint32_t)((uint64_t)n * r / nranks);
        if (qrows_out) qrows_out[r] = (uint32_t)((uint64_t)(n - 1) * r / nranks);
    }
    return MR_OK;
}

}  // extern "C"
