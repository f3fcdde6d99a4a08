/*
 * o_terrain.c -- oracle for the terrain mesh.  TEST INFRASTRUCTURE ONLY (see mr_oracle.h).
 *
 * Reference code followed:
 *   heightmap normalisation   Terrain/Terrain.zig:114-124 (formula :120)
 *   vertex position formula   Terrain/Terrain.zig:24-48 (WGSL, per shader vertex)
 *   corner order / winding    Terrain/Terrain.zig:28-35, Renderer/Pipeline.zig:145-149
 * NEW SPEC (no reference code exists; parity unpinned, defined here once):
 *   the indexed form of the mesh, the u32 index buffer and the clamped
 *   central-difference normals -- see include/myrenderer_b200.h for the formulas.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "mr_oracle.h"

/* Terrain.zig:120 */
static float norm_u16(uint16_t v) { return 1.0f - (float)v / 65535.0f; }

void mr_o_heightmap_normalize(const uint16_t* in, uint64_t count, float* out) {
    uint64_t i;
    for (i = 0; i < count; ++i) out[i] = norm_u16(in[i]);
}

static float texel(const mr_terrain_job* j, uint32_t r, uint32_t c) {
    size_t idx = (size_t)(r - j->height_row0) * j->n + c;
    if (j->height_fmt == MR_HEIGHT_U16) return norm_u16(((const uint16_t*)j->height)[idx]);
    return ((const float*)j->height)[idx];
}

/* Terrain.zig:24-48 evaluated literally for one shader vertex index */
int mr_o_terrain_shader_vertex(const float* h, uint32_t n, uint64_t vi, const mr_terrain_params* p,
                               float out4[4]) {
    static const float quad_vals[6][2] = {{1, 0}, {0, 0}, {1, 1}, {1, 1}, {0, 0}, {0, 1}};
    uint64_t vertex_at = vi % 6u;                       /* :24 */
    uint64_t quad_at = (vi - vertex_at) / 6u;           /* :25 */
    float qr = (float)(uint32_t)(quad_at / n);          /* :26 */
    float qc = (float)(uint32_t)(quad_at % n);
    float size_f = (float)n;
    float vx = p->grid_step * (quad_vals[vertex_at][0] + qr) - p->origin_scale * size_f; /* :36 */
    float vz = p->grid_step * (quad_vals[vertex_at][1] + qc) - p->origin_scale * size_f;
    uint64_t lookup[6];
    lookup[0] = quad_at + n; /* :38-45 */
    lookup[1] = quad_at;
    lookup[2] = quad_at + n + 1u;
    lookup[3] = quad_at + n + 1u;
    lookup[4] = quad_at;
    lookup[5] = quad_at + 1u;
    if (lookup[vertex_at] >= (uint64_t)n * n) return 0;
    out4[0] = vx;
    out4[1] = p->height_scale * h[lookup[vertex_at]]; /* :47-48 */
    out4[2] = vz;
    out4[3] = 1.0f;
    return 1;
}

typedef struct band {
    const mr_terrain_job* j;
    uint32_t r0, r1, q0, q1;
} band;

static void build_band(const band* b) {
    const mr_terrain_job* j = b->j;
    const uint32_t n = j->n;
    const float gs = j->params.grid_step, os = j->params.origin_scale, hs = j->params.height_scale;
    const float org = os * (float)n;
    uint32_t r, c;
    if (j->vtx_out) {
        const int has_normal = j->layout.nattr > 1;
        for (r = b->r0; r < b->r1; ++r) {
            uint32_t rm = r > 0 ? r - 1u : 0u, rp = r + 1u < n ? r + 1u : n - 1u;
            float x = gs * (float)r - org;
            unsigned char* row =
                (unsigned char*)j->vtx_out + (size_t)(r - j->vtx_row0) * n * j->layout.stride;
            for (c = 0; c < n; ++c) {
                unsigned char* v = row + (size_t)c * j->layout.stride;
                float pos[3];
                memset(v, 0, j->layout.stride);
                pos[0] = x;
                pos[1] = hs * texel(j, r, c);
                pos[2] = gs * (float)c - org;
                memcpy(v + j->layout.attr[0].offset, pos, 12);
                if (has_normal) {
                    uint32_t cm = c > 0 ? c - 1u : 0u, cp = c + 1u < n ? c + 1u : n - 1u;
                    float gx = 0.0f, gz = 0.0f, len, inv, nrm[3];
                    if (rp != rm) gx = (hs * (texel(j, rp, c) - texel(j, rm, c))) / (gs * (float)(rp - rm));
                    if (cp != cm) gz = (hs * (texel(j, r, cp) - texel(j, r, cm))) / (gs * (float)(cp - cm));
                    len = sqrtf(((gx * gx) + 1.0f) + (gz * gz));
                    inv = 1.0f / len;
                    nrm[0] = (-gx) * inv;
                    nrm[1] = inv;
                    nrm[2] = (-gz) * inv;
                    memcpy(v + j->layout.attr[1].offset, nrm, 12);
                }
            }
        }
    }
    if (j->idx_out && n > 1) {
        for (r = b->q0; r < b->q1; ++r) {
            uint32_t* o = j->idx_out + (size_t)(r - j->idx_qrow0) * 6u * (n - 1u);
            for (c = 0; c + 1u < n; ++c) {
                uint32_t i00 = r * n + c;
                o[0] = i00 + n;      /* (r+1,c)   Terrain.zig:29,39 */
                o[1] = i00;          /* (r,c)     :30,40 */
                o[2] = i00 + n + 1u; /* (r+1,c+1) :31,41 */
                o[3] = i00 + n + 1u; /* (r+1,c+1) :32,42 */
                o[4] = i00;          /* (r,c)     :33,43 */
                o[5] = i00 + 1u;     /* (r,c+1)   :34,44 */
                o += 6;
            }
        }
    }
}

static void* band_main(void* p) {
    build_band((const band*)p);
    return NULL;
}

int mr_o_terrain_build(const mr_terrain_job* j, int nthreads) {
    int T, i;
    band* bs;
    pthread_t* th;
    if (!j || !j->height || j->n == 0) return MR_E_BADARG;
    if (j->row_end > j->n || j->row_begin > j->row_end) return MR_E_BADARG;
    if (j->n > 1 && (j->qrow_end > j->n - 1u || j->qrow_begin > j->qrow_end)) return MR_E_BADARG;
    T = nthreads <= 0 ? mr_o_hardware_threads() : nthreads;
    bs = (band*)calloc((size_t)T, sizeof(band));
    th = (pthread_t*)calloc((size_t)T, sizeof(pthread_t));
    for (i = 0; i < T; ++i) {
        uint64_t nr = j->row_end - j->row_begin, nq = j->qrow_end - j->qrow_begin;
        bs[i].j = j;
        bs[i].r0 = j->row_begin + (uint32_t)(nr * (uint64_t)i / (uint64_t)T);
        bs[i].r1 = j->row_begin + (uint32_t)(nr * (uint64_t)(i + 1) / (uint64_t)T);
        bs[i].q0 = j->qrow_begin + (uint32_t)(nq * (uint64_t)i / (uint64_t)T);
        bs[i].q1 = j->qrow_begin + (uint32_t)(nq * (uint64_t)(i + 1) / (uint64_t)T);
    }
    if (T == 1) {
        build_band(&bs[0]);
    } else {
        for (i = 0; i < T; ++i) pthread_create(&th[i], NULL, band_main, &bs[i]);
        for (i = 0; i < T; ++i) pthread_join(th[i], NULL);
    }
    free(bs);
    free(th);
    return MR_OK;
}
