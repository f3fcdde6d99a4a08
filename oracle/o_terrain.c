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

/* ---- tiles and culling ---------------------------------------------------
 * Tile boxes: Terrain.zig:103-110 computes ONE box for the terrain, (-bound,0,-bound)..(bound,5,bound); here the
 * same kind of box per tile, tight in y (NEW SPEC: the reference does not tile).
 * Visibility: SceneNode.zig:96-110, statement for statement, with mach.math.Mat4x4.mulVec restated from the mach
 * source (mach is not vendored under /root/reference: PARITY UNPINNED for mulVec's evaluation order). */
static void mach_mul_vec(const float m[16], const float v[4], float out[4]) {
    int i, j;
    for (i = 0; i < 4; ++i) {
        float acc = 0.0f;
        for (j = 0; j < 4; ++j) acc = acc + m[4 * j + i] * v[j]; /* result[i] += matrix.v[j].v[i] * vector.v[j] */
        out[i] = acc;
    }
}

static float texel_full(const void* height, uint32_t fmt, uint32_t n, uint32_t r, uint32_t c) {
    size_t idx = (size_t)r * n + c;
    if (fmt == MR_HEIGHT_U16) return norm_u16(((const uint16_t*)height)[idx]);
    return ((const float*)height)[idx];
}

int mr_o_terrain_tile_bounds(const void* height, uint32_t fmt, uint32_t n, uint32_t tile_rows, uint32_t tile_cols,
                             const mr_terrain_params* p, float* bbox_out) {
    uint32_t tiles_r, tiles_c, tr, tc;
    if (!height || !bbox_out || n < 2 || !tile_rows || !tile_cols) return MR_E_BADARG;
    tiles_r = (n - 1u + tile_rows - 1u) / tile_rows;
    tiles_c = (n - 1u + tile_cols - 1u) / tile_cols;
    for (tr = 0; tr < tiles_r; ++tr)
        for (tc = 0; tc < tiles_c; ++tc) {
            uint32_t r0 = tr * tile_rows, c0 = tc * tile_cols, r, c;
            uint32_t r1 = r0 + tile_rows < n - 1u ? r0 + tile_rows : n - 1u;
            uint32_t c1 = c0 + tile_cols < n - 1u ? c0 + tile_cols : n - 1u;
            float lo = INFINITY, hi = -INFINITY, org = p->origin_scale * (float)n;
            float xa, xb, za, zb, ya, yb, y0, y1;
            float* o = bbox_out + 8u * ((size_t)tr * tiles_c + tc);
            for (r = r0; r <= r1; ++r)
                for (c = c0; c <= c1; ++c) {
                    float v = texel_full(height, fmt, n, r, c);
                    if (v < lo) lo = v;
                    if (v > hi) hi = v;
                }
            xa = p->grid_step * (float)r0 - org;
            xb = p->grid_step * (float)r1 - org;
            za = p->grid_step * (float)c0 - org;
            zb = p->grid_step * (float)c1 - org;
            ya = p->height_scale * lo;
            yb = p->height_scale * hi;
            /* explicit compares (first operand wins ties); a zero y bound is written as +0: with height_scale == 0,
             * or a map that holds both zeros, its sign would otherwise depend on the order of the scan */
            y0 = yb < ya ? yb : ya;
            y1 = yb > ya ? yb : ya;
            if (y0 == 0.0f) y0 = 0.0f;
            if (y1 == 0.0f) y1 = 0.0f;
            o[0] = xb < xa ? xb : xa; o[1] = y0; o[2] = zb < za ? zb : za; o[3] = 1.0f;
            o[4] = xb > xa ? xb : xa; o[5] = y1; o[6] = zb > za ? zb : za; o[7] = 1.0f;
        }
    return MR_OK;
}

/* SceneNode.zig:96-110 for one box: returns should_render */
int mr_o_scene_node_should_render(const float xform[16], const float p0_in[4], const float p1_in[4]) {
    float p0[4], p1[4];
    float mn = fminf(fminf(p0_in[0], p0_in[1]), fminf(p0_in[2], p0_in[3]));
    float mx = fmaxf(fmaxf(p1_in[0], p1_in[1]), fmaxf(p1_in[2], p1_in[3]));
    memcpy(p0, p0_in, 16);
    memcpy(p1, p1_in, 16);
    if (mn != -INFINITY) mach_mul_vec(xform, p0_in, p0); /* :99-101 */
    if (mx != INFINITY) mach_mul_vec(xform, p1_in, p1);  /* :103-105 */
    return (p1[0] > 0.0f && p1[1] > 0.0f && p1[2] > 0.0f && p1[3] > 0.0f) ||
           (p0[0] < 1.0f && p0[1] < 1.0f && p0[2] < 1.0f && p0[3] < 1.0f); /* :111 */
}

int mr_o_terrain_cull(const float* bbox, uint32_t n, uint32_t tile_rows, uint32_t tile_cols, const float xform[16],
                      uint32_t* visible_out, uint32_t* visible_ids_out, uint32_t* idx_out, uint64_t counts_out[2]) {
    uint32_t tiles_r, tiles_c, t, nvis = 0;
    uint64_t nidx = 0;
    if (!bbox || n < 2 || !tile_rows || !tile_cols) return MR_E_BADARG;
    tiles_r = (n - 1u + tile_rows - 1u) / tile_rows;
    tiles_c = (n - 1u + tile_cols - 1u) / tile_cols;
    for (t = 0; t < tiles_r * tiles_c; ++t) {
        int vis = mr_o_scene_node_should_render(xform, bbox + 8u * (size_t)t, bbox + 8u * (size_t)t + 4u);
        if (visible_out) visible_out[t] = vis ? 1u : 0u;
        if (vis) {
            uint32_t tr = t / tiles_c, tc = t % tiles_c, r0 = tr * tile_rows, c0 = tc * tile_cols, r, c;
            uint32_t qr = tile_rows < n - 1u - r0 ? tile_rows : n - 1u - r0;
            uint32_t qc = tile_cols < n - 1u - c0 ? tile_cols : n - 1u - c0;
            if (visible_ids_out) visible_ids_out[nvis] = t;
            ++nvis;
            for (r = r0; r < r0 + qr; ++r)
                for (c = c0; c < c0 + qc; ++c) {
                    if (idx_out) {
                        uint32_t i00 = r * n + c;
                        uint32_t* o = idx_out + nidx;
                        o[0] = i00 + n; o[1] = i00; o[2] = i00 + n + 1u; o[3] = i00 + n + 1u; o[4] = i00; o[5] = i00 + 1u;
                    }
                    nidx += 6u;
                }
        }
    }
    if (counts_out) {
        counts_out[0] = nvis;
        counts_out[1] = nidx;
    }
    return MR_OK;
}
