"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle, bit for bit.

Bars (north-star): index buffers, triangulation output and positions bit-exact; normals within
2 ulp (they are in fact bit-exact: every operation is a single IEEE operation on both sides).
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _torch():
    import torch

    return torch


def _ulp_diff(a: np.ndarray, b: np.ndarray) -> int:
    """max distance in units in the last place between two f32 arrays (finite values)."""
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return int(np.abs(ia - ib).max()) if ia.size else 0


def _check_terrain(ctx, oracle, h, n, layout_fields=None, order="decl"):
    import myrenderer_b200 as mr

    lay = mr.VertexLayout.create(layout_fields or mr.TerrainVertex, order)
    mesh = mr.Terrain(ctx, lay).create_terrain(h)
    ctx.sync()
    olay = (lay.stride, lay.attributes)
    ovtx, oidx = oracle.terrain_build(h, n, layout=olay, nthreads=0)
    gv = mesh.vertex_buffer.vertex_buffer.cpu().numpy()
    gi = mesh.index_buffer.cpu().numpy().view(np.uint32)[: len(oidx)]
    assert np.array_equal(gi, oidx), "index buffer differs"
    V = gv.reshape(n * n, lay.stride)
    Ov = ovtx.reshape(n * n, lay.stride)
    po = lay.attributes[0][0]
    assert np.array_equal(V[:, po:po + 12], Ov[:, po:po + 12]), "positions differ"
    if len(lay.attributes) > 1:
        no = lay.attributes[1][0]
        d = _ulp_diff(np.ascontiguousarray(V[:, no:no + 12]).view(np.float32),
                      np.ascontiguousarray(Ov[:, no:no + 12]).view(np.float32))
        assert d <= 2, f"normals differ by {d} ulp (tolerance 2 ulp)"
    assert np.array_equal(gv, ovtx), "vertex bytes (incl. padding) differ"
    return mesh


def test_terrain_reference_heightmap(ctx, oracle):
    """Config 1: the reference's own 100x100 HEIGHTMAP.png."""
    h = np.load(os.path.join(GOLDEN, "heightmap_100.npy"))
    mesh = _check_terrain(ctx, oracle, h, 100)
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["terrain_100"]
    gv = mesh.vertex_buffer.vertex_buffer.cpu().numpy()
    gi = mesh.index_buffer.cpu().numpy().view(np.uint32)
    assert hashlib.sha256(gv.tobytes()).hexdigest() == kat["vtx_sha256"]
    assert hashlib.sha256(gi.tobytes()).hexdigest() == kat["idx_sha256"]
    assert mesh.index_count == 58806 and mesh.vertex_buffer.vertex_count == 10000
    assert mesh.bounding_box_p0 == (-10.0, 0.0, -10.0) and mesh.bounding_box_p1 == (10.0, 5.0, 10.0)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 255, 256, 257, 300, 1000])
def test_terrain_sizes_u16(ctx, oracle, n):
    h = oracle.synth_heightmap_u16(0x5EED0001, n)
    _check_terrain(ctx, oracle, h, n)


def test_terrain_f32_heights_and_extremes(ctx, oracle):
    n = 130
    h16 = oracle.synth_heightmap_u16(7, n)
    h16[0, :] = 0
    h16[1, :] = 65535
    h16[:, 5] = 65535
    _check_terrain(ctx, oracle, h16, n)
    _check_terrain(ctx, oracle, oracle.heightmap_normalize(h16), n)


def test_terrain_layout_variants(ctx, oracle):
    import myrenderer_b200 as mr

    n = 70
    h = oracle.synth_heightmap_u16(11, n)
    _check_terrain(ctx, oracle, h, n, (("normal", "Vec3"), ("pos", "Vec3")))  # fast path, swapped slots
    # generic path: position only (Vec4 slot), and a 48-byte vertex
    lay = mr.VertexLayout(16, ((0, 3),))
    mesh = mr.Terrain(ctx, lay).create_terrain(h)
    ctx.sync()
    ovtx, _ = oracle.terrain_build(h, n, layout=(16, ((0, 3),)))
    assert np.array_equal(mesh.vertex_buffer.vertex_buffer.cpu().numpy(), ovtx)
    lay = mr.VertexLayout(48, ((4, 3), (32, 3)))
    mesh = mr.Terrain(ctx, lay).create_terrain(h)
    ctx.sync()
    ovtx, _ = oracle.terrain_build(h, n, layout=(48, ((4, 3), (32, 3))))
    assert np.array_equal(mesh.vertex_buffer.vertex_buffer.cpu().numpy(), ovtx)


def test_terrain_positions_match_shader_stream(ctx, oracle):
    """Expanding the indexed mesh reproduces the WGSL vertex stream (Terrain.zig:24-48) bit for bit
    for every quad with r < n-1 and c < n-1 (SURVEY 8-a2')."""
    import myrenderer_b200 as mr

    n = 37
    h16 = oracle.synth_heightmap_u16(3, n)
    hf = oracle.heightmap_normalize(h16)
    mesh = mr.Terrain(ctx).create_terrain(h16)
    ctx.sync()
    V = mesh.vertex_buffer.vertex_buffer.cpu().numpy().reshape(n * n, 32)[:, :12].copy().view(np.float32)
    idx = mesh.index_buffer.cpu().numpy().view(np.uint32)
    k = 0
    for r in range(n - 1):
        for c in range(n - 1):
            quad = r * n + c
            for corner in range(6):
                want = oracle.terrain_shader_vertex(hf, n, quad * 6 + corner)
                got = V[idx[k]]
                assert want is not None and np.array_equal(got.view(np.uint32), want[:3].view(np.uint32))
                k += 1
    assert k == len(idx)


def test_terrain_row_band_sharding_matches_unsharded(ctx, oracle):
    """Logical ranks on one GPU: bands with a 1-row halo and band-local height buffers reproduce the
    unsharded mesh byte for byte."""
    import myrenderer_b200 as mr

    torch = _torch()
    n, G = 203, 4
    h = oracle.synth_heightmap_u16(0x5EED0004, n)
    ovtx, oidx = oracle.terrain_build(h, n)
    T = mr.Terrain(ctx)
    vtx = torch.zeros(n * n * 32, dtype=torch.uint8, device="cuda")
    idx = torch.zeros(6 * (n - 1) * (n - 1), dtype=torch.int32, device="cuda")
    rows = (C.c_uint32 * (G + 1))()
    qrows = (C.c_uint32 * (G + 1))()
    assert ctx.lib.mr_terrain_partition(n, G, rows, qrows) == 0
    for g in range(G):
        r0, r1 = rows[g], rows[g + 1]
        lo, hi = max(r0 - 1, 0), min(r1 + 1, n)
        band = torch.from_numpy(h[lo:hi].copy().view(np.int16)).cuda()  # band + halo only
        T.build(T.job(band, n, rows=(r0, r1), qrows=(qrows[g], qrows[g + 1]), height_row0=lo,
                      height_rows=hi - lo, vtx_out=vtx, vtx_row0=0, idx_out=idx, idx_qrow0=0))
    ctx.sync()
    assert np.array_equal(vtx.cpu().numpy(), ovtx)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), oidx)
    # band-local output buffers (vtx_row0 = row_begin)
    g = 2
    r0, r1 = rows[g], rows[g + 1]
    local = torch.zeros((r1 - r0) * n * 32, dtype=torch.uint8, device="cuda")
    lidx = torch.zeros((qrows[g + 1] - qrows[g]) * 6 * (n - 1), dtype=torch.int32, device="cuda")
    hd = torch.from_numpy(h.view(np.int16)).cuda()
    T.build(T.job(hd, n, rows=(r0, r1), qrows=(qrows[g], qrows[g + 1]), vtx_out=local, vtx_row0=r0,
                  idx_out=lidx, idx_qrow0=qrows[g]))
    ctx.sync()
    assert np.array_equal(local.cpu().numpy(), ovtx[r0 * n * 32: r1 * n * 32])
    assert np.array_equal(lidx.cpu().numpy().view(np.uint32), oidx[qrows[g] * 6 * (n - 1): qrows[g + 1] * 6 * (n - 1)])


def test_terrain_host_pointers(ctx, oracle):
    """The reference-facing call: host buffers in, host buffers out (copies inside the call)."""
    import myrenderer_b200 as mr

    n = 90
    h = oracle.synth_heightmap_u16(5, n)
    vtx = np.zeros(n * n * 32, dtype=np.uint8)
    idx = np.zeros(6 * (n - 1) * (n - 1), dtype=np.uint32)
    T = mr.Terrain(ctx)
    T.build(T.job(h, n, vtx_out=vtx, idx_out=idx))
    ovtx, oidx = oracle.terrain_build(h, n)
    assert np.array_equal(vtx, ovtx) and np.array_equal(idx, oidx)


def test_terrain_full_size_4096(ctx, oracle):
    """Config 2 at full size, compared with the (multi-threaded) oracle byte for byte."""
    n = 4096
    h = oracle.synth_heightmap_u16(0x5EED0001, n)
    _check_terrain(ctx, oracle, h, n)


def test_heightmap_normalize(ctx, oracle):
    torch = _torch()
    v = np.arange(65536, dtype=np.uint16)
    out = torch.empty(65536, dtype=torch.float32, device="cuda")
    d = torch.from_numpy(v.view(np.int16)).cuda()
    ctx.check(ctx.lib.mr_heightmap_normalize(ctx.handle, d.data_ptr(), 65536, out.data_ptr()), "normalize")
    ctx.sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), oracle.heightmap_normalize(v).view(np.uint32))


def test_constant_divisor_quotients_are_exact(ctx):
    """The vertex kernel divides by grid_step and 2*grid_step with q=a*y; r=fma(-q,b,a); q'=fma(r,y,q).
    Exhaustive: all 2^32 dividends against IEEE division, for the divisors the library enables the
    scheme for (and, forced, for a few others to show the guard range is what matters)."""
    for b, force in ((0.2, 0), (0.4, 0), (0.2, 1), (0.4, 1), (65535.0, 1), (0.3, 0)):
        bad = C.c_uint64(123)
        ctx.check(ctx.lib.mr_selftest_fastdiv(ctx.handle, b, force, C.byref(bad)), "selftest")
        assert bad.value == 0, f"divisor {b}: {bad.value} dividends differ from IEEE division"


# ---- polygons ---------------------------------------------------------------------------------
def _check_batch(ctx, oracle, xy, fp, *, offset_prime=None, seed=0, poly_index0=0, order="decl"):
    import myrenderer_b200 as mr

    lay = mr.VertexLayout.create(mr.GPUVertex, order)
    b = mr.Polygon(ctx, lay).create_polygons(xy, fp, offset_prime=offset_prime, seed=seed,
                                             poly_index0=poly_index0)
    ctx.sync()
    ref = oracle.polygon_batch(xy, fp, offset_prime=offset_prime, seed=seed, poly_index0=poly_index0,
                               layout=(lay.stride, lay.attributes), nthreads=0)
    gs = b.status.cpu().numpy().view(np.uint32)
    bad = np.where(gs != ref["status"])[0]
    assert bad.size == 0, f"status differs at {bad[:5]}: gpu {gs[bad[:5]]} oracle {ref['status'][bad[:5]]}"
    assert np.array_equal(b.ntri.cpu().numpy().view(np.uint32), ref["ntri"])
    gv = b.vertex_buffer.cpu().numpy()
    if not np.array_equal(gv, ref["vtx"]):
        per = lay.stride * 3
        ft = ref["first_tri"]
        for i in range(len(fp) - 1):
            a, z = int(ft[i]) * per, int(ft[i + 1]) * per
            assert np.array_equal(gv[a:z], ref["vtx"][a:z]), f"vertices of polygon {i} (n={int(fp[i+1]-fp[i])}) differ"
    gb = b.bbox.cpu().numpy().view(np.uint32)
    assert np.array_equal(gb, ref["bbox"].view(np.uint32)), "bbox differs"
    return b, ref


def test_app_polygons_all_orders(ctx, oracle):
    """Config 1: the two literal polygons of App.zig:68-83 for every (offset, prime), against the
    committed known answers."""
    app = json.load(open(os.path.join(GOLDEN, "app_polygons.json")))
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    for name in ("polygon1", "polygon2"):
        p = np.array(app[name], dtype=np.float32)
        keys = sorted(kat[name].keys())
        ops = np.array([[int(x) for x in k.split(",")] for k in keys], dtype=np.uint32)
        xy = np.tile(p, (len(keys), 1))
        fp = np.arange(len(keys) + 1, dtype=np.uint64) * len(p)
        b, ref = _check_batch(ctx, oracle, xy, fp, offset_prime=ops)
        gv = b.vertex_buffer.cpu().numpy()
        per = 96 * (len(p) - 2)
        for i, k in enumerate(keys):
            assert hashlib.sha256(gv[i * per:(i + 1) * per].tobytes()).hexdigest() == kat[name][k]["vtx_sha256"]
            assert int(b.status.cpu().numpy()[i]) == kat[name][k]["status"]


def test_known_answer_square_linear_order(ctx):
    """SURVEY 8-a hand-derived: polygon2 with (offset 0, prime 1) emits (2,3,1),(3,0,1)."""
    import myrenderer_b200 as mr

    sq = np.array([[10, 10], [40, 10], [40, 40], [10, 40]], dtype=np.float32)
    got = []
    st = mr.Triangulation(ctx).create_polygon(sq, got, lambda c, p: c.append(p), offset_prime=(0, 1))
    assert st == 0
    assert got == [(40.0, 40.0), (10.0, 40.0), (40.0, 10.0), (10.0, 40.0), (10.0, 10.0), (40.0, 10.0)]
    obj = mr.Polygon(ctx).create_polygon(sq, offset_prime=(0, 1))
    raw = obj.vertex_buffer.vertex_buffer.cpu().numpy().reshape(6, 32)
    col = raw[:, 16:28].copy().view(np.uint32).reshape(6, 3)
    assert col[0].tolist() == [0x3EB6B6B7, 0x3E44C4C5, 0x3EBCBCBD]  # palette[0], Polygon.zig:67
    assert col[3].tolist() == [0x3EE0E0E1, 0x3F800000, 0x3F4FCFD0]  # palette[1]
    assert obj.bounding_box_p0 == (0.0, 0.0, 0.0) and obj.bounding_box_p1 == (40.0, 40.0, 0.0)


def test_star_polygons_seeded(ctx, oracle):
    """Config 3 shape at reduced count: star polygons n in [8,64], device-seeded unirand.  Covers every
    status the reference algorithm produces on such input (OK, overflow, underfill, null unwrap)."""
    seed = 0x5EED0003
    fp = oracle.synth_polygon_sizes(seed, 4000, 8, 64)
    xy = oracle.synth_polygons(seed, fp)
    _, ref = _check_batch(ctx, oracle, xy, fp, seed=seed)
    seen = set(np.unique(ref["status"]).tolist())
    assert 0 in seen and any(s & 4 for s in seen) and any(s & 8 for s in seen)


def test_fast_path_is_the_one_that_runs(ctx, oracle):
    """On the benchmark's batch nothing should need the retry tier or the general (global-memory) path:
    the shared-memory fast path handles every polygon, including the ones whose search explodes."""
    import ctypes as C

    import myrenderer_b200 as mr

    seed = 0x5EED0003
    fp = oracle.synth_polygon_sizes(seed, 20000, 8, 64)
    xy = oracle.synth_polygons(seed, fp)
    mr.Polygon(ctx).create_polygons(xy, fp, seed=seed)
    tc = (C.c_uint32 * 8)()
    ctx.check(ctx.lib.mr_triangulate_tier_counts(ctx.handle, tc), "tier counts")
    assert tc[6] == 0 and tc[7] == 0, list(tc)      # general path, >1024 tier
    assert sum(tc[0:6]) <= 20000 // 100, list(tc)    # retry tier: at most a percent
    # coincident points force the general path, and it is counted
    sq = np.array([[0, 0], [10, 0], [10, 0], [10, 10], [0, 10]], dtype=np.float32)
    mr.Polygon(ctx).create_polygons(sq, np.array([0, 5], dtype=np.uint64), seed=1)
    ctx.check(ctx.lib.mr_triangulate_tier_counts(ctx.handle, tc), "tier counts")
    assert tc[6] == 1


def test_zigauto_layout_and_index_offset(ctx, oracle):
    seed = 99
    fp = oracle.synth_polygon_sizes(seed, 300, 3, 40)
    xy = oracle.synth_polygons(seed, fp)
    _check_batch(ctx, oracle, xy, fp, seed=seed, poly_index0=12345, order="zigauto")


def test_generic_vertex_layout_and_odd_unirand_pairs(ctx, oracle):
    """A vertex layout that is not the 32-byte fast case (stride 40, x at 4, colour at 20), and explicit
    (offset, prime) pairs outside what unirand_seed produces: offset >= n, prime >= n, prime == 0, and
    values whose product wraps u32 (unirand.zig:16 is u32 arithmetic) -- same formula in oracle and kernel."""
    import myrenderer_b200 as mr

    rng = np.random.default_rng(11)
    sizes = rng.integers(3, 40, 60)
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    pts = []
    for n in sizes:
        th = 2 * np.pi * (np.arange(n) + 0.8 * rng.random(n) - 0.4) / n
        pts.append(np.stack([100 + 60 * np.cos(th), 100 + 45 * np.sin(th)], 1))
    xy = np.concatenate(pts).astype(np.float32)
    op = np.zeros((60, 2), dtype=np.uint32)
    op[:, 0] = rng.integers(0, 100, 60)
    op[:, 1] = rng.integers(0, 100, 60)
    op[0] = (0, 0)
    op[1] = (5, 0)
    op[2] = (0xFFFFFFF0, 0xFFFFFFF1)
    op[3] = (123456789, 987654321)
    lay = mr.VertexLayout(40, ((4, 2), (20, 3)))
    b = mr.Polygon(ctx, lay).create_polygons(xy, fp, offset_prime=op)
    ctx.sync()
    ref = oracle.polygon_batch(xy, fp, offset_prime=op, layout=(40, ((4, 2), (20, 3))), nthreads=0)
    assert np.array_equal(b.status.cpu().numpy().view(np.uint32), ref["status"])
    assert np.array_equal(b.vertex_buffer.cpu().numpy(), ref["vtx"])
    assert np.array_equal(b.bbox.cpu().numpy().view(np.uint32), ref["bbox"].view(np.uint32))


def test_explicit_orders_that_repeat_edges_every_size_class(ctx, oracle):
    """An explicit (offset, prime) whose prime shares a factor with n visits some edges several times and others never
    (unirand.zig:16 is a plain multiply-add-modulo); the reference then runs add_segment on an edge that is already in
    the DAG.  In the conflict-list classes (n > 64) the edge's list was consumed by the first insertion and the second
    search found nothing (scripts/fuzz_parity.py found it); such orders are now recognised up front (gcd(prime, n) > 1,
    or values >= n whose u32 product wraps) and take the literal search of the next tier.  Pinned here for every size
    class, the team classes and the 3072-point class, with prime 0 (the same edge n times) and offset >= n among the pairs."""
    sizes = np.array([66, 65, 64, 100, 128, 130, 168, 200, 216, 260, 288, 300, 368, 400, 504, 600, 608, 700, 768, 900, 1024, 1500, 3072, 48, 12])
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    rng = np.random.default_rng(77)
    for fam in (oracle.FAMILY_ELLIPSE, oracle.FAMILY_ZIPPER, oracle.FAMILY_STAR):
        xy = oracle.synth_polygons(0xED6E + fam, fp, family=fam)
        for primes in ((2, 3, 4, 6), (0, 1, 5, 10), (7, 12, 25, 64)):
            op = np.zeros((len(sizes), 2), dtype=np.uint32)
            op[:, 0] = rng.integers(0, 5000, len(sizes))
            op[:, 1] = rng.choice(np.array(primes, dtype=np.uint32), len(sizes))
            _check_batch(ctx, oracle, xy, fp, offset_prime=op)


def test_convex_and_large_polygons(ctx, oracle):
    """Sizes up to 1024 (every shared-memory class) and 1025..4096 (global-memory tier)."""
    rng = np.random.default_rng(5)
    sizes = [3, 4, 5, 16, 17, 32, 33, 64, 65, 128, 129, 256, 257, 512, 513, 1024, 1025, 2000, 4096]
    pts = []
    for n in sizes:
        th = 2 * np.pi * (np.arange(n) + 0.8 * rng.random(n) - 0.4) / n
        a, b, ph = 40 + 50 * rng.random(), 40 + 50 * rng.random(), rng.random() * 6.28
        x, y = a * np.cos(th), b * np.sin(th)
        pts.append(np.stack([100 + np.cos(ph) * x - np.sin(ph) * y, 100 + np.sin(ph) * x + np.cos(ph) * y], 1))
    xy = np.concatenate(pts).astype(np.float32)
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    _, ref = _check_batch(ctx, oracle, xy, fp, seed=1)
    assert (ref["status"] == 0).all(), "the reference algorithm handles convex polygons"


def _convex(n, rng):
    th = 2 * np.pi * (np.arange(n) + 0.8 * rng.random(n) - 0.4) / n
    a, b, ph = 40 + 50 * rng.random(), 40 + 50 * rng.random(), rng.random() * 6.28
    x, y = a * np.cos(th), b * np.sin(th)
    return np.stack([100 + np.cos(ph) * x - np.sin(ph) * y, 100 + np.sin(ph) * x + np.cos(ph) * y], 1)


def test_size_class_boundaries_convex_and_star(ctx, oracle):
    """Every size class of the kernel (n <= 64, 128, 168, 216, 288, 368, 504, 608, 768, 1024, then the one-per-SM class
    up to 3072) at its largest size and one past it; several polygons per size so that the persistent warps / warp teams run
    their queue loop, and explicit unirand pairs as well as seeded ones."""
    import myrenderer_b200 as mr

    rng = np.random.default_rng(11)
    sizes = []
    for top in (64, 128, 168, 216, 288, 368, 504, 608, 768, 1024):
        sizes += [top, top + 1] * 3
    xy = np.concatenate([_convex(n, rng) for n in sizes]).astype(np.float32)
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    _, ref = _check_batch(ctx, oracle, xy, fp, seed=21)
    assert (ref["status"] == 0).all()
    tc = (C.c_uint32 * 8)()
    ctx.check(ctx.lib.mr_triangulate_tier_counts(ctx.handle, tc), "tier counts")
    assert sum(tc[0:7]) == 0, list(tc)            # convex input: everything in the first shared-memory pass
    assert tc[7] == 0                              # the three 1025-gons run in the 1025..3072 shared-memory class
    # explicit (offset, prime) pairs, incl. prime 1 and the largest table prime below n
    ops = np.array([[1 + (i * 7) % (n - 1), 1 if i % 2 else 3] for i, n in enumerate(sizes)], dtype=np.uint32)
    _check_batch(ctx, oracle, xy, fp, offset_prime=ops)
    # star-shaped polygons of the same sizes: the reference algorithm fails on many of them (overflow, underfill,
    # null unwrap, arena caps); status and bytes must still agree
    sxy = oracle.synth_polygons(77, fp)
    _check_batch(ctx, oracle, sxy, fp, seed=77)


def test_skewed_sizes_loguniform(ctx, oracle):
    """Config 5 shape at reduced count: log-uniform sizes 8..1024, star-shaped (mostly failing in the
    reference algorithm -> exercises arena caps and re-queueing to the global-memory tier)."""
    seed = 0x5EED0005
    fp = oracle.synth_polygon_sizes(seed, 400, 8, 1024, dist=1)
    xy = oracle.synth_polygons(seed, fp)
    _check_batch(ctx, oracle, xy, fp, seed=seed)


def test_edge_cases(ctx, oracle):
    polys = [
        np.zeros((0, 2)),                                     # empty
        np.array([[1, 1]]),                                   # n = 1
        np.array([[1, 1], [2, 2]]),                           # n = 2
        np.array([[0, 0], [10, 0], [0, 10]]),                 # triangle
        np.array([[0, 0], [10, 0], [np.nan, 10], [0, 5]]),    # NaN
        np.array([[0, 0], [np.inf, 0], [5, 10], [0, 5]]),     # Inf
        np.array([[0, 0], [5, 0], [10, 0], [10, 10], [0, 10]]),       # collinear run on a horizontal edge
        np.array([[0, 0], [10, 0], [10, 0], [10, 10], [0, 10]]),      # coincident consecutive points
        np.array([[5, 5], [5, 5], [5, 5], [5, 5]]),                   # all coincident
        np.array([[0, 0], [10, 10], [10, 0], [0, 10]]),               # self-intersecting (bow tie)
        np.array([[10, 10], [10, 40], [40, 40], [40, 10]]),           # square, opposite orientation
        np.array([[0, 0], [2e10, 5e-41], [1e10, 1e-40]]),             # atan2 corner: pi vs 0 (not acute)
        np.array([[0, 0], [3e38, 0], [3e38, 3e38], [-3e38, 3e38]]),   # differences overflow to inf
    ]
    xy = np.concatenate([p.reshape(-1, 2) for p in polys]).astype(np.float32)
    fp = np.concatenate([[0], np.cumsum([len(p) for p in polys])]).astype(np.uint64)
    for seed in (1, 2, 3):
        _, ref = _check_batch(ctx, oracle, xy, fp, seed=seed)
    assert ref["status"][0] == 1 and ref["status"][1] == 1 and ref["status"][2] == 1
    assert ref["status"][4] == 2 and ref["status"][5] == 2


def test_acute_test_near_the_negative_x_axis(ctx, oracle):
    """push_triangle_if_acute (Triangulation.zig:398-425) compares two atan2 values with pi; the fast path only evaluates
    atan2 when a difference vector comes within 1e-6 rad of the negative x axis.  Thin triangles whose lowest vertex sees
    a neighbour at slopes 0, denormal, 1e-9 .. 1e-4 exercise both sides of that shortcut, at several magnitudes."""
    rng = np.random.default_rng(3)
    polys = []
    for scale in (1.0, 1e-3, 1e10, 3e-20):
        for slope in (0.0, 1e-45, 5e-41, 1e-38, 1e-12, 1e-9, 3e-8, 1e-7, 5e-7, 9e-7, 1.1e-6, 2e-6, 1e-5, 1e-4):
            for k in range(6):
                w = scale * (1.0 + rng.random())
                y_far = w * slope * (1.0 + 0.5 * rng.random())          # far point barely above/at the near point's level
                y_mid = 0.5 * y_far * rng.random()
                tri = np.array([[0.0, 0.0], [2.0 * w, y_far], [w, y_mid]])
                if k % 2:
                    tri = tri[::-1].copy()
                if k % 3 == 0:
                    tri[:, 0] = -tri[:, 0]
                polys.append(tri)
    # quadrilaterals with a nearly flat bottom chain as well
    for slope in (1e-41, 1e-9, 1e-7, 1e-6, 1e-5):
        for w in (1.0, 1e8):
            polys.append(np.array([[0.0, 0.0], [w, w * slope], [2 * w, 2.5 * w * slope], [w, -w]]))
            polys.append(np.array([[0.0, 0.0], [w, -w], [2 * w, 2.5 * w * slope], [w, w * slope]]))
    xy = np.concatenate(polys).astype(np.float32)
    fp = np.concatenate([[0], np.cumsum([len(q) for q in polys])]).astype(np.uint64)
    _, ref = _check_batch(ctx, oracle, xy, fp, seed=4)
    # every polygon must have taken a definite path: statuses agree (checked above) and the batch contains both
    # polygons that emit their full n-2 triangles and ones that do not
    assert (ref["ntri"] > 0).any()


def test_too_large_polygon(ctx, oracle):
    n = 4097
    th = 2 * np.pi * np.arange(n) / n
    xy = np.stack([100 + 50 * np.cos(th), 100 + 50 * np.sin(th)], 1).astype(np.float32)
    xy = np.concatenate([xy, np.array([[0, 0], [10, 0], [0, 10]], dtype=np.float32)])
    fp = np.array([0, n, n + 3], dtype=np.uint64)
    _, ref = _check_batch(ctx, oracle, xy, fp, seed=1)
    assert ref["status"][0] == 32 and ref["status"][1] == 0


def test_polygon_host_pointers_and_subrange(ctx, oracle):
    """Host buffers through the C ABI; and a sub-batch [a,b) addressed with point_base / tri_base
    (the per-rank call of the sharded path)."""
    import myrenderer_b200 as mr

    seed = 21
    fp = oracle.synth_polygon_sizes(seed, 200, 5, 50)
    xy = oracle.synth_polygons(seed, fp)
    ref = oracle.polygon_batch(xy, fp, seed=seed)
    ft = ref["first_tri"]
    P = mr.Polygon(ctx)
    vtx = np.zeros(int(ft[-1]) * 96, dtype=np.uint8)
    status = np.zeros(200, dtype=np.uint32)
    ntri = np.zeros(200, dtype=np.uint32)
    bbox = np.zeros((200, 4), dtype=np.float32)
    for a, z in ((0, 77), (77, 200)):  # two "ranks" writing into one buffer
        j = P.job(xy, fp[a:z + 1].copy(), z - a, vtx_out=vtx, first_tri=ft[a:z + 1].copy(), bbox_out=bbox[a:z],
                  status_out=status[a:z], ntri_out=ntri[a:z], seed=seed, poly_index0=a)
        P.triangulate(j)
    assert np.array_equal(status, ref["status"]) and np.array_equal(ntri, ref["ntri"])
    assert np.array_equal(vtx, ref["vtx"]) and np.array_equal(bbox.view(np.uint32), ref["bbox"].view(np.uint32))


def test_polygon_pinned_host_output_zero_copy(ctx, oracle):
    """Pinned host vertex buffers are written by the kernels directly (zero-copy over PCIe); pageable
    ones are staged.  Both must give the oracle's bytes."""
    import myrenderer_b200 as mr

    torch = _torch()
    seed = 33
    fp = oracle.synth_polygon_sizes(seed, 3000, 8, 64)
    xy = oracle.synth_polygons(seed, fp)
    ref = oracle.polygon_batch(xy, fp, seed=seed, nthreads=0)
    ft = ref["first_tri"]
    P = mr.Polygon(ctx)
    pinned = torch.full((int(ft[-1]) * 96,), 0xAB, dtype=torch.uint8).pin_memory()
    status = np.zeros(3000, dtype=np.uint32)
    P.triangulate(P.job(xy, fp, 3000, vtx_out=pinned, first_tri=ft, status_out=status, seed=seed))
    assert np.array_equal(status, ref["status"])
    assert np.array_equal(pinned.numpy(), ref["vtx"])


def test_unirand_device_port(ctx, oracle):
    torch = _torch()
    tops = np.concatenate([np.arange(2, 300), [509, 521, 1013, 1024, 1723, 1724, 4096]])
    fp = np.concatenate([[0], np.cumsum(tops)]).astype(np.uint64)
    out = torch.empty(2 * len(tops), dtype=torch.int32, device="cuda")
    d = torch.from_numpy(fp.view(np.int64)).cuda()
    ctx.check(ctx.lib.mr_unirand_seed_batch(ctx.handle, d.data_ptr(), len(tops), 0xABCDEF, 7, out.data_ptr()), "seed")
    ctx.sync()
    got = out.cpu().numpy().view(np.uint32).reshape(-1, 2)
    for i, t in enumerate(tops):
        assert tuple(got[i]) == oracle.unirand_seed(int(t), 0xABCDEF, 7 + i)


def test_synth_generators_match_host_definition(ctx, oracle):
    torch = _torch()
    n = 257
    out = torch.empty(n * n, dtype=torch.int16, device="cuda")
    ctx.check(ctx.lib.mr_synth_heightmap_u16(ctx.handle, 0x5EED0001, n, 0, n, out.data_ptr()), "synth h")
    fp = oracle.synth_polygon_sizes(0x5EED0003, 3000, 8, 64)
    d = torch.from_numpy(fp.view(np.int64)).cuda()
    xy = torch.empty(int(fp[-1]) * 2, dtype=torch.float32, device="cuda")
    ctx.check(ctx.lib.mr_synth_polygons(ctx.handle, 0x5EED0003, 0, d.data_ptr(), 3000, xy.data_ptr()), "synth p")
    ctx.sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint16).reshape(n, n), oracle.synth_heightmap_u16(0x5EED0001, n))
    want = oracle.synth_polygons(0x5EED0003, fp)
    assert np.array_equal(xy.cpu().numpy().view(np.uint32), want.reshape(-1).view(np.uint32))


def test_polygon_offsets_device(ctx, oracle):
    torch = _torch()
    fp = oracle.synth_polygon_sizes(4, 5000, 1, 70)
    d = torch.from_numpy(fp.view(np.int64)).cuda()
    out = torch.empty(5001, dtype=torch.int64, device="cuda")
    ctx.check(ctx.lib.mr_polygon_offsets(ctx.handle, d.data_ptr(), 5000, out.data_ptr()), "offsets")
    ctx.sync()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), oracle.polygon_offsets(fp))


# ---- round 2 additions ---------------------------------------------------------------------------
def test_synth_families_match_host_definition(ctx, oracle):
    torch = _torch()
    fp = oracle.synth_polygon_sizes(0x5EED0005, 400, 3, 1024, dist=1)
    d = torch.from_numpy(fp.view(np.int64)).cuda()
    for fam in (oracle.FAMILY_STAR, oracle.FAMILY_ELLIPSE, oracle.FAMILY_ZIPPER):
        xy = torch.empty(int(fp[-1]) * 2, dtype=torch.float32, device="cuda")
        ctx.check(ctx.lib.mr_synth_polygons_family(ctx.handle, fam, 0xFA111, 17, d.data_ptr(), 400, xy.data_ptr()), "synth")
        ctx.sync()
        want = oracle.synth_polygons(0xFA111, fp, poly_index0=17, family=fam)
        assert np.array_equal(xy.cpu().numpy().view(np.uint32), want.reshape(-1).view(np.uint32)), fam


@pytest.mark.parametrize("family", ["ellipse", "zipper"])
def test_sound_families_every_size_class(ctx, oracle, family):
    """The two families the reference triangulates correctly for every edge order (see
    tests/test_reference_soundness_cpu.py), at and around every size-class boundary of the kernels, up to 1024
    points: bit-exact and all status OK.  The zipper family is the non-convex one."""
    fam = {"ellipse": oracle.FAMILY_ELLIPSE, "zipper": oracle.FAMILY_ZIPPER}[family]
    sizes = []
    for b in (64, 128, 168, 216, 288, 368, 504, 608, 768, 1024):
        sizes += [b - 1, b, b + 1] if b < 1024 else [b - 1, b]
    sizes = np.array(sizes * 6 + [3, 4, 5, 7, 8, 9, 31, 32, 33] * 4)
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    xy = oracle.synth_polygons(0x21BB, fp, family=fam)
    _, ref = _check_batch(ctx, oracle, xy, fp, seed=0x21BB)
    assert (ref["status"] == 0).all()
    tc = (C.c_uint32 * 8)()
    ctx.check(ctx.lib.mr_triangulate_tier_counts(ctx.handle, tc), "tier counts")
    assert tc[6] == 0 and sum(tc[0:6]) == 0, list(tc)  # the first shared-memory pass handles all of them


def test_zipper_loguniform_batch(ctx, oracle):
    seed = 0x5EED0005
    fp = oracle.synth_polygon_sizes(seed, 6000, 8, 1024, dist=1)
    xy = oracle.synth_polygons(seed, fp, family=oracle.FAMILY_ZIPPER)
    _, ref = _check_batch(ctx, oracle, xy, fp, seed=seed)
    assert (ref["status"] == 0).all()


def test_small_batch_path_single_polygon(ctx, oracle):
    """Polygon.create_polygon's own call shape (Polygon.zig:81-107, App.zig:68-83): ONE polygon, every buffer in
    host memory.  One kernel launch, results equal to the oracle, for every edge order of both App polygons."""
    import myrenderer_b200 as mr

    app = json.load(open(os.path.join(GOLDEN, "app_polygons.json")))
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    P = mr.Polygon(ctx)
    for name in ("polygon1", "polygon2"):
        p = np.array(app[name], dtype=np.float32)
        n = len(p)
        fp = np.array([0, n], dtype=np.uint64)
        ft = np.array([0, n - 2], dtype=np.uint64)
        for key, want in kat[name].items():
            op = np.array([int(x) for x in key.split(",")], dtype=np.uint32)
            vtx = np.full((n - 2) * 96, 0xCD, dtype=np.uint8)
            bbox = np.zeros(4, dtype=np.float32)
            st = np.full(1, 77, dtype=np.uint32)
            nt = np.zeros(1, dtype=np.uint32)
            l0 = ctx.launch_count
            P.triangulate(P.job(p, fp, 1, vtx_out=vtx, first_tri=ft, bbox_out=bbox, status_out=st, ntri_out=nt, offset_prime=op))
            assert ctx.launch_count - l0 == 1, "the single-polygon call must be one kernel launch"
            assert int(st[0]) == want["status"] and int(nt[0]) == n - 2
            assert hashlib.sha256(vtx.tobytes()).hexdigest() == want["vtx_sha256"]
            assert bbox.view(np.uint32).tolist() == want["bbox_bits"]


def test_small_batch_path_mixed(ctx, oracle):
    """Small host-memory batches through the one-block path: every size class incl. > 1024 points, degenerate and
    too-large polygons, exploding stars (which need the retry tiers), seeded and explicit edge orders, optional
    outputs left out, sub-ranges with point_base / tri_base."""
    import myrenderer_b200 as mr

    P = mr.Polygon(ctx)
    seed = 0xBEEF
    sizes = np.array([7, 4, 2, 3, 64, 65, 130, 300, 520, 700, 1024, 1500, 36, 36, 5000, 12, 0, 1, 900, 450])
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    for fam in (oracle.FAMILY_STAR, oracle.FAMILY_ZIPPER):
        xy = oracle.synth_polygons(seed, fp, family=fam)
        ref = oracle.polygon_batch(xy, fp, seed=seed, nthreads=0)
        ft = ref["first_tri"]
        npoly = len(sizes)
        vtx = np.full(max(int(ft[-1]) * 96, 1), 0xEE, dtype=np.uint8)
        st = np.zeros(npoly, dtype=np.uint32)
        nt = np.zeros(npoly, dtype=np.uint32)
        bbox = np.zeros((npoly, 4), dtype=np.float32)
        P.triangulate(P.job(xy, fp, npoly, vtx_out=vtx, first_tri=ft, bbox_out=bbox, status_out=st, ntri_out=nt, seed=seed))
        assert np.array_equal(st, ref["status"]), (st, ref["status"])
        assert np.array_equal(nt, ref["ntri"]) and np.array_equal(vtx[: int(ft[-1]) * 96], ref["vtx"])
        assert np.array_equal(bbox.view(np.uint32), ref["bbox"].view(np.uint32))
        # sub-range [5, 12) into the same buffers, only the vertex output requested
        vtx2 = np.full_like(vtx, 0x11)
        P.triangulate(P.job(xy, fp[5:13].copy(), 7, vtx_out=vtx2, first_tri=ft[5:13].copy(), seed=seed, poly_index0=5))
        a, z = int(ft[5]) * 96, int(ft[12]) * 96
        assert np.array_equal(vtx2[a:z], ref["vtx"][a:z])
        assert (vtx2[:a] == 0x11).all() and (vtx2[z:] == 0x11).all()  # nothing outside the job's range is touched


def test_pinned_vertex_output_without_other_host_outputs(ctx, oracle):
    """a pinned host vtx_out is written by the kernels directly; the call must not return before they
    are done even when no other host output forces a copy-back."""
    import myrenderer_b200 as mr

    torch = _torch()
    seed = 34
    fp = oracle.synth_polygon_sizes(seed, 20000, 8, 64)
    xy = oracle.synth_polygons(seed, fp)
    ref = oracle.polygon_batch(xy, fp, seed=seed, nthreads=0, want_ids=False)
    ft = ref["first_tri"]
    P = mr.Polygon(ctx)
    xy_d = torch.from_numpy(xy).cuda()
    fp_d = torch.from_numpy(fp.view(np.int64)).cuda()
    ft_d = torch.from_numpy(ft.view(np.int64)).cuda()
    st_d = torch.empty(20000, dtype=torch.int32, device="cuda")
    for _ in range(3):
        pinned = torch.full((int(ft[-1]) * 96,), 0xAB, dtype=torch.uint8).pin_memory()
        P.triangulate(P.job(xy_d, fp_d, 20000, vtx_out=pinned, first_tri=ft_d, status_out=st_d, seed=seed))
        got = pinned.numpy().copy()  # no ctx.sync(): the call itself must have waited
        assert np.array_equal(got, ref["vtx"])


def test_argument_validation_alignment_and_empty_bands(ctx, oracle):
    import myrenderer_b200 as mr

    torch = _torch()
    n = 64
    h = torch.from_numpy(oracle.synth_heightmap_u16(1, n).view(np.int16)).cuda()
    T = mr.Terrain(ctx)
    vtx = torch.empty(n * n * 32 + 64, dtype=torch.uint8, device="cuda")
    idx = torch.empty(6 * (n - 1) * (n - 1) + 16, dtype=torch.int32, device="cuda")
    with pytest.raises(mr.MrError):  # idx_out off by one u32: would fault in the 8-byte stores
        T.build(T.job(h, n, vtx_out=vtx, idx_out=idx.data_ptr() + 4))
    with pytest.raises(mr.MrError):
        T.build(T.job(h, n, vtx_out=vtx.data_ptr() + 2, idx_out=idx))
    with pytest.raises(mr.MrError):  # vtx_row0 beyond the (empty) band: used to wrap into a huge allocation
        T.build(T.job(h, n, rows=(3, 3), vtx_out=np.zeros(16, dtype=np.uint8), vtx_row0=9))
    T.build(T.job(h, n, rows=(5, 5), qrows=(2, 2), vtx_out=vtx, idx_out=idx))  # empty band: a no-op
    T.build(T.job(h, n, vtx_out=vtx, idx_out=idx))
    ctx.sync()  # the context is still healthy
    ovtx, oidx = oracle.terrain_build(h.cpu().numpy().view(np.uint16).reshape(n, n), n)
    assert np.array_equal(vtx.cpu().numpy()[: n * n * 32], ovtx)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32)[: len(oidx)], oidx)


def test_context_trim_releases_scratch(ctx, oracle):
    n = 512
    h = oracle.synth_heightmap_u16(3, n)
    import myrenderer_b200 as mr

    T = mr.Terrain(ctx)
    vtx = np.zeros(n * n * 32, dtype=np.uint8)
    idx = np.zeros(6 * (n - 1) * (n - 1), dtype=np.uint32)
    T.build(T.job(h, n, vtx_out=vtx, idx_out=idx))
    b = C.c_uint64()
    ctx.check(ctx.lib.mr_context_scratch_bytes(ctx.handle, C.byref(b)), "scratch bytes")
    assert b.value >= n * n * 32
    ctx.check(ctx.lib.mr_context_trim(ctx.handle), "trim")
    ctx.check(ctx.lib.mr_context_scratch_bytes(ctx.handle, C.byref(b)), "scratch bytes")
    assert b.value == 0
    vtx2 = np.zeros_like(vtx)
    T.build(T.job(h, n, vtx_out=vtx2, idx_out=idx))  # and it works again afterwards
    assert np.array_equal(vtx, vtx2)
    ovtx, oidx = oracle.terrain_build(h, n)
    assert np.array_equal(vtx, ovtx) and np.array_equal(idx, oidx)


def test_terrain_host_band_into_larger_buffer(ctx, oracle):
    """Host vtx_out / idx_out with vtx_row0 / idx_qrow0 below the band: only the band's bytes are written."""
    import myrenderer_b200 as mr

    n = 200
    h = oracle.synth_heightmap_u16(5, n)
    T = mr.Terrain(ctx)
    vtx = np.full(n * n * 32, 0x5A, dtype=np.uint8)
    idx = np.full(6 * (n - 1) * (n - 1), 0x5A5A5A5A, dtype=np.uint32)
    T.build(T.job(h, n, rows=(50, 120), qrows=(40, 90), vtx_out=vtx, vtx_row0=0, idx_out=idx, idx_qrow0=0))
    ovtx, oidx = oracle.terrain_build(h, n)
    a, z = 50 * n * 32, 120 * n * 32
    assert np.array_equal(vtx[a:z], ovtx[a:z]) and (vtx[:a] == 0x5A).all() and (vtx[z:] == 0x5A).all()
    a, z = 40 * 6 * (n - 1), 90 * 6 * (n - 1)
    assert np.array_equal(idx[a:z], oidx[a:z]) and (idx[:a] == 0x5A5A5A5A).all() and (idx[z:] == 0x5A5A5A5A).all()


def test_terrain_tiles_and_cull_match_oracle(ctx, oracle):
    """per-tile bounding boxes and the SceneNode.zig:96-110 visibility test on the GPU, with the
    visible tiles compacted (ascending) into one index buffer -- bit-exact against the oracle, device and host pointers."""
    from test_terrain_cull_cpu import camera_matrix, mat_image

    torch = _torch()
    lib = ctx.lib
    for n, tr, tc in ((130, 32, 24), (1000, 64, 64), (257, 8, 256), (300, 500, 7), (1024, 33, 16), (264, 64, 24)):  # the last three u16 cases take 128-bit loads
        h = oracle.synth_heightmap_u16(0x5EED0001, n)
        want_box = oracle.terrain_tile_bounds(h, n, tr, tc)
        ntiles = want_box.shape[0]
        hd = torch.from_numpy(h.view(np.int16)).cuda()
        box_d = torch.empty(ntiles * 8, dtype=torch.float32, device="cuda")
        ctx.check(lib.mr_terrain_tile_bounds(ctx.handle, hd.data_ptr(), 0, n, tr, tc, None, box_d.data_ptr()), "tile bounds")
        ctx.sync()
        assert np.array_equal(box_d.cpu().numpy().view(np.uint32), want_box.reshape(-1).view(np.uint32))
        box_h = np.zeros((ntiles, 8), dtype=np.float32)  # host pointers, f32 heights
        hf = oracle.heightmap_normalize(h)
        ctx.check(lib.mr_terrain_tile_bounds(ctx.handle, hf.ctypes.data, 1, n, tr, tc, None, box_h.ctypes.data), "tile bounds")
        assert np.array_equal(box_h.view(np.uint32), want_box.view(np.uint32))
        if n == 1000:  # one-column tiles (512 tiles per strip), zero and negative height scales
            for tr2, tc2, prm in ((16, 1, (0.2, 0.1, 5.0)), (7, 1, (0.2, 0.1, 0.0)), (64, 64, (0.2, 0.1, 0.0)), (33, 2, (0.3, -0.5, -4.0))):
                wb = oracle.terrain_tile_bounds(h, n, tr2, tc2, prm)
                bd = torch.empty(wb.size, dtype=torch.float32, device="cuda")
                pp = (C.c_float * 3)(*prm)
                ctx.check(lib.mr_terrain_tile_bounds(ctx.handle, hd.data_ptr(), 0, n, tr2, tc2, pp, bd.data_ptr()), "tile bounds")
                ctx.sync()
                assert np.array_equal(bd.cpu().numpy().view(np.uint32), wb.reshape(-1).view(np.uint32)), (tr2, tc2, prm)
                hneg = (oracle.heightmap_normalize(h) - np.float32(0.5)) * np.float32(0.0)  # a float map of +0 and -0
                hneg[::3] = np.float32(-0.0)
                wb = oracle.terrain_tile_bounds(hneg, n, tr2, tc2, prm)
                ctx.check(lib.mr_terrain_tile_bounds(ctx.handle, torch.from_numpy(hneg).cuda().data_ptr(), 1, n, tr2, tc2, pp, bd.data_ptr()), "tile bounds")
                ctx.sync()
                assert np.array_equal(bd.cpu().numpy().view(np.uint32), wb.reshape(-1).view(np.uint32)), ("zeros", tr2, tc2, prm)
        mats = [camera_matrix(), camera_matrix((5.0, 60.0, 5.0), (40.0, 0.0, 40.0)), mat_image(np.array([[0, 0, 0, 1]] * 4, dtype=np.float32)),
                mat_image(np.array([[0, 0, 0, -1], [0, 0, 0, 2], [0, 0, 0, 0], [0, 0, 0, 1]], dtype=np.float32))]
        for m in mats:
            ref = oracle.terrain_cull(want_box, n, tr, tc, m)
            vis = torch.empty(ntiles, dtype=torch.int32, device="cuda")
            ids = torch.empty(ntiles, dtype=torch.int32, device="cuda")
            idx = torch.full((6 * (n - 1) * (n - 1),), -1, dtype=torch.int32, device="cuda")
            cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
            ctx.check(lib.mr_terrain_cull(ctx.handle, box_d.data_ptr(), n, tr, tc, m.ctypes.data, vis.data_ptr(), ids.data_ptr(),
                                          idx.data_ptr(), cnt.data_ptr()), "cull")
            ctx.sync()
            c = cnt.cpu().numpy()
            assert c.tolist() == ref["counts"].tolist()
            assert np.array_equal(vis.cpu().numpy().view(np.uint32), ref["visible"])
            assert np.array_equal(ids.cpu().numpy().view(np.uint32)[: c[0]], ref["ids"])
            g = idx.cpu().numpy().view(np.uint32)
            assert np.array_equal(g[: c[1]], ref["idx"]) and (g[c[1]:] == 0xFFFFFFFF).all()
            # host pointers, ids not requested
            idx_h = np.zeros(6 * (n - 1) * (n - 1), dtype=np.uint32)
            cnt_h = np.zeros(2, dtype=np.uint64)
            ctx.check(lib.mr_terrain_cull(ctx.handle, want_box.ctypes.data, n, tr, tc, m.ctypes.data, None, None,
                                          idx_h.ctypes.data, cnt_h.ctypes.data), "cull host")
            assert cnt_h.tolist() == ref["counts"].tolist() and np.array_equal(idx_h[: int(cnt_h[1])], ref["idx"])


def test_xl_class_1025_to_3072_points(ctx, oracle):
    """Polygons of 1025..3072 points run in shared memory too (one polygon per SM, a team of 8 warps); above that, and
    whenever such a polygon outgrows the typical-case arenas (exploding stars), the global-memory general path takes
    over.  Bit-exact either way."""
    sizes = np.array([1025, 1026, 1500, 2047, 2048, 2049, 3000, 3071, 3072, 3073, 4096, 1024, 900])
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    tc = (C.c_uint32 * 8)()
    for fam in (oracle.FAMILY_ELLIPSE, oracle.FAMILY_ZIPPER):
        xy = oracle.synth_polygons(0x71A5, fp, family=fam)
        _, ref = _check_batch(ctx, oracle, xy, fp, seed=0x71A5)
        assert (ref["status"] == 0).all()
        ctx.check(ctx.lib.mr_triangulate_tier_counts(ctx.handle, tc), "tier counts")
        assert tc[5] == 0 and tc[6] == 0 and tc[7] == 2, list(tc)  # only 3073 and 4096 take the general path
    # stars of that size explode: handed over, same bytes as the oracle (mostly MR_POLY_ARENA)
    sizes = np.array([1100, 1300, 2000, 2500, 3072, 1025])
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    xy = oracle.synth_polygons(0x71A6, fp, family=oracle.FAMILY_STAR)
    _check_batch(ctx, oracle, xy, fp, seed=0x71A6)
    ctx.check(ctx.lib.mr_triangulate_tier_counts(ctx.handle, tc), "tier counts")
    assert tc[5] >= 1 and tc[7] == 0, list(tc)
    # host pointers: the small-batch path schedules the same classes
    import myrenderer_b200 as mr

    P = mr.Polygon(ctx)
    sizes = np.array([2048, 7, 3073, 1025])
    fp = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    xy = oracle.synth_polygons(0x71A7, fp, family=oracle.FAMILY_ZIPPER)
    ref = oracle.polygon_batch(xy, fp, seed=0x71A7, nthreads=0)
    ft = ref["first_tri"]
    vtx = np.zeros(int(ft[-1]) * 96, dtype=np.uint8)
    st = np.zeros(4, dtype=np.uint32)
    P.triangulate(P.job(xy, fp, 4, vtx_out=vtx, first_tri=ft, status_out=st, seed=0x71A7))
    assert np.array_equal(st, ref["status"]) and np.array_equal(vtx, ref["vtx"])
