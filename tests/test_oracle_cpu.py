"""CPU tests: pin the oracle (known answers, independent second restatement, invariants)."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _app():
    return json.load(open(os.path.join(GOLDEN, "app_polygons.json")))


def test_hand_derived_known_answer(oracle):
    """polygon2 with the linear order (offset 0, prime 1): 18 nodes, mountains in the order
    key(1,2), key(0,3); triangles (2,3,1) then (3,0,1).  Derived by hand from the Zig source."""
    sq = np.array(_app()["polygon2"], dtype=np.float32)
    r = oracle.polygon_batch(sq, np.array([0, 4]), offset_prime=[0, 1], want_stats=True)
    assert r["ids"].tolist() == [2, 3, 1, 3, 0, 1]
    assert r["stats"]["nodes"] == 18 and r["stats"]["mountains"] == 2 and r["stats"]["sum_stack"] == 5
    assert r["status"][0] == 0
    xy = r["vtx"].reshape(6, 32)[:, :8].copy().view(np.float32).reshape(6, 2)
    assert xy.tolist() == [[40, 40], [10, 40], [40, 10], [10, 40], [10, 10], [40, 10]]
    # bbox with the as-written rule of Polygon.zig:73-76, starting from (0,0),(0,0)
    assert r["bbox"][0].tolist() == [0.0, 0.0, 40.0, 40.0]


def test_palette_bits(oracle):
    """Polygon.zig:50-57,66-71: byte-reversed channel order, f32(u8)/255 (SURVEY 8-a12 bit patterns)."""
    import ctypes as C

    pal = (C.c_float * 12)()
    oracle.lib().mr_o_palette(pal)
    bits = np.array(pal[:], dtype=np.float32).view(np.uint32).reshape(4, 3).tolist()
    assert bits == [[0x3EB6B6B7, 0x3E44C4C5, 0x3EBCBCBD], [0x3EE0E0E1, 0x3F800000, 0x3F4FCFD0],
                    [0x3EE0E0E1, 0x3F23A3A4, 0x3E70F0F1], [0x3F2BABAC, 0x3EB6B6B7, 0x3E969697]]


def test_golden_app_polygons(oracle):
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    app = _app()
    for name in ("polygon1", "polygon2"):
        p = np.array(app[name], dtype=np.float32)
        for key, want in kat[name].items():
            off, prime = (int(x) for x in key.split(","))
            r = oracle.polygon_batch(p, np.array([0, len(p)]), offset_prime=[off, prime], want_stats=True)
            assert r["ids"].tolist() == want["ids"] and int(r["status"][0]) == want["status"]
            assert r["bbox"].view(np.uint32)[0].tolist() == want["bbox_bits"]
            assert hashlib.sha256(r["vtx"].tobytes()).hexdigest() == want["vtx_sha256"]
            assert r["stats"]["nodes"] == want["nodes"]
    # polygon1 is triangulated into the same 5 triangles whatever the edge order
    sets = {frozenset(tuple(t) for t in np.array(v["ids"]).reshape(-1, 3).tolist()) for v in kat["polygon1"].values()}
    assert len(sets) == 1


def _shoelace(P):
    x, y = P[:, 0].astype(np.float64), P[:, 1].astype(np.float64)
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def test_app_polygons_are_valid_triangulations(oracle):
    app = _app()
    for name in ("polygon1", "polygon2"):
        P = np.array(app[name], dtype=np.float32)
        assert _shoelace(P) > 0  # "clockwise" of Triangulation.zig:443 == positive shoelace in raw (x,y)
        r = oracle.polygon_batch(P, np.array([0, len(P)]), offset_prime=[1, 1])
        T = P[r["ids"].reshape(-1, 3)].astype(np.float64)
        ar = 0.5 * ((T[:, 1, 0] - T[:, 0, 0]) * (T[:, 2, 1] - T[:, 0, 1]) - (T[:, 2, 0] - T[:, 0, 0]) * (T[:, 1, 1] - T[:, 0, 1]))
        assert abs(np.abs(ar).sum() - _shoelace(P)) < 1e-3 * _shoelace(P)
        assert (ar > 0).all()  # every triangle keeps the polygon's orientation


def test_independent_python_restatement_agrees(oracle):
    """oracle/pyref.py was transliterated from the Zig source separately from the C file; both must
    give the same emit sequence and the same failure classification."""
    from oracle import pyref

    rng = np.random.default_rng(42)
    atan2 = lambda y, x: np.float32(oracle.atan2f(y, x))
    checked = fails = 0
    for trial in range(70):
        n = int(rng.integers(3, 15))
        u1, u2 = rng.random(n), rng.random(n)
        th = 2 * np.pi * (np.arange(n) + 0.8 * u1 - 0.4) / n
        rad = (20 + 70 * u2) if trial % 2 else np.full(n, 60.0)
        P = np.stack([100 + rad * np.cos(th), 100 + rad * np.sin(th)], 1).astype(np.float32)
        off, prime = oracle.unirand_seed(n, 1234, trial)
        order = oracle.unirand_sequence(n, off, prime)
        ids, st = pyref.triangulate_ids(P, order, atan2=atan2)
        r = oracle.polygon_batch(P, np.array([0, n]), offset_prime=[off, prime])
        cap = 3 * (n - 2)
        if st == "null_unwrap":
            assert r["status"][0] & 4
            fails += 1
        else:
            exp = np.full(cap, 0xFFFFFFFF, dtype=np.uint32)
            m = min(cap, len(ids))
            exp[:m] = ids[:m]
            assert np.array_equal(exp, r["ids"]), (trial, n)
            assert (len(ids) > cap) == bool(r["status"][0] & 8)
            assert (len(ids) < cap) == bool(r["status"][0] & 128)
        checked += 1
    assert checked == 70 and fails > 0  # the sample includes inputs on which the reference algorithm fails


def test_structural_invariants_convex(oracle):
    """On convex input the reference algorithm is sound: n-2 triangles that tile the polygon, every
    emitted vertex an input vertex, for every edge order unirand can pick."""
    rng = np.random.default_rng(3)
    for n in (3, 5, 8, 13, 32, 100):
        th = 2 * np.pi * (np.arange(n) + 0.8 * rng.random(n) - 0.4) / n
        P = np.stack([100 + 60 * np.cos(th), 100 + 45 * np.sin(th)], 1).astype(np.float32)
        for trial in range(6):
            r = oracle.polygon_batch(P, np.array([0, n]), seed=trial, want_stats=True)
            assert r["status"][0] == 0 and r["ntri"][0] == n - 2
            ids = r["ids"].reshape(-1, 3)
            assert ids.max() < n
            T = P[ids].astype(np.float64)
            ar = 0.5 * ((T[:, 1, 0] - T[:, 0, 0]) * (T[:, 2, 1] - T[:, 0, 1]) - (T[:, 2, 0] - T[:, 0, 0]) * (T[:, 1, 1] - T[:, 0, 1]))
            assert abs(ar.sum() - _shoelace(P)) < 1e-4 * _shoelace(P) and (ar > -1e-9).all()
            assert r["stats"]["not_acute"] == 0


def test_statuses_and_caps(oracle):
    # degenerate / non-finite / too large
    xy = np.array([[0, 0], [1, 1], [0, 0], [1, 0], [np.nan, 1]], dtype=np.float32)
    r = oracle.polygon_batch(xy, np.array([0, 2, 5]), seed=1)
    assert r["status"].tolist() == [1, 2]
    # resource caps: on star 1024-gons the reference's DFS pushes the same trapezoid many times; the
    # contract abandons such a polygon once the stack or the node arena passes its cap
    fp = np.arange(31, dtype=np.uint64) * 1024
    P = oracle.synth_polygons(0x5EED0005, fp)
    r = oracle.polygon_batch(P, fp, seed=0x5EED0005, want_stats=True, nthreads=0)
    arena = (r["status"] & 64) != 0
    assert arena.any()
    per = 1022 * 96
    for i in np.where(arena)[0]:
        assert r["ntri"][i] == 0 and not r["vtx"][i * per:(i + 1) * per].any() and not r["bbox"][i].any()
    assert r["stats"]["max_stack"] <= 16 * 1024 + 64


def test_not_acute_corner_exists(oracle):
    """Triangulation.zig:403 is false for |atan2 - atan2| == (float)pi: reachable with finite input."""
    assert oracle.atan2f(5e-41, -1e10) == np.float32(math.pi)
    assert oracle.atan2f(1e-40, 1e10) == 0.0
    P = np.array([[0, 0], [2e10, 5e-41], [1e10, 1e-40]], dtype=np.float32)
    r = oracle.polygon_batch(P, np.array([0, 3]), seed=1, want_stats=True)
    assert r["stats"]["not_acute"] >= 1


def test_acute_shortcut_condition_is_sound(oracle):
    """The fast path skips atan2 in push_triangle_if_acute (Triangulation.zig:398-425) when both difference vectors
    satisfy  dy > 0 and (dx >= 0 or dy > 1e-6 * -dx),  or  dy == 0 and dx > 0  (triangulate.cu: angle_clear_of_pi).
    Its claim, checked here against the musl-exact atan2f of the oracle: such a vector's angle lies in
    [0, (f32)pi - 5e-7], so |a - b| <= max(a, b) < (f32)pi and the comparison is decided without evaluating it."""
    rng = np.random.default_rng(12)
    pi32 = float(np.float32(math.pi))
    worst = 0.0
    samples = []
    for mag in (1e-30, 1e-10, 1e-3, 1.0, 37.5, 1e6, 1e20, 1e35):
        for ratio in (1.0000001e-6, 1.001e-6, 1.5e-6, 1e-5, 1e-3, 0.5, 1.0, 3.0, 1e3, 1e6):
            for _ in range(40):
                dx = -mag * (1.0 + rng.random())
                dy = abs(dx) * ratio * (1.0 + 1e-3 * rng.random())
                samples.append((dy, dx))
        for _ in range(200):  # dx >= 0 branch, incl. dx = 0 and tiny dy
            samples.append((mag * rng.random() + 1e-45, mag * rng.random() * (rng.random() > 0.2)))
        samples.append((0.0, mag))  # dy == 0 and dx > 0
    for dy, dx in samples:
        with np.errstate(over="ignore"):  # 1e35 * 1e6 -> inf is a legitimate sample (atan2 = pi/2)
            dy32, dx32 = np.float32(dy), np.float32(dx)
        clear = (dy32 > 0 and (dx32 >= 0 or dy32 > np.float32(1e-6) * -dx32)) or (dy32 == 0 and dx32 > 0)
        if not clear:
            continue
        a = oracle.atan2f(float(dy32), float(dx32))
        assert 0.0 <= a <= pi32 - 5e-7, (dy, dx, a)
        worst = max(worst, a)
    assert worst > 3.14159  # the sample does reach the neighbourhood of the threshold


# ---- unirand ------------------------------------------------------------------------------------
PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 101, 103,
          107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179, 181, 191, 193, 197, 199, 211, 223,
          227, 229, 233, 239, 241, 251, 257, 263, 269, 271, 277, 281, 283, 293, 307, 311, 313, 317, 331, 337, 347,
          349, 353, 359, 367, 373, 379, 383, 389, 397, 401, 409, 419, 421, 431, 433, 439, 443, 449, 457, 461, 463,
          467, 479, 487, 491, 499, 503, 509, 521, 523, 541, 601, 659, 733, 809, 863, 941, 1013, 1069, 1151, 1283,
          1289, 1367, 1447, 1499, 1579, 1637, 1723, 429494501, 429493501, 429486647, 100001053, 100002421, 10001567]


def _splitmix_stream(seed, index):
    M = (1 << 64) - 1
    state = seed ^ ((0x9E3779B97F4A7C15 * (index + 1)) & M)
    while True:
        state = (state + 0x9E3779B97F4A7C15) & M
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        yield z >> 32


def _unirand_seed_py(top, seed, index):
    """unirand.zig:26-50 transcribed independently, drawing from the documented stream."""
    g = _splitmix_stream(seed, index)
    offset = next(g) % (top - 1) + 1
    best = 1
    for p in PRIMES:
        if p < top and top % p != 0 and next(g) % 3 > 0:
            best = p
    return offset, best


def test_unirand_seed_matches_reference_rule(oracle):
    assert len(PRIMES) == 123
    for top in list(range(2, 140)) + [509, 1013, 1024, 1723, 1724, 4096]:
        for index in (0, 5):
            assert oracle.unirand_seed(top, 0xC0FFEE, index) == _unirand_seed_py(top, 0xC0FFEE, index)


def test_unirand_is_a_permutation(oracle):
    for top in range(2, 130):
        off, prime = oracle.unirand_seed(top, 99, top)
        assert 1 <= off <= top - 1 and (prime == 1 or (prime < top and top % prime != 0))
        seq = oracle.unirand_sequence(top, off, prime)
        assert sorted(seq) == list(range(top))
        assert seq[0] == off % top
    assert oracle.unirand_sequence(0, 0, 1) == []


def test_atan2f_restatement(oracle):
    rng = np.random.default_rng(0)
    y = (rng.standard_normal(4000) * 10.0 ** rng.integers(-6, 6, 4000)).astype(np.float32)
    x = (rng.standard_normal(4000) * 10.0 ** rng.integers(-6, 6, 4000)).astype(np.float32)
    got = np.array([oracle.atan2f(float(a), float(b)) for a, b in zip(y, x)], dtype=np.float32)
    want = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
    assert (np.abs(got.astype(np.float64) - want) <= 1.01 * ulp).all()
    assert oracle.atan2f(0.0, 1.0) == 0.0 and oracle.atan2f(0.0, -1.0) == np.float32(math.pi)
    assert oracle.atan2f(1.0, 0.0) == np.float32(np.float32(math.pi) / 2)
    assert oracle.atan2f(1.0, 1.0) == np.float32(0.7853981852531433)


# ---- terrain --------------------------------------------------------------------------------------
def _terrain_numpy(h16, n, gs=np.float32(0.2), os_=np.float32(0.1), hs=np.float32(5.0)):
    """Vectorised float32 restatement of the terrain spec (independent of the C oracle)."""
    f = np.float32
    h = (f(1.0) - (h16.astype(np.float32) / f(65535.0))).astype(np.float32).reshape(n, n)
    org = f(os_ * f(n))
    ar = np.arange(n)
    x = (gs * ar.astype(np.float32) - org).astype(np.float32)
    pos = np.zeros((n, n, 3), dtype=np.float32)
    pos[:, :, 0] = x[:, None]
    pos[:, :, 1] = hs * h
    pos[:, :, 2] = x[None, :]
    rm, rp = np.maximum(ar - 1, 0), np.minimum(ar + 1, n - 1)
    with np.errstate(invalid="ignore", divide="ignore"):
        gx = (hs * (h[rp, :] - h[rm, :])).astype(np.float32) / (gs * (rp - rm).astype(np.float32))[:, None]
        gz = (hs * (h[:, rp] - h[:, rm])).astype(np.float32) / (gs * (rp - rm).astype(np.float32))[None, :]
    if n == 1:
        gx[:] = 0
        gz[:] = 0
    ln = np.sqrt(((gx * gx).astype(np.float32) + f(1.0)).astype(np.float32) + (gz * gz).astype(np.float32)).astype(np.float32)
    inv = (f(1.0) / ln).astype(np.float32)
    nrm = np.stack([(-gx) * inv, inv, (-gz) * inv], -1).astype(np.float32)
    r, c = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    i00 = (r * n + c).astype(np.uint32)
    idx = np.stack([i00 + n, i00, i00 + n + 1, i00 + n + 1, i00, i00 + 1], -1).reshape(-1)
    return pos, nrm, idx


def test_terrain_oracle_vs_numpy_restatement(oracle):
    for n in (1, 2, 3, 17, 100):
        h = np.load(os.path.join(GOLDEN, "heightmap_100.npy")) if n == 100 else oracle.synth_heightmap_u16(9, n)
        vtx, idx = oracle.terrain_build(h, n)
        pos, nrm, widx = _terrain_numpy(h, n)
        V = vtx.reshape(n * n, 32)
        assert np.array_equal(V[:, 0:12].copy().view(np.uint32), pos.reshape(-1, 3).view(np.uint32))
        assert np.array_equal(V[:, 16:28].copy().view(np.uint32), nrm.reshape(-1, 3).view(np.uint32))
        assert not V[:, 12:16].any() and not V[:, 28:32].any()
        assert np.array_equal(idx, widx)
        ln = np.linalg.norm(nrm.astype(np.float64), axis=-1)
        assert np.abs(ln - 1).max() < 1e-6


def test_terrain_golden_hashes(oracle):
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["terrain_100"]
    h = np.load(os.path.join(GOLDEN, "heightmap_100.npy"))
    assert h.min() == 2313 and h.max() == 65535 and int(h.astype(np.float64).mean()) == 47987
    vtx, idx = oracle.terrain_build(h, 100)
    assert hashlib.sha256(vtx.tobytes()).hexdigest() == kat["vtx_sha256"]
    assert hashlib.sha256(idx.tobytes()).hexdigest() == kat["idx_sha256"]
    assert len(idx) == 58806 and idx.max() == 9999
    assert idx[:6].tolist() == [100, 0, 101, 101, 0, 1]  # (r+1,c) (r,c) (r+1,c+1) (r+1,c+1) (r,c) (r,c+1)


def test_terrain_matches_shader_stream(oracle):
    """Indexed mesh expanded through the index buffer == WGSL formula per shader vertex (Terrain.zig:24-48)."""
    n = 23
    h16 = oracle.synth_heightmap_u16(1, n)
    hf = oracle.heightmap_normalize(h16)
    vtx, idx = oracle.terrain_build(h16, n)
    V = vtx.reshape(n * n, 32)[:, :12].copy().view(np.float32)
    k = 0
    for r in range(n - 1):
        for c in range(n - 1):
            for corner in range(6):
                want = oracle.terrain_shader_vertex(hf, n, (r * n + c) * 6 + corner)
                assert np.array_equal(V[idx[k]].view(np.uint32), want[:3].view(np.uint32))
                k += 1
    # the reference's draw also covers quads with r = n-1 (reads past the heightmap): no such lookup here
    assert oracle.terrain_shader_vertex(hf, n, ((n - 1) * n + 3) * 6 + 0) is None


def test_terrain_bands_equal_whole(oracle):
    n = 61
    h = oracle.synth_heightmap_u16(2, n)
    vtx, idx = oracle.terrain_build(h, n)
    for (r0, r1) in ((0, 20), (20, 45), (45, 61)):
        lo, hi = max(r0 - 1, 0), min(r1 + 1, n)
        q0, q1 = r0, min(r1, n - 1)
        v, i = oracle.terrain_build(h[lo:hi], n, rows=(r0, r1), qrows=(q0, q1), height_row0=lo)
        assert np.array_equal(v, vtx[r0 * n * 32:r1 * n * 32])
        assert np.array_equal(i, idx[q0 * 6 * (n - 1):q1 * 6 * (n - 1)])


# ---- synthetic workloads -----------------------------------------------------------------------------
def test_synthetic_polygons_are_simple_and_positively_oriented(oracle):
    fp = oracle.synth_polygon_sizes(0x5EED0003, 500, 8, 64)
    n = np.diff(fp.astype(np.int64))
    assert n.min() >= 8 and n.max() <= 64 and abs(n.mean() - 36) < 2
    xy = oracle.synth_polygons(0x5EED0003, fp)
    for i in range(0, 500, 7):
        P = xy[fp[i]:fp[i + 1]].astype(np.float64)
        ang = np.unwrap(np.arctan2(P[:, 1] - 100, P[:, 0] - 100))
        assert (np.diff(ang) > 0).all() and ang[-1] - ang[0] < 2 * np.pi  # star-shaped about (100,100) => simple
        assert _shoelace(P) > 0
        rad = np.hypot(P[:, 0] - 100, P[:, 1] - 100)
        assert rad.min() >= 19.99 and rad.max() <= 90.01
    fpl = oracle.synth_polygon_sizes(0x5EED0005, 2000, 8, 1024, dist=1)
    nl = np.diff(fpl.astype(np.int64))
    assert nl.min() >= 8 and nl.max() <= 1024 and 150 < nl.mean() < 270


def test_independent_restatement_agrees_on_fuzzed_inputs_and_repeating_orders(oracle):
    """The same cross-check on the inputs the GPU fuzz draws (scripts/fuzz_parity.py): the three synthetic families under
    reversal, axis swap, quantised coordinates (ties), collinear runs, duplicated vertices and random rings -- and with
    explicit edge orders whose prime shares a factor with n, so that add_segment runs on an edge that is already in the
    DAG (the case the conflict-list kernels once got wrong: what the reference does there is pinned by two restatements)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "scripts"))
    import fuzz_parity as FP
    from oracle import pyref

    rng = np.random.default_rng(2026)
    atan2 = lambda y, x: np.float32(oracle.atan2f(y, x))
    fams = [oracle.FAMILY_STAR, oracle.FAMILY_ELLIPSE, oracle.FAMILY_ZIPPER]
    hows = ["identity", "reverse", "swap_axes", "quantise", "collinear", "duplicate", "noise", "noise_grid", "same_y"]
    checked = repeats = fails = 0
    for trial in range(160):
        n = int(rng.integers(3, 19))
        fp = np.array([0, n], dtype=np.uint64)
        P = FP.transform(rng, oracle.synth_polygons(1000 + trial, fp, family=fams[trial % 3]), fp, hows[trial % len(hows)])
        P = np.ascontiguousarray(P, dtype=np.float32).reshape(n, 2)
        if trial % 2:
            off, prime = int(rng.integers(0, n)), int(rng.choice([0, 1, 2, 3, 4, 6]))
            repeats += int(np.gcd(prime, n) != 1)
        else:
            off, prime = oracle.unirand_seed(n, 77, trial)
        order = oracle.unirand_sequence(n, off, prime)
        r = oracle.polygon_batch(P, fp, offset_prime=[off, prime])
        st_c = int(r["status"][0])
        if st_c & ~(4 | 8 | 128):  # arena caps, coincident-point or non-finite statuses: outside pyref's vocabulary
            continue
        ids, st = pyref.triangulate_ids(P, order, atan2=atan2)
        cap = 3 * (n - 2)
        if st == "null_unwrap":
            assert st_c & 4, (trial, n, off, prime)
            fails += 1
        else:
            exp = np.full(max(cap, 1), 0xFFFFFFFF, dtype=np.uint32)
            m = min(cap, len(ids))
            exp[:m] = ids[:m]
            assert np.array_equal(exp[:cap], r["ids"][:cap]), (trial, n, off, prime, hows[trial % len(hows)])
            assert (len(ids) > cap) == bool(st_c & 8) and (len(ids) < cap) == bool(st_c & 128)
        checked += 1
    assert checked >= 120 and repeats >= 20 and fails > 0
