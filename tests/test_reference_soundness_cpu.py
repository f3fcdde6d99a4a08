"""CPU tests that pin what can be pinned about the reference's behaviour on non-convex input.

1. The failure precondition of Triangulation.zig's segment search, as a geometric predicate on the input.
2. A non-convex family (MR_FAMILY_ZIPPER) on which the reference is sound for every edge order, up to 1024 points.
3. What the contract's resource caps (MR_NODE_CAP / MR_STACK_CAP) cut off, measured with the caps multiplied.
4. The distance between the NEW-SPEC normal formula and the three-quotient form SURVEY 8-a4 proposed.
"""
import numpy as np
import pytest


def _shoelace(P):
    x, y = P[:, 0].astype(np.float64), P[:, 1].astype(np.float64)
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def _ccw(a, b, c):
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def _is_simple(p):
    n = len(p)
    for i in range(n):
        a, b = p[i], p[(i + 1) % n]
        for j in range(n):
            if j == i or j == (i + 1) % n or (j + 1) % n == i:
                continue
            c, d = p[j], p[(j + 1) % n]
            if _ccw(a, b, c) * _ccw(a, b, d) <= 0 and _ccw(c, d, a) * _ccw(c, d, b) <= 0:
                return False
    return True


def _above(a, b):  # Triangulation.zig:128-136
    return (a[1] < b[1]) or (a[1] == b[1] and a[0] < b[0])


def _left_of(P, a, b):  # Triangulation.zig:117-126, separately rounded f32 operations
    f = np.float32
    return f(f(f(b[0]) - f(a[0])) * f(f(P[1]) - f(a[1]))) - f(f(f(b[1]) - f(a[1])) * f(f(P[0]) - f(a[0]))) > 0


def has_bad_containment_pair(p):
    """The failure precondition of add_segment's side test (Triangulation.zig:275-286).

    `bottom_is_below` (:276) is computed as point_is_above(lower, other.point2), i.e. it is TRUE when the
    new segment's lower point is ABOVE the other segment's lower point.  So the branch the comments call
    "contains the other one vertically" (:277-281) really handles overlap -- and tests a point that lies
    inside the other's y-range, which is fine -- while true containment (top above, bottom below) lands in
    the branch "our bottom point is adjacent to the line" (:282-285) and tests the new segment's LOWER
    endpoint against the other segment's line.  That endpoint lies below the other segment, so the test is
    against the *extension* of the line and gives the wrong side whenever the endpoint lies across it.
    Returns True when some ordered pair of non-adjacent edges (S contains O vertically) has that property."""
    n = len(p)
    E = []
    for i in range(n):
        a, b = p[i], p[(i + 1) % n]
        E.append((a, b) if _above(a, b) else (b, a))
    for i, (su, sl) in enumerate(E):
        for j, (ou, ol) in enumerate(E):
            if i == j or j == (i + 1) % n or i == (j + 1) % n:
                continue
            if _above(su, ou) and not _above(sl, ol):  # S strictly contains O in the (y, x) order
                wrong = _left_of(sl, ou, ol) != (not _left_of(ou, su, sl))  # as written vs the true side
                if wrong:
                    return True
    return False


def _all_orders(n):
    primes = [1] + [q for q in (2, 3, 5, 7, 11, 13) if q < n and n % q]
    return [(o, pr) for o in range(1, n) for pr in primes]


def test_failure_precondition_of_the_segment_search(oracle):
    """On 1,500 random simple polygons (4..8 integer points, every edge order unirand can produce): whenever
    the reference fails for SOME order, the input has a bad containment pair; without one it is sound for
    EVERY order.  (The converse does not hold: a bad pair has to be met by the search to do harm.)"""
    rng = np.random.default_rng(11)
    seen = {"fail_with": 0, "ok_with": 0, "ok_without": 0}
    for n in (4, 5, 6, 7, 8):
        cnt = 0
        while cnt < 300:
            p = rng.integers(0, 20, size=(n, 2)).astype(np.float32)
            if len({(a, b) for a, b in p}) < n or abs(_shoelace(p)) < 1 or not _is_simple(p):
                continue
            if _shoelace(p) < 0:
                p = p[::-1].copy()
            cnt += 1
            prs = _all_orders(n)
            fp = np.arange(0, (len(prs) + 1) * n, n, dtype=np.uint64)
            r = oracle.polygon_batch(np.concatenate([p] * len(prs)), fp, offset_prime=np.array(prs, dtype=np.uint32),
                                     want_ids=False)
            fail = bool((r["status"] != 0).any())
            bad = has_bad_containment_pair(p)
            assert not (fail and not bad), f"failure without the precondition: {p.tolist()}"
            seen["fail_with" if fail else ("ok_with" if bad else "ok_without")] += 1
    assert seen["fail_with"] > 300 and seen["ok_without"] > 500, seen  # both populations are represented


def test_smallest_failing_case_by_hand(oracle):
    """The quadrilateral (7,2),(1,6),(1,9),(0,0) with edge order 1,0,3,2 (offset 1, prime 3): inserting edge
    (3,2) meets segment node (0,1); top (0,0) is above (7,2), bottom (1,9) is below (1,6) -> :282-285 tests
    (1,9) against the line through (7,2),(1,6): d = (1-7)*(9-2) - (6-2)*(1-7) = -18 -> "right", although the
    new edge lies left of (0,1) wherever both exist.  Traced by hand from the Zig source."""
    p = np.array([[7, 2], [1, 6], [1, 9], [0, 0]], dtype=np.float32)
    assert _shoelace(p) > 0 and _is_simple(p) and has_bad_containment_pair(p)
    assert oracle.unirand_sequence(4, 1, 3) == [1, 0, 3, 2]
    r = oracle.polygon_batch(p, np.array([0, 4]), offset_prime=[1, 3])
    assert int(r["status"][0]) != 0
    r = oracle.polygon_batch(p, np.array([0, 4]), offset_prime=[2, 1])  # an order in which the pair never meets
    assert int(r["status"][0]) == 0


@pytest.mark.parametrize("family", ["ellipse", "zipper"])
def test_sound_families_every_order_small(oracle, family):
    fam = {"ellipse": oracle.FAMILY_ELLIPSE, "zipper": oracle.FAMILY_ZIPPER}[family]
    for n in (3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16):
        for k in range(6):
            p = oracle.synth_polygons(0x21BB + k, np.array([0, n], dtype=np.uint64), poly_index0=k, family=fam)
            assert _shoelace(p) > 0 and _is_simple(p.astype(np.float64))
            assert not has_bad_containment_pair(p)
            prs = _all_orders(n)
            fp = np.arange(0, (len(prs) + 1) * n, n, dtype=np.uint64)
            r = oracle.polygon_batch(np.concatenate([p] * len(prs)), fp, offset_prime=np.array(prs, dtype=np.uint32))
            assert (r["status"] == 0).all(), (family, n, k)
            # n-2 triangles that tile the polygon (areas add up, all with the polygon's orientation)
            ids = r["ids"].reshape(len(prs), n - 2, 3)
            q = p.astype(np.float64)
            tri = q[ids]
            a2 = ((tri[:, :, 1, 0] - tri[:, :, 0, 0]) * (tri[:, :, 2, 1] - tri[:, :, 0, 1])
                  - (tri[:, :, 1, 1] - tri[:, :, 0, 1]) * (tri[:, :, 2, 0] - tri[:, :, 0, 0]))
            assert np.allclose(np.abs(a2).sum(1) * 0.5, _shoelace(p), rtol=1e-9)


@pytest.mark.parametrize("family", ["ellipse", "zipper"])
def test_sound_families_up_to_1024_points(oracle, family):
    """Every member ends with status OK at every size class of the kernel, and stays far below the contract caps:
    on the families the reference handles, MR_POLY_ARENA can never be the reason for a difference."""
    fam = {"ellipse": oracle.FAMILY_ELLIPSE, "zipper": oracle.FAMILY_ZIPPER}[family]
    sizes = [17, 31, 64, 65, 128, 129, 168, 216, 288, 289, 368, 504, 608, 768, 1000, 1024]
    for n in sizes:
        reps = 12
        fp = np.arange(0, (reps + 1) * n, n, dtype=np.uint64)
        xy = oracle.synth_polygons(0x5EED0005, fp, family=fam)
        r = oracle.polygon_batch(xy, fp, seed=0x5EED0005, want_ids=False, want_stats=True, nthreads=0)
        assert (r["status"] == 0).all(), (family, n)
        assert r["stats"]["nodes"] / (reps * n) < 5.0  # contract cap: 8n + 64
        assert r["stats"]["max_stack"] <= (4 if family == "ellipse" else 2)  # contract cap: 16n + 64
    # mixed sizes, log-uniform like BASELINE config 5
    fp = oracle.synth_polygon_sizes(0x5EED0005, 1500, 8, 1024, dist=1)
    xy = oracle.synth_polygons(0x5EED0005, fp, family=fam)
    r = oracle.polygon_batch(xy, fp, seed=0x5EED0005, want_ids=False, nthreads=0)
    assert (r["status"] == 0).all()


def test_what_the_arena_caps_cut_off(oracle):
    """MR_POLY_ARENA is a DEFINED DIVERGENCE, not parity: the reference has no bound.  This test measures it
    instead of hiding it.  3,000 log-uniform star polygons (8..1024 points): those the contract caps abandon
    are re-run with the caps multiplied by 8 (a literally unbounded run does not end in practical time: the
    DFS re-pushes merged trapezoids once per DAG path and pass 2 is quadratic in the stack).
      * no polygon that ends OK ever came near the NODE cap (nodes stay ~5n): only the STACK cap binds;
      * under 1 % of the abandoned polygons would end OK with 8x caps; the others fail anyway (overflow,
        underfill, null unwrap, or still exploding)."""
    seed = 0x5EED0005
    npoly = 3000
    fp = oracle.synth_polygon_sizes(seed, npoly, 8, 1024, dist=1)
    xy = oracle.synth_polygons(seed, fp)
    r = oracle.polygon_batch(xy, fp, seed=seed, nthreads=0, want_ids=False)
    arena = np.nonzero(r["status"] & 64)[0]
    assert 200 < len(arena) < 900  # ~15 % of the star polygons explode
    fp2 = np.zeros(len(arena) + 1, dtype=np.uint64)
    parts = []
    for k, i in enumerate(arena):
        a, b = int(fp[i]), int(fp[i + 1])
        parts.append(xy[a:b])
        fp2[k + 1] = fp2[k] + (b - a)
    ops = np.array([oracle.unirand_seed(int(fp[i + 1] - fp[i]), seed, int(i)) for i in arena], dtype=np.uint32)
    oracle.lift_caps(8)
    try:
        r2 = oracle.polygon_batch(np.concatenate(parts), fp2, offset_prime=ops, nthreads=0, want_ids=False)
        ok = np.nonzero(r2["status"] == 0)[0]
        assert len(ok) <= 0.01 * len(arena) + 1, (len(ok), len(arena))
        for k in ok:  # the ones that would finish: few nodes, an enormous stack of duplicates
            a, b = int(fp2[k]), int(fp2[k + 1])
            n = b - a
            r3 = oracle.polygon_batch(np.concatenate(parts)[a:b], np.array([0, n], dtype=np.uint64),
                                      offset_prime=ops[k:k + 1], want_stats=True, want_ids=False)
            assert r3["stats"]["nodes"] < 8 * n + 64 and r3["stats"]["max_stack"] > 16 * n + 64
    finally:
        oracle.lift_caps(1)
    # and the capped oracle is back: same statuses as before
    r3 = oracle.polygon_batch(xy, fp, seed=seed, nthreads=0, want_ids=False)
    assert np.array_equal(r3["status"], r["status"])


def test_normals_vs_the_three_quotient_form(oracle):
    """SURVEY 8-a4 proposed n = (-gx, 1, -gz) / len with three IEEE quotients; the header, oracle and kernel use
    inv = 1/len and two products.  On a 1024^2 hash-noise map (large gradients, heavy cancellation) plus the
    reference's own heightmap the two differ by at most 1 ulp per component -- inside the north-star's 2 ulp."""
    import os

    def three_quotient(h, n):
        f = np.float32
        hp = np.pad(h, 1, mode="edge")
        r = np.arange(n)
        span = (np.minimum(r + 1, n - 1) - np.maximum(r - 1, 0)).astype(np.float32)
        gx = (f(5.0) * (hp[2:, 1:-1] - hp[:-2, 1:-1])) / (f(0.2) * span)[:, None]
        gz = (f(5.0) * (hp[1:-1, 2:] - hp[1:-1, :-2])) / (f(0.2) * span)[None, :]
        ln = np.sqrt((gx * gx + f(1.0)) + gz * gz)
        return np.stack([-gx / ln, f(1.0) / ln, -gz / ln], -1).astype(np.float32)

    def ulps(a, b):
        ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
        ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
        ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
        return np.abs(ia - ib)

    cases = [oracle.synth_heightmap_u16(0x5EED0001, 1024)]
    png = os.path.join(os.path.dirname(__file__), "golden", "heightmap_100.npy")
    if os.path.exists(png):
        cases.append(np.load(png).astype(np.uint16))
    worst = 0
    for u16 in cases:
        n = u16.shape[0]
        vtx, _ = oracle.terrain_build(u16, n, want_idx=False, nthreads=0)
        got = vtx.view(np.float32).reshape(n, n, 8)[:, :, 4:7]
        want = three_quotient(oracle.heightmap_normalize(u16), n)
        d = ulps(np.ascontiguousarray(got), np.ascontiguousarray(want))
        worst = max(worst, int(d.max()))
    assert worst <= 1, worst
