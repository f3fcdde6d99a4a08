"""CPU tests of the oracle's tile boxes and of the restated SceneNode visibility test."""
import numpy as np


def mat_image(rows):
    """mach Mat4x4.init(r0..r3) stores the ROWS it is given as four column vectors: image[4*col + row]."""
    return np.ascontiguousarray(np.array(rows, dtype=np.float32).T.reshape(16))


def perspective(fovy, aspect, near, far):  # math.zig:22-31
    f = np.float32
    halftan = f(np.tan(f(fovy) / f(2.0)))
    return [[f(1.0) / (f(aspect) * halftan), 0, 0, 0], [0, f(1.0) / halftan, 0, 0],
            [0, 0, f(far) / (f(far) - f(near)), -f(far) * f(near) / (f(far) - f(near))], [0, 0, 1, 0]]


def look_at(camera, target, up_ref):  # math.zig:9-20
    c, t, u = (np.array(v, dtype=np.float32) for v in (camera, target, up_ref))
    norm = lambda v: v / np.float32(np.sqrt(np.float32(np.dot(v, v))))
    fwd = norm(t - c)
    right = norm(np.cross(u, fwd))
    up = norm(np.cross(fwd, right))
    return [[*right, -np.dot(right, c)], [*up, -np.dot(up, c)], [*fwd, -np.dot(fwd, c)], [0, 0, 0, 1]]


def camera_matrix(cam=(30.0, 25.0, -40.0), target=(0.0, 0.0, 0.0)):
    m = np.array(perspective(1.2, 1.0, 0.1, 200.0), dtype=np.float32) @ np.array(look_at(cam, target, (0, 1, 0)), dtype=np.float32)
    return mat_image(m)


def test_tile_boxes_are_tight_and_cover_the_terrain(oracle):
    n, tr, tc = 97, 16, 20  # ragged last tiles in both directions
    h = oracle.synth_heightmap_u16(0x5EED0001, n)
    box = oracle.terrain_tile_bounds(h, n, tr, tc)
    vtx, _ = oracle.terrain_build(h, n, want_idx=False)
    pos = vtx.view(np.float32).reshape(n, n, 8)[:, :, :3]
    tiles_r, tiles_c = oracle.tile_count(n, tr, tc)
    assert box.shape == (tiles_r * tiles_c, 8) and (box[:, 3] == 1).all() and (box[:, 7] == 1).all()
    for t in range(tiles_r * tiles_c):
        a, b = divmod(t, tiles_c)
        r0, c0 = a * tr, b * tc
        r1, c1 = min(r0 + tr, n - 1), min(c0 + tc, n - 1)
        p = pos[r0:r1 + 1, c0:c1 + 1].reshape(-1, 3)
        assert np.array_equal(box[t, :3], p.min(0)) and np.array_equal(box[t, 4:7], p.max(0))  # tight, bit-exact
    # every tile lies inside the terrain's own box (Terrain.zig:103-110: (-bound,0,-bound)..(bound,5,bound))
    bound = np.float32(n) * np.float32(0.1)
    assert (box[:, 0] >= -bound).all() and (box[:, 2] >= -bound).all() and (box[:, 1] >= 0).all()
    assert (box[:, 4] <= bound).all() and (box[:, 6] <= bound).all() and (box[:, 5] <= 5).all()


def test_visibility_test_as_written(oracle):
    """SceneNode.zig:96-110: should_render = all(M*p1 > 0) or all(M*p0 < 1), all FOUR components, no divide; a box with
    an infinite component is not transformed (the default node box is +-inf and therefore always rendered)."""
    ident = mat_image(np.eye(4))
    inf = np.inf
    assert oracle.scene_node_should_render(ident, [-inf, -inf, -inf, 1], [inf, inf, inf, 1])       # SceneNode.zig:11-22
    assert oracle.scene_node_should_render(ident, [2, 2, 2, 1], [3, 3, 3, 1])                      # p1 > 0 everywhere
    assert oracle.scene_node_should_render(ident, [-3, -3, -3, 0.5], [-2, -2, -2, 1])              # p0 < 1 everywhere (w too)
    assert not oracle.scene_node_should_render(ident, [-3, 5, -3, 1], [-2, 6, -2, 1])              # p1.x < 0 and p0.y >= 1
    assert not oracle.scene_node_should_render(ident, [-3, -3, -3, 1], [-2, -2, -2, 1])            # p0.w == 1 is not < 1
    # the transform is mach's mulVec on the column image: result[i] = sum_j m[4j+i] * v[j], j ascending from 0
    rng = np.random.default_rng(3)
    for _ in range(200):
        m = rng.normal(size=16).astype(np.float32)
        p0 = np.append(rng.normal(size=3) * 3, 1).astype(np.float32)
        p1 = np.append(rng.normal(size=3) * 3, 1).astype(np.float32)
        f = np.float32

        def mul(v):
            out = []
            for i in range(4):
                acc = f(0)
                for j in range(4):
                    acc = f(acc + f(m[4 * j + i] * v[j]))
                out.append(acc)
            return np.array(out, dtype=np.float32)

        want = bool((mul(p1) > 0).all() or (mul(p0) < 1).all())
        assert oracle.scene_node_should_render(m, p0, p1) == want


def test_cull_compaction_and_index_ranges(oracle):
    n, tr, tc = 130, 32, 24
    h = oracle.synth_heightmap_u16(0x5EED0001, n)
    box = oracle.terrain_tile_bounds(h, n, tr, tc)
    _, full = oracle.terrain_build(h, n, want_vtx=False)
    quads_full = full.reshape(-1, 6)
    # a matrix under which every tile passes: the compacted buffer is the full index buffer regrouped by tile
    allvis = oracle.terrain_cull(box, n, tr, tc, mat_image(np.array([[0, 0, 0, 1]] * 4, dtype=np.float32)))
    assert allvis["visible"].all() and int(allvis["counts"][1]) == full.size
    got = allvis["idx"].reshape(-1, 6)
    assert np.array_equal(got[np.lexsort(got.T[::-1])], quads_full[np.lexsort(quads_full.T[::-1])])
    # a real camera: a strict subset survives; ids ascending; every visible tile contributes its quads row-major
    res = oracle.terrain_cull(box, n, tr, tc, camera_matrix())
    vis = res["visible"].astype(bool)
    assert 0 < vis.sum() < vis.size
    assert np.array_equal(res["ids"], np.nonzero(vis)[0])
    tiles_r, tiles_c = oracle.tile_count(n, tr, tc)
    want = []
    for t in res["ids"]:
        a, b = divmod(int(t), tiles_c)
        for r in range(a * tr, min(a * tr + tr, n - 1)):
            want.append(quads_full[r * (n - 1) + b * tc: r * (n - 1) + min(b * tc + tc, n - 1)])
    want = np.concatenate(want)
    assert np.array_equal(res["idx"].reshape(-1, 6), want) and int(res["counts"][1]) == want.size
    none = oracle.terrain_cull(box, n, tr, tc, mat_image(-np.eye(4, dtype=np.float32) * 0 + np.diag([0, 0, 0, -1]).astype(np.float32)))
    # M*p = (0,0,0,-1): p1 not > 0, p0 < 1 everywhere -> visible through the second clause: the test is an OR
    assert none["visible"].all()
    none = oracle.terrain_cull(box, n, tr, tc, mat_image(np.array([[0, 0, 0, -1], [0, 0, 0, 2], [0, 0, 0, 0], [0, 0, 0, 1]], dtype=np.float32)))
    assert not none["visible"].any() and int(none["counts"][0]) == 0 and none["idx"].size == 0


def test_height_conversion_is_monotone_so_tile_extremes_can_be_taken_on_texels(oracle):
    """terrain_tile_bounds_wide_k reduces the raw u16 texels (packed min / max) and converts only the two results:
    exact because 1 - v/65535 (Terrain.zig:120) is monotone non-increasing over all 65,536 texel values, so the smallest
    height of a tile is the height of its largest texel and vice versa."""
    h = oracle.heightmap_normalize(np.arange(65536, dtype=np.uint16))
    assert (np.diff(h) <= 0).all() and h[0] == 1.0 and h[65535] == 0.0
    rng = np.random.default_rng(5)
    for _ in range(50):
        t = rng.integers(0, 65536, size=int(rng.integers(1, 400))).astype(np.uint16)
        ht = oracle.heightmap_normalize(t)
        assert ht.min() == h[int(t.max())] and ht.max() == h[int(t.min())]


def test_tile_box_zero_bounds_are_positive_zero_and_one_column_tiles(oracle):
    """A y bound that is zero is written as +0.0f whatever the order of the reduction (height_scale == 0, or a float map
    holding both zeros); tiles one quad wide give n - 1 tiles per row (the strip kernels take 512 of them per CTA)."""
    n = 40
    h = np.zeros((n, n), dtype=np.float32)
    h[::2] = np.float32(-0.0)
    h[1, 1] = np.float32(-3.0)
    for prm in ((0.2, 0.1, 0.0), (0.2, 0.1, 5.0), (0.2, 0.1, -2.0)):
        box = oracle.terrain_tile_bounds(h, n, 4, 1, prm)
        assert box.shape[0] == ((n - 1 + 3) // 4) * (n - 1)
        y = box[:, [1, 5]].copy().view(np.uint32)
        assert not (y == 0x80000000).any(), prm  # no negative zero
        assert (box[:, 1] <= box[:, 5]).all()
    u = oracle.synth_heightmap_u16(9, n)
    box = oracle.terrain_tile_bounds(u, n, 7, 1)
    hf = oracle.heightmap_normalize(u)
    for tc in (0, 17, n - 2):
        lo = hf[0:8, tc:tc + 2].min() * np.float32(5.0)
        assert box[tc, 1] == np.float32(lo)
