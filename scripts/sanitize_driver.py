"""Small workload that touches every kernel and every polygon tier, for compute-sanitizer.

    compute-sanitizer --tool {memcheck,racecheck,synccheck,initcheck} python scripts/sanitize_driver.py

Host (numpy) buffers only -- the C ABI stages them -- so torch's allocator stays out of the report.
Results are compared with the CPU oracle (this script is test infrastructure, like tests/).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import myrenderer_b200 as mr  # noqa: E402
from oracle import oracle as O  # noqa: E402

SCALE = int(os.environ.get("MR_SANITIZE_SCALE", "1"))


def convex(n, rng):
    th = 2 * np.pi * (np.arange(n) + 0.8 * rng.random(n) - 0.4) / n
    a, b, ph = 40 + 50 * rng.random(), 40 + 50 * rng.random(), 6.28 * rng.random()
    x, y = a * np.cos(th), b * np.sin(th)
    return np.stack([100 + np.cos(ph) * x - np.sin(ph) * y, 100 + np.sin(ph) * x + np.cos(ph) * y], 1).astype(np.float32)


def main():
    ctx = mr.Context(0, use_torch_stream=False)
    lib = ctx.lib
    bad = 0

    # ---- terrain: the reference heightmap, odd sizes, u16 and f32 ---------------------------------
    T = mr.Terrain(ctx)
    h100 = np.load(os.path.join(ROOT, "tests", "golden", "heightmap_100.npy"))
    for h, n in [(h100, 100), (O.synth_heightmap_u16(1, 257), 257), (O.synth_heightmap_u16(2, 31), 31),
                 (O.heightmap_normalize(O.synth_heightmap_u16(3, 130)), 130), (O.synth_heightmap_u16(4, 1), 1)]:
        vtx = np.zeros(n * n * 32, dtype=np.uint8)
        idx = np.zeros(max(6 * (n - 1) * (n - 1), 1), dtype=np.uint32)
        T.build(T.job(h, n, vtx_out=vtx, idx_out=idx if n > 1 else None))
        ctx.sync()
        ov, oi = O.terrain_build(h, n)
        ok = np.array_equal(vtx, ov) and np.array_equal(idx[: len(oi)], oi)
        print(f"terrain n={n} {h.dtype}: {'ok' if ok else 'MISMATCH'}")
        bad += not ok
    out = np.zeros(h100.size, dtype=np.float32)
    ctx.check(lib.mr_heightmap_normalize(ctx.handle, h100.ctypes.data, h100.size, out.ctypes.data), "normalize")
    ctx.sync()
    bad += not np.array_equal(out, O.heightmap_normalize(h100).reshape(-1))

    # ---- polygons: star (incl. failing / exploding searches), convex in every size tier, coincident
    # points (general path), > 1024 points (global-memory tier), degenerate ----------------------------
    rng = np.random.default_rng(5)
    polys = []
    fp = O.synth_polygon_sizes(0x5EED0003, 150 * SCALE, 8, 64)
    sxy = O.synth_polygons(0x5EED0003, fp)
    polys += [sxy[int(fp[i]):int(fp[i + 1])] for i in range(len(fp) - 1)]
    for n in [3, 4, 8, 33, 64, 65, 100, 128, 200, 256, 300, 512, 700, 1024, 1025, 1500]:
        polys.append(convex(n, rng))
    fpl = O.synth_polygon_sizes(11, 6 * SCALE, 65, 600, dist=1)
    lxy = O.synth_polygons(11, fpl)
    polys += [lxy[int(fpl[i]):int(fpl[i + 1])] for i in range(len(fpl) - 1)]  # large star polygons
    dup = convex(20, rng)
    dup[7] = dup[3]
    polys.append(dup)                                              # coincident points
    polys.append(np.array([[0, 0], [1, 1]], dtype=np.float32))     # n < 3
    polys.append(np.array([[10, 10], [40, 10], [40, 40], [10, 40]], dtype=np.float32))  # App.zig polygon2
    xy = np.concatenate(polys).astype(np.float32)
    first_point = np.zeros(len(polys) + 1, dtype=np.uint64)
    first_point[1:] = np.cumsum([len(p) for p in polys])
    first_tri = mr.polygon_offsets_host(first_point)
    npoly, ntri = len(polys), int(first_tri[-1])
    P = mr.Polygon(ctx)
    vtx = np.zeros(ntri * 3 * 32, dtype=np.uint8)
    bbox = np.zeros((npoly, 4), dtype=np.float32)
    status = np.zeros(npoly, dtype=np.uint32)
    nt = np.zeros(npoly, dtype=np.uint32)
    P.triangulate(P.job(xy, first_point, npoly, vtx_out=vtx, first_tri=first_tri, bbox_out=bbox, status_out=status,
                        ntri_out=nt, seed=77))
    ctx.sync()
    o = O.polygon_batch(xy, first_point, seed=77, nthreads=0, want_ids=False)
    ok = (np.array_equal(vtx, o["vtx"]) and np.array_equal(status, o["status"]) and np.array_equal(nt, o["ntri"])
          and np.array_equal(bbox.view(np.uint32), o["bbox"].view(np.uint32)))
    tiers = (C.c_uint32 * 8)()
    lib.mr_triangulate_tier_counts(ctx.handle, tiers)
    print(f"polygons {npoly} ({int(first_point[-1])} points): {'ok' if ok else 'MISMATCH'}; "
          f"status ok {int((status == 0).sum())}; tiers {list(tiers)}")
    bad += not ok

    # ---- device-side generators and unirand ------------------------------------------------------------
    op = np.zeros(2 * npoly, dtype=np.uint32)
    ctx.check(lib.mr_unirand_seed_batch(ctx.handle, first_point.ctypes.data, npoly, 77, 0, op.ctypes.data), "unirand batch")
    g16 = np.zeros((40, 64), dtype=np.uint16)
    ctx.check(lib.mr_synth_heightmap_u16(ctx.handle, 9, 64, 3, 40, g16.ctypes.data), "synth heightmap")
    gxy = np.zeros_like(sxy)
    ctx.check(lib.mr_synth_polygons(ctx.handle, 0x5EED0003, 0, fp.ctypes.data, len(fp) - 1, gxy.ctypes.data), "synth polygons")
    ctx.sync()
    bad += not np.array_equal(g16, O.synth_heightmap_u16(9, 64, 3, 40))
    bad += not np.array_equal(gxy, sxy)
    ctx.close()
    print("RESULT", "ok" if bad == 0 else f"{bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
