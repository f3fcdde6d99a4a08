"""Single-GPU timing of BASELINE configs 4 and 5 (full sizes), device-resident, CUDA events."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import myrenderer_b200 as mr
from oracle import oracle as O
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_fullsize import _ellipse_batch

ctx = mr.Context(0); lib = ctx.lib
ev = lambda: torch.cuda.Event(enable_timing=True)
out = {}
# config 4: 16384^2
n = 16384
height = torch.empty(n * n, dtype=torch.int16, device="cuda")
ctx.check(lib.mr_synth_heightmap_u16(ctx.handle, 0x5EED0004, n, 0, n, height.data_ptr()), "synth")
vtx = torch.empty(n * n * 32, dtype=torch.uint8, device="cuda"); idx = torch.empty(6 * (n - 1) ** 2, dtype=torch.int32, device="cuda")
T = mr.Terrain(ctx); jv = T.job(height, n, qrows=(0, 0), vtx_out=vtx); ji = T.job(height, n, rows=(0, 0), idx_out=idx)
for _ in range(3): T.build(jv); T.build(ji)
tv, ti = [], []
for _ in range(5):
    a, b, c = ev(), ev(), ev(); a.record(); T.build(jv); b.record(); T.build(ji); c.record(); torch.cuda.synchronize()
    tv.append(a.elapsed_time(b)); ti.append(b.elapsed_time(c))
bv, bi = 34 * n * n, 24 * (n - 1) ** 2
out["config4_terrain_16384"] = {"ms_vertices": min(tv), "ms_indices": min(ti), "gverts_per_s": n * n / ((min(tv) + min(ti)) * 1e-3) / 1e9,
    "vertices_gb_per_s": bv / (min(tv) * 1e-3) / 1e9, "indices_gb_per_s": bi / (min(ti) * 1e-3) / 1e9,
    "total_gb_per_s": (bv + bi) / ((min(tv) + min(ti)) * 1e-3) / 1e9}
del vtx, idx, height
# config 5: 1M polygons log-uniform 8..1024 (convex family)
npoly = 1_000_000
fp = O.synth_polygon_sizes(0x5EED0005, npoly, 8, 1024, dist=1)
xy, _ = _ellipse_batch(fp, 1234)
P = mr.Polygon(ctx); ft = mr.polygon_offsets_host(fp)
fp_d = torch.from_numpy(fp.view(np.int64)).cuda(); ft_d = torch.from_numpy(ft.view(np.int64)).cuda()
pv = torch.empty(int(ft[-1]) * 96, dtype=torch.uint8, device="cuda"); st = torch.empty(npoly, dtype=torch.int32, device="cuda")
job = P.job(xy, fp_d, npoly, vtx_out=pv, first_tri=ft_d, status_out=st, seed=0x5EED0005)
P.triangulate(job); torch.cuda.synchronize()
tp = []
for _ in range(3):
    a, b = ev(), ev(); a.record(); P.triangulate(job); b.record(); torch.cuda.synchronize(); tp.append(a.elapsed_time(b))
out["config5_polygons_1m"] = {"ms": min(tp), "polygons_per_s": npoly / (min(tp) * 1e-3), "points": int(fp[-1]),
    "mpoints_per_s": int(fp[-1]) / (min(tp) * 1e-3) / 1e6, "status_ok": int((st == 0).sum().item())}
print(json.dumps(out, indent=1))
