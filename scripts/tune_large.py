"""Timing of the polygon kernel on convex ellipses; `log`: sizes log-uniform 8..1024, `big`: 512..1024, `small`: 8..64
(the bench's convex batch).  Also the target of the ncu captures in profiles/."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import myrenderer_b200 as mr
from oracle import oracle as O

def ellipses(fp, seed):
    rng = np.random.default_rng(seed)
    n = np.diff(fp.astype(np.int64)); tot = int(fp[-1])
    pid = np.repeat(np.arange(len(n)), n); k = np.arange(tot) - np.repeat(fp[:-1].astype(np.int64), n)
    nn = n[pid].astype(np.float64)
    th = 2*np.pi*(k + 0.8*rng.random(tot) - 0.4)/nn
    a = (40+50*rng.random(len(n)))[pid]; b = (40+50*rng.random(len(n)))[pid]; ph = (rng.random(len(n))*6.28)[pid]
    x, y = a*np.cos(th), b*np.sin(th)
    return np.stack([100+np.cos(ph)*x-np.sin(ph)*y, 100+np.sin(ph)*x+np.cos(ph)*y], 1).astype(np.float32)

which = sys.argv[1] if len(sys.argv) > 1 else "log"
npoly = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
fp = (O.synth_polygon_sizes(0x5EED0005, npoly, 8, 1024, 1) if which == "log" else
      O.synth_polygon_sizes(0x5EED0003, npoly, 8, 64, 0) if which == "small" else O.synth_polygon_sizes(1, npoly, 512, 1024, 0))
xy = ellipses(fp, 3)
ctx = mr.Context(0); P = mr.Polygon(ctx)
for tune in ("default",):
    b = P.create_polygons(xy, fp, seed=5); ctx.sync()
    ok = int((b.status == 0).sum().item())
    import ctypes as C
    tc = (C.c_uint32 * 8)(); ctx.lib.mr_triangulate_tier_counts(ctx.handle, tc); print("tier counts", list(tc))
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xy_d = torch.from_numpy(xy).cuda(); ft = mr.polygon_offsets_host(fp)
        fp_d = torch.from_numpy(fp.view(np.int64)).cuda(); ft_d = torch.from_numpy(ft.view(np.int64)).cuda()
        vtx = torch.empty(int(ft[-1])*96, dtype=torch.uint8, device="cuda")
        st = torch.empty(npoly, dtype=torch.int32, device="cuda")
        j = P.job(xy_d, fp_d, npoly, vtx_out=vtx, first_tri=ft_d, status_out=st, seed=5)
        torch.cuda.synchronize(); e0.record(); P.triangulate(j); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(which, tune, "ok", ok, "ms", round(min(ts), 3), "Mpoly/s", round(npoly/min(ts)/1e3, 3), "Mpts/s", round(int(fp[-1])/min(ts)/1e3, 1), flush=True)
