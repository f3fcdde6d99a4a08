"""One mr_triangulate_batch over a sample of BASELINE config 5 (log-uniform 8..1024 points), for ncu.

    python scripts/profile_config5.py [npoly] [family: 1 ellipse | 2 zipper | 0 star] [nmin] [nmax]
Prints the event-timed duration of the measured call; run it under
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python scripts/profile_config5.py
for the per-kernel (= per size class) split, or under `ncu --set full -k regex:triangulate_team_k ...` for one kernel.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import myrenderer_b200 as mr

npoly = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
family = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nmin = int(sys.argv[3]) if len(sys.argv) > 3 else 8
nmax = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
dist = int(sys.argv[5]) if len(sys.argv) > 5 else 1
seed = 0x5EED0005
ctx = mr.Context(0)
lib = ctx.lib
fp = np.zeros(npoly + 1, dtype=np.uint64)
lib.mr_synth_polygon_sizes(seed, 0, npoly, nmin, nmax, dist, fp.ctypes.data)
ft = mr.polygon_offsets_host(fp)
dev = torch.device("cuda", 0)
fp_d = torch.from_numpy(fp.view(np.int64)).to(dev)
ft_d = torch.from_numpy(ft.view(np.int64)).to(dev)
xy = torch.empty(int(fp[-1]) * 2, dtype=torch.float32, device=dev)
ctx.check(lib.mr_synth_polygons_family(ctx.handle, family, seed, 0, fp_d.data_ptr(), npoly, xy.data_ptr()), "synth")
pv = torch.empty(int(ft[-1]) * 96, dtype=torch.uint8, device=dev)
st = torch.empty(npoly, dtype=torch.int32, device=dev)
P = mr.Polygon(ctx)
job = P.job(xy, fp_d, npoly, vtx_out=pv, first_tri=ft_d, status_out=st, seed=seed)
P.triangulate(job)
ctx.sync()
ts = []
for _ in range(int(os.environ.get("REPS", "1"))):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    P.triangulate(job)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"npoly {npoly} family {family} n {nmin}..{nmax} points {int(fp[-1])}: {min(ts):.3f} ms, ok {int((st == 0).sum())}")
ctx.close()
