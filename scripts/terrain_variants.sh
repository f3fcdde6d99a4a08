#!/bin/bash
# experiment: terrain kernel variants (rows per tile, store cache policy); restores the default library afterwards
cp myrenderer_b200/lib/libmyrenderer_b200.so /tmp/lib_default.so
for f in gpurun_variants/lib_*.so; do
  cp $f myrenderer_b200/lib/libmyrenderer_b200.so
  python - "$f" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import myrenderer_b200 as mr
ctx = mr.Context(0); T = mr.Terrain(ctx); lib = ctx.lib
n = int(os.environ.get("TN", "4096"))
height = torch.empty(n*n, dtype=torch.int16, device="cuda")
lib.mr_synth_heightmap_u16(ctx.handle, 0x5EED0001, n, 0, n, height.data_ptr())
vtx = torch.empty(n*n*32, dtype=torch.uint8, device="cuda"); idx = torch.empty(6*(n-1)**2, dtype=torch.int32, device="cuda")
jv = T.job(height, n, qrows=(0,0), vtx_out=vtx); ji = T.job(height, n, rows=(0,0), idx_out=idx)
flush = torch.empty(256<<20, dtype=torch.uint8, device="cuda")
for _ in range(3): T.build(jv); T.build(ji); flush.zero_()
tv, ti = [], []
for _ in range(20):
    a,b,c = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    a.record(); T.build(jv); b.record(); T.build(ji); c.record(); flush.zero_(); torch.cuda.synchronize()
    tv.append(a.elapsed_time(b)); ti.append(b.elapsed_time(c))
print(sys.argv[1], "vertices us mean %.1f min %.1f  | indices mean %.1f" % (1e3*sum(tv)/len(tv), 1e3*min(tv), 1e3*sum(ti)/len(ti)))
PY
done
cp /tmp/lib_default.so myrenderer_b200/lib/libmyrenderer_b200.so
