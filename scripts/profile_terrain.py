"""One 4096^2 terrain build (u16 heightmap, 32-byte vertices + indices) for ncu captures of the terrain kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import myrenderer_b200 as mr

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = mr.Context(0)
h = torch.empty(n * n, dtype=torch.int16, device="cuda")
ctx.check(ctx.lib.mr_synth_heightmap_u16(ctx.handle, 0x5EED0001, n, 0, n, h.data_ptr()), "synth")
T = mr.Terrain(ctx)
vtx = torch.empty(n * n * 32, dtype=torch.uint8, device="cuda")
idx = torch.empty(6 * (n - 1) ** 2, dtype=torch.int32, device="cuda")
job = T.job(h, n, vtx_out=vtx, idx_out=idx)
for _ in range(3):
    T.build(job)
ctx.sync()
print("done")
ctx.close()
