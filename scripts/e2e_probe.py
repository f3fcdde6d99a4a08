import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import myrenderer_b200 as mr
ctx = mr.Context(0); T = mr.Terrain(ctx); lib = ctx.lib
n = 4096
height = torch.empty(n * n, dtype=torch.int16, device="cuda")
ctx.check(lib.mr_synth_heightmap_u16(ctx.handle, 1, n, 0, n, height.data_ptr()), "synth")
h_h = torch.empty(n * n, dtype=torch.int16, pin_memory=True); h_h.copy_(height)
h_v = torch.empty(n * n * 32, dtype=torch.uint8, pin_memory=True)
h_i = torch.empty(6 * (n - 1) ** 2, dtype=torch.int32, pin_memory=True)
d_v = torch.empty(n * n * 32, dtype=torch.uint8, device="cuda"); d_i = torch.empty(6 * (n - 1) ** 2, dtype=torch.int32, device="cuda")
job_host = T.job(h_h, n, vtx_out=h_v, idx_out=h_i)
job_dev = T.job(height, n, vtx_out=d_v, idx_out=d_i)
def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
    return round(min(ts), 2), round(sum(ts) / len(ts), 2)
print("library host path      ", timeit(lambda: T.build(job_host)))
def manual():
    height.copy_(h_h, non_blocking=True); T.build(job_dev); h_v.copy_(d_v, non_blocking=True); h_i.copy_(d_i, non_blocking=True)
print("manual torch copies    ", timeit(manual))
print("D2H vtx only           ", timeit(lambda: h_v.copy_(d_v, non_blocking=True)))
print("D2H idx only           ", timeit(lambda: h_i.copy_(d_i, non_blocking=True)))
print("H2D only               ", timeit(lambda: height.copy_(h_h, non_blocking=True)))
print("kernels only           ", timeit(lambda: T.build(job_dev)))
