"""BASELINE configs 4 and 5 at full size, strong-scaled over the GPUs of one box.

    python scripts/bench_configs45.py                                  # one GPU
    python -m torch.distributed.run --nproc-per-node N ... scripts/bench_configs45.py

config 4: 16384 x 16384 hash-noise heightmap (seed 0x5EED0004), row bands with a one-row halo;
config 5: 1,000,000 polygons, sizes log-uniform in [8,1024] (seed 0x5EED0005), convex family,
          contiguous cost-balanced ranges.
Per config two times are reported (CUDA events, max over ranks): `compute` = every rank builds its
shard into its own HBM; `gather` = every rank builds its shard straight into rank 0's buffer
through an IPC-mapped peer pointer (NVLink stores), so the result is one buffer on rank 0.
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import myrenderer_b200 as mr
from myrenderer_b200 import sharding
from myrenderer_b200.workloads import ellipse_batch

sys.stdout.flush()
_OUT = os.dup(1)  # stdout carries only the JSON: library banners (NCCL version, ...) go to stderr
os.dup2(2, 1)
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
ctx = mr.Context(local)
lib = ctx.lib
ev = lambda: torch.cuda.Event(enable_timing=True)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rmax(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    barrier()
    ts = []
    for _ in range(reps):
        a, b = ev(), ev()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        barrier()
    return rmax(min(ts))


def shared_buffer(nbytes):
    """rank 0 allocates, everyone gets a pointer to it (IPC peer mapping on the other ranks)."""
    base = C.c_void_p()
    handle = torch.zeros(64, dtype=torch.uint8, device=dev)
    if rank == 0:
        ctx.check(lib.mr_device_alloc(ctx.handle, nbytes, C.byref(base)), "alloc")
        hb = (C.c_ubyte * 64)()
        ctx.check(lib.mr_ipc_export(ctx.handle, base, hb), "export")
        handle.copy_(torch.frombuffer(bytearray(hb), dtype=torch.uint8))
    if world > 1:
        dist.broadcast(handle, src=0)
        if rank != 0:
            hb = (C.c_ubyte * 64).from_buffer_copy(bytes(handle.cpu().numpy().tobytes()))
            ctx.check(lib.mr_ipc_open(ctx.handle, hb, C.byref(base)), "open")
    return base


def release(base):
    barrier()
    if rank != 0:
        lib.mr_ipc_close(ctx.handle, base)
    barrier()
    if rank == 0:
        lib.mr_device_free(ctx.handle, base)


out = {"n_gpus": world}

# ---- config 4 ----------------------------------------------------------------------------------------
n = 16384
sh = sharding.plan_terrain(n, rank, world)
(r0, r1), (q0, q1), (lo, hi) = sh.rows, sh.qrows, sh.halo_rows
height = torch.empty((hi - lo) * n, dtype=torch.int16, device=dev)
ctx.check(lib.mr_synth_heightmap_u16(ctx.handle, 0x5EED0004, n, lo, hi - lo, height.data_ptr()), "synth")
vtx = torch.empty((r1 - r0) * n * 32, dtype=torch.uint8, device=dev)
idx = torch.empty((q1 - q0) * 6 * (n - 1), dtype=torch.int32, device=dev)
T = mr.Terrain(ctx)
job_local = T.job(height, n, rows=(r0, r1), qrows=(q0, q1), height_row0=lo, height_rows=hi - lo, vtx_out=vtx, vtx_row0=r0,
                  idx_out=idx, idx_qrow0=q0)
ms_c = timed(lambda: T.build(job_local))
bytes4 = 34 * n * n + 24 * (n - 1) ** 2
res4 = {"compute_ms": ms_c, "compute_gverts_per_s": n * n / (ms_c * 1e-3) / 1e9, "compute_gb_per_s_aggregate": bytes4 / (ms_c * 1e-3) / 1e9}
if world > 1:
    gv = shared_buffer(n * n * 32)
    gi = shared_buffer(6 * (n - 1) ** 2 * 4)
    job_g = T.job(height, n, rows=(r0, r1), qrows=(q0, q1), height_row0=lo, height_rows=hi - lo, vtx_out=gv.value, vtx_row0=0,
                  idx_out=gi.value, idx_qrow0=0)
    ms_g = timed(lambda: T.build(job_g))
    nv = (n * n * 32 + 6 * (n - 1) ** 2 * 4) * (world - 1) / world
    res4.update({"gather_ms": ms_g, "gather_gverts_per_s": n * n / (ms_g * 1e-3) / 1e9, "nvlink_bytes_into_rank0": int(nv),
                 "rank0_ingest_gb_per_s": nv / (ms_g * 1e-3) / 1e9})
    release(gv)
    release(gi)
out["config4_terrain_16384"] = res4
del vtx, idx, height
torch.cuda.empty_cache()

# ---- config 5 ----------------------------------------------------------------------------------------
npoly = 1_000_000
fp_all = np.zeros(npoly + 1, dtype=np.uint64)
lib.mr_synth_polygon_sizes(0x5EED0005, 0, npoly, 8, 1024, 1, fp_all.ctypes.data)
ft_all = mr.polygon_offsets_host(fp_all)
ps = sharding.plan_polygons(fp_all, ft_all, rank, world)
fp = np.ascontiguousarray(fp_all[ps.begin:ps.end + 1])
ft_glob = np.ascontiguousarray(ft_all[ps.begin:ps.end + 1])
ft_loc = ft_glob - ft_glob[0]
cnt = ps.end - ps.begin
xy, _ = ellipse_batch(fp, 1234 + rank, device=dev)
fp_d = torch.from_numpy(fp.view(np.int64)).to(dev)
ftl_d = torch.from_numpy(ft_loc.view(np.int64)).to(dev)
ftg_d = torch.from_numpy(ft_glob.view(np.int64)).to(dev)
pv = torch.empty(int(ft_loc[-1]) * 96, dtype=torch.uint8, device=dev)
st = torch.empty(cnt, dtype=torch.int32, device=dev)
P = mr.Polygon(ctx)
job_l = P.job(xy, fp_d, cnt, vtx_out=pv, first_tri=ftl_d, status_out=st, seed=0x5EED0005, poly_index0=ps.begin, point_base=int(fp[0]))
ms_c = timed(lambda: P.triangulate(job_l), reps=2)
ok = int((st == 0).sum().item())
if world > 1:
    t = torch.tensor([ok], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    ok = int(t.item())
res5 = {"compute_ms": ms_c, "compute_polygons_per_s": npoly / (ms_c * 1e-3), "compute_mpoints_per_s": int(fp_all[-1]) / (ms_c * 1e-3) / 1e6,
        "status_ok": ok, "points": int(fp_all[-1]), "range_points_this_rank": int(fp[-1] - fp[0])}
if world > 1:
    total_bytes = int(ft_all[-1]) * 96
    gp = shared_buffer(total_bytes)
    job_g = P.job(xy, fp_d, cnt, vtx_out=gp.value, first_tri=ftg_d, tri_base=0, status_out=st, seed=0x5EED0005, poly_index0=ps.begin,
                  point_base=int(fp[0]))
    ms_g = timed(lambda: P.triangulate(job_g), reps=2)
    nv = total_bytes * (world - 1) / world
    res5.update({"gather_ms": ms_g, "gather_polygons_per_s": npoly / (ms_g * 1e-3), "nvlink_bytes_into_rank0": int(nv)})
    release(gp)
out["config5_polygons_1m"] = res5
if rank == 0:
    os.write(_OUT, (json.dumps(out, indent=1) + "\n").encode())
if world > 1:
    dist.destroy_process_group()
ctx.close()
