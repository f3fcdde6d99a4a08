"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run).

Every rank builds its terrain row band and its polygon range STRAIGHT INTO RANK 0's buffers through an
IPC-mapped peer pointer (mr_ipc_export / mr_ipc_open: NVLink stores from the kernels, no staging, no NCCL);
rank 0 then compares the gathered buffers with the CPU oracle byte for byte.  The oracle is the checker only.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import myrenderer_b200 as mr  # noqa: E402
from myrenderer_b200 import sharding  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    dev = torch.device("cuda", local)
    ctx = mr.Context(local)
    lib = ctx.lib

    def shared(nbytes):
        base = C.c_void_p()
        h = torch.zeros(64, dtype=torch.uint8)
        if rank == 0:
            ctx.check(lib.mr_device_alloc(ctx.handle, nbytes, C.byref(base)), "alloc")
            ctx.check(lib.mr_fill_zero(ctx.handle, base, nbytes), "zero")
            ctx.sync()
            hb = (C.c_ubyte * 64)()
            ctx.check(lib.mr_ipc_export(ctx.handle, base, hb), "export")
            h.copy_(torch.frombuffer(bytearray(hb), dtype=torch.uint8))
        dist.broadcast(h, src=0)
        if rank != 0:
            hb = (C.c_ubyte * 64).from_buffer_copy(bytes(h.numpy().tobytes()))
            ctx.check(lib.mr_ipc_open(ctx.handle, hb, C.byref(base)), "open")
        return base

    def fetch(base, nbytes):
        out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ctx.check(lib.mr_copy(ctx.handle, out.data_ptr(), base, nbytes), "copy")
        ctx.sync()
        return out.cpu().numpy()

    # ---- terrain: row bands with a one-row halo -----------------------------------------------------------
    n = 1031  # odd size: bands of unequal height, ragged last tile
    sh = sharding.plan_terrain(n, rank, world)
    (r0, r1), (lo, hi) = sh.rows, sh.halo_rows
    height = torch.empty((hi - lo) * n, dtype=torch.int16, device=dev)
    ctx.check(lib.mr_synth_heightmap_u16(ctx.handle, 0x5EED0004, n, lo, hi - lo, height.data_ptr()), "synth")
    gv = shared(n * n * 32)
    T = mr.Terrain(ctx)
    gidx = torch.empty(6 * (n - 1) ** 2, dtype=torch.int32, device=dev) if rank == 0 else None
    T.build(T.job(height, n, rows=(r0, r1), qrows=(0, n - 1) if rank == 0 else (0, 0), height_row0=lo, height_rows=hi - lo,
                  vtx_out=gv.value, vtx_row0=0, idx_out=gidx, idx_qrow0=0))
    ctx.sync()
    dist.barrier()

    # ---- polygons: cost-balanced contiguous ranges, global triangle offsets -------------------------------
    seed = 0x5EED0005
    npoly = 3000
    fp_all = np.zeros(npoly + 1, dtype=np.uint64)
    lib.mr_synth_polygon_sizes(seed, 0, npoly, 3, 1024, 1, fp_all.ctypes.data)
    ft_all = mr.polygon_offsets_host(fp_all)
    ps = sharding.plan_polygons(fp_all, ft_all, rank, world)
    fp = np.ascontiguousarray(fp_all[ps.begin:ps.end + 1])
    ftg = np.ascontiguousarray(ft_all[ps.begin:ps.end + 1])
    cnt = ps.end - ps.begin
    fp_d = torch.from_numpy(fp.view(np.int64)).to(dev)
    ftg_d = torch.from_numpy(ftg.view(np.int64)).to(dev)
    xy = torch.empty(int(fp[-1] - fp[0]) * 2, dtype=torch.float32, device=dev)
    ctx.check(lib.mr_synth_polygons_family(ctx.handle, 2, seed, ps.begin, fp_d.data_ptr(), cnt, xy.data_ptr()), "synth")
    gp = shared(int(ft_all[-1]) * 96)
    gs = shared(npoly * 4)
    P = mr.Polygon(ctx)
    P.triangulate(P.job(xy, fp_d, cnt, vtx_out=gp.value, first_tri=ftg_d, tri_base=0, status_out=gs.value + 4 * ps.begin,
                        seed=seed, poly_index0=ps.begin, point_base=int(fp[0])))
    ctx.sync()
    dist.barrier()

    if rank == 0:
        from oracle import oracle as O  # the checker

        h_full = O.synth_heightmap_u16(0x5EED0004, n)
        ovtx, oidx = O.terrain_build(h_full, n, nthreads=0)
        assert np.array_equal(fetch(gv, n * n * 32), ovtx), "gathered terrain vertices differ from the oracle"
        assert np.array_equal(gidx.cpu().numpy().view(np.uint32), oidx), "rank 0's index buffer differs from the oracle"
        xy_all = O.synth_polygons(seed, fp_all, family=O.FAMILY_ZIPPER)
        ref = O.polygon_batch(xy_all, fp_all, seed=seed, nthreads=0, want_ids=False)
        assert np.array_equal(fetch(gp, int(ft_all[-1]) * 96), ref["vtx"]), "gathered polygon vertices differ from the oracle"
        assert np.array_equal(fetch(gs, npoly * 4).view(np.uint32), ref["status"]), "gathered status differs"
        assert (ref["status"][fp_all[1:] - fp_all[:-1] >= 3] == 0).all()
        print(f"gather ok: {world} ranks, terrain {n}x{n}, {npoly} polygons")
    dist.barrier()
    for b in (gv, gp, gs):
        if rank != 0:
            lib.mr_ipc_close(ctx.handle, b)
    dist.barrier()
    if rank == 0:
        for b in (gv, gp, gs):
            lib.mr_device_free(ctx.handle, b)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
