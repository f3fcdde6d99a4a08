"""The N>1 path on CPU: world_size-2 gloo.  Each rank plans its shard with the library's partition
helpers (myrenderer_b200.sharding), executes it with the CPU oracle standing in for the kernels,
and the shards are gathered into rank 0's buffer; the result must equal the unsharded build byte
for byte (terrain bands with halo; polygon ranges at global output offsets)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from myrenderer_b200 import sharding
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- terrain: row bands with a one-row halo ----
        n = 97
        sh = sharding.plan_terrain(n, rank, world)
        lo, hi = sh.halo_rows
        band = O.synth_heightmap_u16(0x5EED0004, n, lo, hi - lo)  # only the band + halo is materialised
        v, i = O.terrain_build(band, n, rows=sh.rows, qrows=sh.qrows, height_row0=lo)
        shards = [sharding.plan_terrain(n, r, world) for r in range(world)]
        gv = sharding.gather_to_rank0(torch.from_numpy(v), [s.vertex_bytes for s in shards])
        gi = sharding.gather_to_rank0(torch.from_numpy(i.view(np.int32)), [s.index_count for s in shards])
        # ---- polygons: cost-balanced ranges, global offsets ----
        seed = 0x5EED0003
        fp = O.synth_polygon_sizes(seed, 600, 8, 64)
        ft = O.polygon_offsets(fp)
        ps = sharding.plan_polygons(fp, ft, rank, world)
        sub_fp = np.ascontiguousarray(fp[ps.begin:ps.end + 1])
        xy = O.synth_polygons(seed, sub_fp, poly_index0=ps.begin)  # this rank generates only its polygons
        r = O.polygon_batch(xy, sub_fp, seed=seed, poly_index0=ps.begin, want_ids=False)
        pshards = [sharding.plan_polygons(fp, ft, k, world) for k in range(world)]
        gp = sharding.gather_to_rank0(torch.from_numpy(r["vtx"]), [(s.tri_range[1] - s.tri_range[0]) * 96 for s in pshards])
        gs = sharding.gather_to_rank0(torch.from_numpy(r["status"].view(np.int32)), [s.end - s.begin for s in pshards])
        if rank == 0:
            wv, wi = O.terrain_build(O.synth_heightmap_u16(0x5EED0004, n), n)
            whole = O.polygon_batch(O.synth_polygons(seed, fp), fp, seed=seed, want_ids=False)
            ok = (np.array_equal(gv.numpy(), wv) and np.array_equal(gi.numpy().view(np.uint32), wi)
                  and np.array_equal(gp.numpy(), whole["vtx"]) and np.array_equal(gs.numpy().view(np.uint32), whole["status"]))
            q.put(("ok" if ok else "mismatch", [s.rows for s in shards], [(s.begin, s.end) for s in pshards]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_equals_unsharded_gloo(world):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    verdict, rows, pranges = q.get(timeout=5)
    assert verdict == "ok"
    assert rows[0][0] == 0 and rows[-1][1] == 97 and pranges[0][0] == 0 and pranges[-1][1] == 600
