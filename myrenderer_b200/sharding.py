"""Sharding of the hot path over the GPUs of one box (one process per GPU).

Both halves of the path partition into independent units, so there is no data-path
collective: every rank builds its shard into its own HBM.
    terrain   row bands; rank g needs rows [r0-1, r1+1) of the heightmap (a one-row halo that is an
              *input* overlap, not an exchange) and writes vertex rows [r0,r1) and quad rows [q0,q1)
    polygons  contiguous ranges balanced by the cost model w(n) = n*log2(n)+n; output offsets are the
              global exclusive prefix sum of (n_i-2), so shards land at disjoint final positions
`gather_to_rank0` assembles the shards in rank 0's buffer where a caller wants one buffer
(torch.distributed: NCCL on GPUs, gloo in the CPU tests).  The partition arithmetic itself lives in
the C library (mr_terrain_partition / mr_polygon_partition) so a Zig host gets the same split.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _capi as capi


@dataclass(frozen=True)
class TerrainShard:
    n: int
    rows: tuple       # vertex rows [r0, r1)
    qrows: tuple      # quad rows [q0, q1)
    halo_rows: tuple  # heightmap rows needed [lo, hi)

    @property
    def vertex_bytes(self):
        return (self.rows[1] - self.rows[0]) * self.n * 32

    @property
    def index_count(self):
        return (self.qrows[1] - self.qrows[0]) * 6 * max(self.n - 1, 0)


@dataclass(frozen=True)
class PolygonShard:
    begin: int  # polygon range [begin, end)
    end: int
    point_range: tuple  # [first_point[begin], first_point[end])
    tri_range: tuple    # [first_tri[begin], first_tri[end])


def plan_terrain(n: int, rank: int, world: int) -> TerrainShard:
    lib = capi.load()
    rows = (C.c_uint32 * (world + 1))()
    qrows = (C.c_uint32 * (world + 1))()
    rc = lib.mr_terrain_partition(n, world, rows, qrows)
    if rc != 0:
        raise capi.MrError(rc, "mr_terrain_partition")
    r0, r1 = rows[rank], rows[rank + 1]
    return TerrainShard(n, (r0, r1), (qrows[rank], qrows[rank + 1]), (max(r0 - 1, 0), min(r1 + 1, n)))


def plan_polygons(first_point: np.ndarray, first_tri: np.ndarray, rank: int, world: int) -> PolygonShard:
    lib = capi.load()
    fp = np.ascontiguousarray(first_point, dtype=np.uint64)
    npoly = len(fp) - 1
    ranges = (C.c_uint32 * (world + 1))()
    rc = lib.mr_polygon_partition(fp.ctypes.data, npoly, world, ranges)
    if rc != 0:
        raise capi.MrError(rc, "mr_polygon_partition")
    a, b = ranges[rank], ranges[rank + 1]
    return PolygonShard(a, b, (int(fp[a]), int(fp[b])), (int(first_tri[a]), int(first_tri[b])))


def gather_to_rank0(shard, sizes, dst=None):
    """Concatenate every rank's 1-D tensor `shard` (lengths `sizes`, known to all ranks) into rank 0's `dst`.
    Returns dst on rank 0, None elsewhere.  Point-to-point sends: only rank 0 ingests."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    if rank == 0:
        if dst is None:
            dst = torch.empty(int(sum(sizes)), dtype=shard.dtype, device=shard.device)
        offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        dst[offs[0]:offs[1]].copy_(shard)
        reqs = [dist.irecv(dst[offs[r]:offs[r + 1]], src=r) for r in range(1, world) if sizes[r] > 0]
        for q in reqs:
            q.wait()
        return dst
    if sizes[rank] > 0:
        dist.send(shard, dst=0)
    return None
