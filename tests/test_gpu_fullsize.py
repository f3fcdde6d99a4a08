"""Full BASELINE sizes on one B200 (configs 4 and 5), checked through size-independent properties
plus bit-exact comparison of sampled pieces against the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_terrain_16384_properties_and_sampled_bands(ctx, oracle):
    """Config 4's heightmap (16384^2, seed 0x5EED0004) meshed on one GPU: 8.6 GB of vertices, 6.4 GB of
    indices (byte offsets beyond 2^32).  Properties: closed-form checksum of the whole index buffer;
    x/z position planes are the closed form; normals are unit length; sampled row bands bit-exact."""
    import torch

    import myrenderer_b200 as mr

    n = 16384
    seed = 0x5EED0004
    height = torch.empty(n * n, dtype=torch.int16, device="cuda")
    ctx.check(ctx.lib.mr_synth_heightmap_u16(ctx.handle, seed, n, 0, n, height.data_ptr()), "synth")
    vtx = torch.empty(n * n * 32, dtype=torch.uint8, device="cuda")
    idx = torch.empty(6 * (n - 1) * (n - 1), dtype=torch.int32, device="cuda")
    T = mr.Terrain(ctx)
    T.build(T.job(height, n, vtx_out=vtx, idx_out=idx))
    ctx.sync()
    # ---- index checksum: sum over quads of (6*i00 + 3n + 3), i00 = r*n + c
    m = n - 1
    sr = m * (m - 1) // 2  # sum of r (or c) over 0..m-1
    want = 6 * (n * sr * m + sr * m) + (3 * n + 3) * m * m
    got = 0
    step = 1 << 28
    for a in range(0, idx.numel(), step):
        got += int(idx[a:a + step].to(torch.int64).sum().item())
    assert got == want
    assert int(idx.max().item()) == n * n - 1 and int(idx.min().item()) == 0
    # ---- positions: x depends on the row only, z on the column only, both the closed form
    V = vtx.view(torch.float32).view(n, n, 8)
    r = torch.arange(n, device="cuda", dtype=torch.float32)
    plane = (torch.tensor(0.2, device="cuda") * r - torch.tensor(0.1, device="cuda") * torch.tensor(float(n), device="cuda"))
    for rows in (slice(0, 2048), slice(7000, 9048), slice(n - 2048, n)):
        assert torch.equal(V[rows, :, 0], plane[rows, None].expand(-1, n))
        assert torch.equal(V[rows, :, 2], plane[None, :].expand(rows.stop - rows.start, -1))
        nl = (V[rows, :, 4:7].double() ** 2).sum(-1)
        assert float((nl - 1).abs().max().item()) < 1e-6
        assert not bool(V[rows, :, 3].any().item()) and not bool(V[rows, :, 7].any().item())  # pad lanes
    # ---- sampled bands, bit-exact against the oracle (which generates the same heightmap rows on the host)
    for r0, r1 in ((0, 6), (8190, 8196), (n - 6, n)):
        lo, hi = max(r0 - 1, 0), min(r1 + 1, n)
        band = oracle.synth_heightmap_u16(seed, n, lo, hi - lo)
        q0, q1 = r0, min(r1, n - 1)
        ov, oi = oracle.terrain_build(band, n, rows=(r0, r1), qrows=(q0, q1), height_row0=lo, nthreads=0)
        assert np.array_equal(vtx[r0 * n * 32:r1 * n * 32].cpu().numpy(), ov)
        assert np.array_equal(idx[q0 * 6 * m:q1 * 6 * m].cpu().numpy().view(np.uint32), oi)


def _ellipse_batch(first_point, seed):
    from myrenderer_b200.workloads import ellipse_batch

    return ellipse_batch(first_point, seed)


def test_polygons_1m_skewed_sizes(ctx, oracle):
    """Config 5 on one GPU: 1,000,000 polygons, sizes log-uniform in [8,1024] (2.1e8 points in, 20 GB of
    vertices out).  Properties over the WHOLE batch: every polygon ends OK with n-2 triangles, the signed
    areas of its triangles sum to its shoelace area (a valid triangulation), every triangle keeps the
    polygon's orientation; a random sample is compared with the oracle bit for bit."""
    import torch

    import myrenderer_b200 as mr

    npoly = 1_000_000
    seed = 0x5EED0005
    fp = oracle.synth_polygon_sizes(seed, npoly, 8, 1024, dist=1)
    npts = int(fp[-1])
    assert 1.8e8 < npts < 2.4e8
    xy, pid = _ellipse_batch(fp, 1234)
    P = mr.Polygon(ctx)
    batch = P.create_polygons(xy, fp, seed=seed)
    ctx.sync()
    status = batch.status.cpu().numpy().view(np.uint32)
    ntri = batch.ntri.cpu().numpy().view(np.uint32)
    n = np.diff(fp.astype(np.int64))
    assert (status == 0).all(), f"statuses: {np.unique(status, return_counts=True)}"
    assert np.array_equal(ntri, (n - 2).astype(np.uint32))
    # ---- area checksum over all ~2.1e8 triangles, in chunks of polygons
    ft = torch.from_numpy(batch.first_tri.astype(np.int64)).cuda()
    fpd = torch.from_numpy(fp.astype(np.int64)).cuda()
    X = xy.double()
    nxt = torch.arange(npts, device="cuda") + 1
    last = fpd[1:] - 1
    nxt[last] = fpd[:-1]  # wrap the last vertex of each polygon to its first
    cross = X[:, 0] * X[nxt, 1] - X[nxt, 0] * X[:, 1]
    poly_area = torch.zeros(npoly, device="cuda", dtype=torch.float64).index_add_(0, pid, cross) * 0.5
    del cross, nxt
    tri_area_sum = torch.zeros(npoly, device="cuda", dtype=torch.float64)
    min_tri = torch.full((1,), 1e300, device="cuda", dtype=torch.float64)
    Vb = batch.vertex_buffer
    chunk = 100_000
    for a in range(0, npoly, chunk):
        z = min(a + chunk, npoly)
        t0, t1 = int(ft[a]), int(ft[z])
        v = Vb[t0 * 96:t1 * 96].view(torch.float32).view(-1, 3, 8)[:, :, :2].double()  # x attribute of 3 vertices
        ar = 0.5 * ((v[:, 1, 0] - v[:, 0, 0]) * (v[:, 2, 1] - v[:, 0, 1]) - (v[:, 2, 0] - v[:, 0, 0]) * (v[:, 1, 1] - v[:, 0, 1]))
        owner = torch.repeat_interleave(torch.arange(a, z, device="cuda"), ft[a + 1:z + 1] - ft[a:z])
        tri_area_sum.index_add_(0, owner, ar)
        min_tri = torch.minimum(min_tri, ar.min().reshape(1))
    rel = ((tri_area_sum - poly_area).abs() / poly_area).max().item()
    assert rel < 1e-5, rel
    assert float(min_tri.item()) > -1e-3  # consistent orientation (degenerate slivers allowed)
    # ---- random sample, bit for bit
    rng = np.random.default_rng(7)
    xy_h = None
    for i in rng.choice(npoly, 600, replace=False):
        a, z = int(fp[i]), int(fp[i + 1])
        pts = xy[a:z].cpu().numpy()
        ref = oracle.polygon_batch(pts, np.array([0, z - a]), seed=seed, poly_index0=int(i), want_ids=False)
        t0, t1 = int(batch.first_tri[i]), int(batch.first_tri[i + 1])
        assert np.array_equal(Vb[t0 * 96:t1 * 96].cpu().numpy(), ref["vtx"]), f"polygon {i} (n={z - a})"
        assert np.array_equal(batch.bbox[i].cpu().numpy().view(np.uint32), ref["bbox"].view(np.uint32)[0])


def test_polygons_1m_zipper_family(ctx, oracle):
    """Config 5 with the NON-CONVEX family on which the reference is sound (MR_FAMILY_ZIPPER), generated on the device:
    all 1,000,000 polygons end OK with n-2 triangles whose areas add up to the polygon's area; a random sample --
    inputs regenerated on the CPU from the same counter hash -- is compared with the oracle bit for bit."""
    import torch

    import myrenderer_b200 as mr

    npoly = 1_000_000
    seed = 0x5EED0005
    fp = oracle.synth_polygon_sizes(seed, npoly, 8, 1024, dist=1)
    npts = int(fp[-1])
    fpd = torch.from_numpy(fp.astype(np.int64)).cuda()
    xy = torch.empty(npts * 2, dtype=torch.float32, device="cuda")
    ctx.check(ctx.lib.mr_synth_polygons_family(ctx.handle, oracle.FAMILY_ZIPPER, seed, 0, fpd.data_ptr(), npoly, xy.data_ptr()), "synth")
    xy = xy.view(-1, 2)
    P = mr.Polygon(ctx)
    batch = P.create_polygons(xy, fp, seed=seed)
    ctx.sync()
    status = batch.status.cpu().numpy().view(np.uint32)
    n = np.diff(fp.astype(np.int64))
    assert (status == 0).all(), f"statuses: {np.unique(status, return_counts=True)}"
    assert np.array_equal(batch.ntri.cpu().numpy().view(np.uint32), (n - 2).astype(np.uint32))
    # areas: sum over each polygon's triangles == shoelace area
    pid = torch.repeat_interleave(torch.arange(npoly, device="cuda"), fpd[1:] - fpd[:-1])
    X = xy.double()
    nxt = torch.arange(npts, device="cuda") + 1
    nxt[fpd[1:] - 1] = fpd[:-1]
    cross = X[:, 0] * X[nxt, 1] - X[nxt, 0] * X[:, 1]
    poly_area = torch.zeros(npoly, device="cuda", dtype=torch.float64).index_add_(0, pid, cross) * 0.5
    del cross, nxt, pid, X
    ft = torch.from_numpy(batch.first_tri.astype(np.int64)).cuda()
    tri_area_sum = torch.zeros(npoly, device="cuda", dtype=torch.float64)
    Vb = batch.vertex_buffer
    chunk = 100_000
    for a in range(0, npoly, chunk):
        z = min(a + chunk, npoly)
        t0, t1 = int(ft[a]), int(ft[z])
        v = Vb[t0 * 96:t1 * 96].view(torch.float32).view(-1, 3, 8)[:, :, :2].double()
        ar = 0.5 * ((v[:, 1, 0] - v[:, 0, 0]) * (v[:, 2, 1] - v[:, 0, 1]) - (v[:, 2, 0] - v[:, 0, 0]) * (v[:, 1, 1] - v[:, 0, 1]))
        owner = torch.repeat_interleave(torch.arange(a, z, device="cuda"), ft[a + 1:z + 1] - ft[a:z])
        tri_area_sum.index_add_(0, owner, ar.abs())
    rel = ((tri_area_sum - poly_area).abs() / poly_area).max().item()
    assert rel < 1e-5, rel  # the triangles tile the polygon
    rng = np.random.default_rng(8)
    for i in rng.choice(npoly, 400, replace=False):
        a, z = int(fp[i]), int(fp[i + 1])
        pts = oracle.synth_polygons(seed, np.array([0, z - a], dtype=np.uint64), poly_index0=int(i), family=oracle.FAMILY_ZIPPER)
        assert np.array_equal(pts, xy[a:z].cpu().numpy()), f"generator differs at polygon {i}"
        ref = oracle.polygon_batch(pts, np.array([0, z - a]), seed=seed, poly_index0=int(i), want_ids=False)
        t0, t1 = int(batch.first_tri[i]), int(batch.first_tri[i + 1])
        assert np.array_equal(Vb[t0 * 96:t1 * 96].cpu().numpy(), ref["vtx"]), f"polygon {i} (n={z - a})"
