"""Synthetic polygon workloads built with torch on the GPU (benchmark / test plumbing, not the product).

`ellipse_batch`: convex polygons -- the family the reference triangulator handles correctly at every
size: vertex k of an n-gon at angle 2*pi*(k + jitter)/n on a randomly rotated
ellipse centred at (100,100), positive shoelace area in raw (x, y).  The star-shaped family of SURVEY
8-d lives in the library itself (mr_synth_polygons)."""
from __future__ import annotations

import numpy as np


def ellipse_batch(first_point: np.ndarray, seed: int, device="cuda"):
    """Returns (xy float32 [npts,2] on `device`, polygon id per point)."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    fp = torch.from_numpy(first_point.astype(np.int64)).to(device)
    nper = fp[1:] - fp[:-1]
    npoly, tot = len(nper), int(fp[-1] - fp[0])
    pid = torch.repeat_interleave(torch.arange(npoly, device=device), nper)
    k = torch.arange(tot, device=device) - (fp[:-1] - fp[0])[pid]
    th = 2 * np.pi * (k.double() + 0.8 * torch.rand(tot, generator=g, device=device, dtype=torch.float64) - 0.4) / nper[pid].double()
    a = (40 + 50 * torch.rand(npoly, generator=g, device=device, dtype=torch.float64))[pid]
    b = (40 + 50 * torch.rand(npoly, generator=g, device=device, dtype=torch.float64))[pid]
    ph = (6.28 * torch.rand(npoly, generator=g, device=device, dtype=torch.float64))[pid]
    x, y = a * torch.cos(th), b * torch.sin(th)
    xy = torch.stack([100 + torch.cos(ph) * x - torch.sin(ph) * y, 100 + torch.sin(ph) * x + torch.cos(ph) * y], 1)
    return xy.float().contiguous(), pid
