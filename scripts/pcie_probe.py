import torch, time
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
for name, (dst, src) in (("D2H", (h, d)), ("H2D", (d, h))):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(name, "GB/s", round(5 * (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1))
# two concurrent D2H halves on two streams
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h[: 1 << 29].copy_(d[: 1 << 29], non_blocking=True)
    with torch.cuda.stream(s2): h[1 << 29:].copy_(d[1 << 29:], non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("D2H two streams GB/s", round(5 * (1 << 30) / dt / 1e9, 1))
# bidirectional
h2 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("bidirectional each-way GB/s", round(5 * (1 << 30) / dt / 1e9, 1))
