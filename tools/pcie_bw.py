"""Measure pinned D2H / H2D bandwidth on this box (context for bench.py's e2e number: 9 B per ray must cross PCIe)."""
import time
import torch

n = 75 << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
hp = torch.empty(n, dtype=torch.uint8)
for name, dst, src in (("D2H pinned", h, d), ("H2D pinned", d, h), ("D2H pageable", hp, d)):
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name}: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms per 75 MiB)")
# chunked: 8 chunks x 3 arrays like ort_trace_frame
chunks = [(i * (n // 24), n // 24) for i in range(24)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    for o, l in chunks:
        h[o:o + l].copy_(d[o:o + l], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"D2H pinned in 24 pieces: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms)")
