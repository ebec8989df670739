"""Pinned host->device copy bandwidth of the box (ceiling of the e2e leg)."""
import time, torch
n = 448 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"H2D pinned {n/1e6:.0f} MB: {dt*1e3:.2f} ms = {n/dt/1e9:.1f} GB/s")
# 8 chunks of 56 MB on a side stream
s = torch.cuda.Stream()
t0 = time.perf_counter()
with torch.cuda.stream(s):
    for i in range(8):
        d[i * 56_000_000:(i + 1) * 56_000_000].copy_(h[i * 56_000_000:(i + 1) * 56_000_000], non_blocking=True)
s.synchronize()
dt = time.perf_counter() - t0
print(f"H2D 8 x 56 MB chunks: {dt*1e3:.2f} ms = {n/dt/1e9:.1f} GB/s")
