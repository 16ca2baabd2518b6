"""Probe: pinned host -> device copy rate of this box (the ceiling of the e2e number, which copies every frame)."""
import time, torch
n = 3 * 1024**3
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(h, non_blocking=True); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("H2D pinned 3 GiB: %.2f GB/s (best of 5)" % (n / best / 1e6))
