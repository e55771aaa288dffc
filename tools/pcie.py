"""Host<->device copy bandwidth of this box (pinned memory), for reading bench.py's e2e number."""
import time
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("H2D", d, h), ("D2H", h, d)):
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    print(name, f"{5 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("bidirectional", f"{10 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s total")
