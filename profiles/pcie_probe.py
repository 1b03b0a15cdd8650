#!/usr/bin/env python
"""Host<->device copy bandwidth / latency of this box (pinned memory), to put the e2e number in context."""
import time, torch
dev = torch.device("cuda", 0)
for mb in (0.064, 0.25, 1, 4, 16, 64):
    n = int(mb * (1 << 20))
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(f"{name} {mb:7.3f} MiB  {dt*1e6:8.1f} us  {n/dt/1e9:6.1f} GB/s")
# both directions at once on two streams
n = 16 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); d1 = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"duplex 16 MiB each way: {dt*1e6:.1f} us  {2*n/dt/1e9:.1f} GB/s total")
