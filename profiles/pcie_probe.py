#!/usr/bin/env python
"""Host<->device copy bandwidth / latency of this box (pinned memory), to put the e2e number in context.  Under
torch.distributed.run every rank probes its own GPU at the same time (barrier before every size): what N ranks get when
they share the host's memory / PCIe root."""
import os, time, torch
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
def barrier():
    if world > 1:
        dist.barrier()
for mb in (0.25, 1, 4, 16, 64):
    n = int(mb * (1 << 20))
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); barrier(); t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        t = torch.tensor([n / dt / 1e9], device=dev)
        if world > 1:
            ts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(ts, t)
            if rank == 0:
                v = [float(x) for x in ts]
                print(f"{name} {mb:7.3f} MiB x {world} ranks at once: per rank min {min(v):6.1f} max {max(v):6.1f} GB/s, sum {sum(v):7.1f} GB/s")
        else:
            print(f"{name} {mb:7.3f} MiB  {dt*1e6:8.1f} us  {n/dt/1e9:6.1f} GB/s")
# both directions at once on two streams
n = 16 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); d1 = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); barrier(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
if rank == 0:
    print(f"duplex 16 MiB each way (rank 0 of {world}): {dt*1e6:.1f} us  {2*n/dt/1e9:.1f} GB/s total")
if world > 1:
    dist.destroy_process_group()
