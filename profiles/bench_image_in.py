#!/usr/bin/env python
"""Per-kernel device time of the image-in head on the bench batch, L2 flushed between iterations: the two launches of SURVEY 8f f1
(cgic_entropy_route = entropy maps + per-image routing; cgic_route_mix = fine mask + mask-mix), and with --three the three
stand-alone kernels (entropy maps, router, mask-mix) they replace.  For kernel tuning (honours CGIC_B200_LIB)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, workload
import cgic_b200 as cg
name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else bench.DEFAULT_WORKLOAD
B, H, W, c, m = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(7)
x = torch.rand(B, 3, H, W, generator=g).to(dev)
cbk, _ = workload.codebook_and_counts()
hc, hm, hf = (t.to(dev) for t in workload.heads(B, H, W, cbk, 1000))
flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
lib = cg._lib.lib()
def step():
    if "--three" in sys.argv:
        e8, e16 = cg.ops.entropy_maps(x)
        mc, mm, mf, gate, mode = cg.ops.router(e16, e8, c, m, per_image=True)
        return cg.ops.mask_mix(hc, hm, hf, mc, mm, mf)
    e8, e16, mc, mm, near, mode = cg.ops.entropy_route(x, c, m)
    return cg.ops.route_mix(hc, hm, hf, mc, mm, mode)[2]
for _ in range(3):
    step()
lib.cgic_prof_enable(1)
n = 20
for _ in range(n):
    flush.zero_()
    step()
buf = ctypes.create_string_buffer(8192)
lib.cgic_prof_report(buf, 8192)
lib.cgic_prof_enable(0)
tot = 0
for line in buf.value.decode().splitlines():
    k, cnt, ms = line.split()
    us = 1e3 * float(ms) / int(cnt)
    tot += us * int(cnt) / n
    print(f"{k:24s} {int(cnt)/n:4.1f}/step {us:8.2f} us")
print(f"image-in head: {tot:.1f} us per step; image bytes {x.numel()*4/1e6:.1f} MB -> {x.numel()*4/tot/1e3:.0f} GB/s")
