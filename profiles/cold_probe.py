#!/usr/bin/env python
"""How much of each stage's time is cold caches (data + instructions)?  Times every stage of the bench step
(a) with the L2 flushed right before the stage, (b) inside the step (the predecessor's outputs in L2),
(c) repeated back to back (everything warm).  Honours CGIC_B200_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, workload
import cgic_b200 as cg

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
B, H, W, c, m = bench.WORKLOADS[name]
h, w = H // 4, W // 4
dev = torch.device("cuda", 0)
cbk, counts = workload.codebook_and_counts()
table = cg.ops.HuffTable(counts.numpy(), workload.lexicographic_order()).upload()
cb = cbk.to(dev)
prepared = cg.ops.Codebook(cb)
e16, e8 = workload.entropy_maps(B, H, W, 1000)
mc, mm, mf, _, mode = cg.ops.router(e16.to(dev), e8.to(dev), c, m, per_image=True)
hc, hm, hf = (t.to(dev) for t in workload.heads(B, H, W, cbk, 1000))
z = cg.ops.mask_mix(hc, hm, hf, mc, mm, mf)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
idx, zq, sq = cg.ops.vq_assign(z, prepared)
packed, sizes = cg.ops.pack(idx, mc, mm, mf, mode, table, h, w)
stages = {
    "vq": lambda: cg.ops.vq_assign(z, prepared),
    "pack": lambda: cg.ops.pack(idx, mc, mm, mf, mode, table, h, w),
    "unpack": lambda: cg.ops.unpack(packed, sizes, mode, table, cb, h, w),
}
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, pre):
    tot = 0.0
    for it in range(iters + 3):
        pre()
        a, b = ev(), ev()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if it >= 3: tot += a.elapsed_time(b)
    return round(1e3 * tot / iters, 2)
res = {}
for k, fn in stages.items():
    cold = timed(fn, lambda: flush.zero_())
    warm = timed(fn, lambda: (fn(), torch.cuda._sleep(400000)))  # same kernel just ran; a spin kernel (no memory traffic) hides the launch latency
    res[k] = {"l2_flushed_us": cold, "warm_us": warm}
print(os.environ.get("CGIC_B200_LIB", "default"), name, res)
