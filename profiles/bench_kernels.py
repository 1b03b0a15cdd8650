#!/usr/bin/env python
"""Per-stage device timing of the bench step (events around each C-ABI call, L2 flushed, GPU kept
busy by the flush so that launch latency is hidden).  For kernel tuning; honours CGIC_B200_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, workload
import cgic_b200 as cg

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
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
acc = {"vq": 0.0, "pack": 0.0, "unpack": 0.0}
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(iters + 5):
    flush.zero_()
    e = [ev() for _ in range(4)]
    e[0].record()
    idx, zq, sq = cg.ops.vq_assign(z, prepared)
    e[1].record()
    packed, sizes = cg.ops.pack(idx, mc, mm, mf, mode, table, h, w)
    e[2].record()
    out = cg.ops.unpack(packed, sizes, mode, table, cb, h, w)
    e[3].record()
    torch.cuda.synchronize()
    if it >= 5:
        acc["vq"] += e[0].elapsed_time(e[1]); acc["pack"] += e[1].elapsed_time(e[2]); acc["unpack"] += e[2].elapsed_time(e[3])
assert torch.equal(out[3].view(-1), idx) and int(out[5].abs().sum()) == 0
print(os.environ.get("CGIC_B200_LIB", "default"), name, {k: round(1e3 * v / iters, 2) for k, v in acc.items()}, "us")
