#!/usr/bin/env python
"""Runs the bench step (encode + decode of one batch) a few times, eagerly, for use under ncu:
    ncu --set full --import-source on -k regex:<kernel> -s <skip> -c <n> -o gpurun_out/x python profiles/prof_step.py [workload] [steps]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, workload
import cgic_b200 as cg

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(steps):
    flush.zero_()
    idx, zq, sq = cg.ops.vq_assign(z, prepared)
    packed, sizes = cg.ops.pack(idx, mc, mm, mf, mode, table, h, w)
    out = cg.ops.unpack(packed, sizes, mode, table, cb, h, w)
torch.cuda.synchronize()
assert torch.equal(out[3].view(-1), idx) and int(out[5].abs().sum()) == 0
print("ok", name, steps)
