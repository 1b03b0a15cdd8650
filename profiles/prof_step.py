#!/usr/bin/env python
"""Runs the bench step (cgic_encode + cgic_unpack of one batch) a few times, eagerly, L2 flushed before each, for use under ncu:
    ncu --set full --import-source on -k regex:<kernel> -s <skip> -c <n> -o gpurun_out/x python profiles/prof_step.py [workload] [steps]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B, H, W, c, m = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
hp = bench.HotPath(dev)
z, masks, mode = hp.inputs(B, H, W, c, m, 0)
step = hp.step_fn([(z, masks, mode)])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(steps):
    flush.zero_()
    out = step()
torch.cuda.synchronize()
bench.check_roundtrip(torch, out)
print("ok", name, steps)
