#!/usr/bin/env python
"""Dumps the captured step graph (cudaGraphDebugDotPrint via torch) to check that the PDL launches became programmatic edges."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, workload
import cgic_b200 as cg
B, H, W, c, m = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
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
def step():
    idx, zq, sq = cg.ops.vq_assign(z, prepared)
    packed, sizes = cg.ops.pack(idx, mc, mm, mf, mode, table, h, w)
    return cg.ops.unpack(packed, sizes, mode, table, cb, h, w)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
g.enable_debug_mode()
with torch.cuda.graph(g):
    out = step()
g.debug_dump(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_graph.dot")
print("dumped")
