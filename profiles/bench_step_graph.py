#!/usr/bin/env python
"""The bench step replayed from a CUDA graph, device time only (event pair per step, L2 flushed between steps) --
bench.py's `value` without the e2e / CPU legs, for A/B runs of library variants (CGIC_B200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, workload
import cgic_b200 as cg
name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
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
def step():
    idx, zq, sq = cg.ops.vq_assign(z, prepared)
    packed, sizes = cg.ops.pack(idx, mc, mm, mf, mode, table, h, w)
    return idx, cg.ops.unpack(packed, sizes, mode, table, cb, h, w)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    idx, out = step()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
for _ in range(10):
    flush.zero_(); g.replay()
torch.cuda.synchronize()
for a, b in ev:
    flush.zero_(); a.record(); g.replay(); b.record()
torch.cuda.synchronize()
assert torch.equal(out[3].view(-1), idx) and int(out[5].abs().sum()) == 0
ts = sorted(a.elapsed_time(b) for a, b in ev)
print(f"{os.environ.get('CGIC_B200_LIB', 'default'):36s} {name} step mean {1e3 * sum(ts) / len(ts):.2f} us  median {1e3 * ts[len(ts) // 2]:.2f}  min {1e3 * ts[0]:.2f}")
