#!/usr/bin/env python
"""Timeline of the bench step INSIDE the CUDA graph (a -DCGIC_TRACE -DCGIC_VQ_TRACE build: %globaltimer stamps of the VQ
and decode CTAs): CGIC_B200_LIB=build/variants/lib_trace.so python profiles/trace_graph.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, workload
import cgic_b200 as cg
B, H, W, c, m = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
dev = torch.device("cuda", 0)
hp = bench.HotPath(dev)
z, masks, mode = hp.inputs(B, H, W, c, m, 0)
step = hp.step_fn([(z, masks, mode)])
use_graph = "--eager" not in sys.argv
run, g, _ = bench.capture(torch, step, not use_graph)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(5):
    flush.zero_(); ev[0].record(); run(); ev[1].record()
torch.cuda.synchronize()
print("step (events):", round(1e3 * ev[0].elapsed_time(ev[1]), 2), "us", "graph" if use_graph else "eager")
buf = np.zeros(1024 * 16, np.uint64)
assert ctypes.CDLL(cg._lib.LIB_PATH).cgic_trace_unpack(buf.ctypes.data_as(ctypes.c_void_p)) == 0
SL = 16
un = buf.reshape(1024, SL)[: 4 * B].astype(np.float64)
# the VQ stamps live in the vq workspace of the stream the kernel ran on: take every cached one and keep the latest stamps
best = None
for k, v in cg.ops._ws_cache.items():
    if k[0][0] not in ("vq", "encode"): continue
    raw = v[256:].view(torch.int64)[512:512 + 300 * 8].cpu().numpy().reshape(300, 8)
    raw = raw[raw[:, 0] > 0]
    if len(raw) and (best is None or raw[:, 0].max() > best[:, 0].max()): best = raw
vq = best[:, :7].astype(np.float64)
t0 = vq[:, 0].min()
rel = lambda a: (a - t0) / 1e3
print(f"vq     first start 0.00, last start {rel(vq[:,0].max()):.2f}, inputs visible (pdl wait) median {rel(np.median(vq[:,1])):.2f}, tiles done median {rel(np.median(vq[:,5])):.2f}, end median {rel(np.median(vq[:,6])):.2f} max {rel(vq[:,6].max()):.2f}")
names = {0: "medium", 1: "fine", 2: "coarse", 3: "masks"}
for k in range(4):
    rows = un[k * B:(k + 1) * B]
    s, e = rows[:, 0], rows[:, 6] if k < 3 else rows[:, 0]
    extra = ""
    if k < 3:
        extra = (f" | tables {rel(np.median(rows[:,1])):.2f} chunk {rel(np.median(rows[:,2])):.2f} staged {rel(np.median(rows[:,3])):.2f} A0 {rel(np.median(rows[:,7])):.2f}"
                 f" A {rel(np.median(rows[:,4])):.2f} B {rel(np.median(rows[:,5])):.2f} end median {rel(np.median(e)):.2f} max {rel(e.max()):.2f}")
    print(f"decode {names[k]:6s} start min {rel(s.min()):.2f} median {rel(np.median(s)):.2f} max {rel(s.max()):.2f}{extra}")

def unit(name, n):
    b_ = np.zeros(1024 * 16, np.uint64)
    assert getattr(ctypes.CDLL(cg._lib.LIB_PATH), "cgic_trace_" + name)(b_.ctypes.data_as(ctypes.c_void_p)) == 0
    return b_.reshape(1024, 16)[:n].astype(np.float64)
pk = unit("pack", 4 * B)
for k, nm in enumerate(("fine", "medium", "coarse", "masks")):
    r = pk[k * B:(k + 1) * B]
    print(f"pack {nm:13s} start median {rel(np.median(r[:,0])):.2f} max {rel(r[:,0].max()):.2f} | released {rel(np.median(r[:,2])):.2f} | stream done median {rel(np.median(r[:,1])):.2f} max {rel(r[:,1].max()):.2f}")
asm_ = unit("assemble", 4 * B)
print(f"assemble start min {rel(asm_[:,0].min()):.2f} median {rel(np.median(asm_[:,0])):.2f} max {rel(asm_[:,0].max()):.2f} | released median {rel(np.median(asm_[:,1])):.2f} max {rel(asm_[:,1].max()):.2f} | thread-0 done median {rel(np.median(asm_[:,2])):.2f} max {rel(asm_[:,2].max()):.2f}")
