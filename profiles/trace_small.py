#!/usr/bin/env python
"""Per-CTA phase timestamps of the fused small-grid kernels (encode_small_kernel, unpack_small_kernel) from a
-DCGIC_TRACE build, inside the CUDA graph of the bench step:
    CGIC_B200_LIB=build/variants/lib_trace.so python profiles/trace_small.py [images]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import cgic_b200 as cg
B, H, W, c, m = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
if len(sys.argv) > 1:
    B = int(sys.argv[1])
dev = torch.device("cuda", 0)
hp = bench.HotPath(dev)
z, masks, mode = hp.inputs(B, H, W, c, m, 0)
step = hp.step_fn([(z, masks, mode)])
runner, graph, _ = bench.capture(torch, step, "--eager" in sys.argv)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(5):
    if "--no-flush" not in sys.argv:
        flush.zero_()
    ev[0].record(); runner(); ev[1].record()
torch.cuda.synchronize()
print("step (events):", round(1e3 * ev[0].elapsed_time(ev[1]), 2), "us")
SL = 16
def unit(name):
    b_ = np.zeros(1024 * SL, np.uint64)
    assert getattr(ctypes.CDLL(cg._lib.LIB_PATH), "cgic_trace_" + name)(b_.ctypes.data_as(ctypes.c_void_p)) == 0
    return b_.reshape(1024, SL)[:min(B, 1024)].astype(np.float64)
pk, un = unit("pack"), unit("unpack")
t0 = pk[:, 0].min()
rel = lambda a: (a - t0) / 1e3
def show(title, arr, names):
    print(title)
    for k, n in names:
        col = arr[:, k]
        col = col[col > 0]
        if len(col):
            print(f"   {n:28s} min {rel(col.min()):7.2f}  median {rel(np.median(col)):7.2f}  max {rel(col.max()):7.2f}")
show("encode_small_kernel", pk, [(0, "start"), (2, "tables issued (pdl_wait done)"), (7, "VQ tiles done (per thread 0)"), (3, "barrier after VQ"), (4, "scan done"),
                                 (5, "codes staged"), (6, "streams stored"), (1, "mask streams stored")])
show("unpack_small_kernel", un, [(0, "start"), (1, "pdl_wait done"), (2, "mask levels + tables"), (3, "words staged (last batch)"), (7, "A0 len8 done"),
                                 (4, "DP done"), (10, "B1 done"), (11, "B2 done"), (5, "B3 done"), (8, "C done"), (6, "batches done"), (9, "assembled")])
