#!/usr/bin/env python
"""Phase stamps of the step's kernels for ANY workload (a -DCGIC_TRACE -DCGIC_VQ_TRACE build), without the CTA -> (image,
stream) mapping of trace_graph.py: per kernel and stamp index, min / median / max over the CTAs that wrote it, in us from
the first VQ CTA's start.   CGIC_B200_LIB=build/variants/lib_trace.so python profiles/trace_phases.py [workload]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, workload
import cgic_b200 as cg
wl = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
B, H, W, c, m = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
hp = bench.HotPath(dev)
z, masks, mode = hp.inputs(B, H, W, c, m, 0)
step = hp.step_fn([(z, masks, mode)])
run, g, _ = bench.capture(torch, step, False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(5):
    flush.zero_(); ev[0].record(); run(); ev[1].record()
torch.cuda.synchronize()
print(wl, "step (events):", round(1e3 * ev[0].elapsed_time(ev[1]), 2), "us")
best = None
for k, v in cg.ops._ws_cache.items():
    if k[0][0] not in ("vq", "encode"): continue
    raw = v[256:].view(torch.int64)[512:512 + 300 * 8].cpu().numpy().reshape(300, 8)
    raw = raw[raw[:, 0] > 0]
    if len(raw) and (best is None or raw[:, 0].max() > best[:, 0].max()): best = raw
t0 = float(best[:, 0].min())
def show(name, rows):
    rows = rows.astype(np.float64)
    for k in range(rows.shape[1]):
        col = rows[:, k]; col = col[col > t0 - 1e6]
        if len(col): print(f"  {name:9s} stamp {k:2d}: n {len(col):4d}  min {(col.min()-t0)/1e3:7.2f}  median {(np.median(col)-t0)/1e3:7.2f}  max {(col.max()-t0)/1e3:7.2f}")
show("vq", best)
for name in ("pack", "unpack", "assemble"):
    b_ = np.zeros(1024 * 16, np.uint64)
    assert getattr(ctypes.CDLL(cg._lib.LIB_PATH), "cgic_trace_" + name)(b_.ctypes.data_as(ctypes.c_void_p)) == 0
    r = b_.reshape(1024, 16); show(name, r[r[:, 0] > 0])
if "--ctas" in sys.argv:  # per-CTA view of the decode kernel: linear block id, start, end (us), "busy" = has a chunk stamp
    b_ = np.zeros(1024 * 16, np.uint64)
    ctypes.CDLL(cg._lib.LIB_PATH).cgic_trace_unpack(b_.ctypes.data_as(ctypes.c_void_p))
    r = b_.reshape(1024, 16).astype(np.float64)
    k0 = r[:, 0][r[:, 0] > 0].min()
    for lin in range(0, 1024, 8):
        row = r[lin]
        if row[0] <= 0: continue
        busy = row[2] >= row[0]
        print(f"  cta {lin:4d}  start {(row[0]-k0)/1e3:6.2f}  end {((row[6] if row[6] >= row[0] else row[0])-k0)/1e3:6.2f}  {'busy' if busy else ''}")
if "--pack-ctas" in sys.argv:  # per-CTA view of the packer: linear block id and its stamps relative to the first start
    b_ = np.zeros(1024 * 16, np.uint64)
    ctypes.CDLL(cg._lib.LIB_PATH).cgic_trace_pack(b_.ctypes.data_as(ctypes.c_void_p))
    r = b_.reshape(1024, 16).astype(np.float64)
    k0 = r[:, 0][r[:, 0] > 0].min()
    for lin in range(1024):
        row = r[lin]
        if row[0] <= 0: continue
        print(f"  pack cta {lin:4d} " + " ".join(f"{k}:{(row[k]-k0)/1e3:5.2f}" if row[k] >= row[0] else f"{k}:  -  " for k in (0, 2, 3, 4, 5, 6, 7, 1)))
if "--route" in sys.argv:  # routing tail of the fused entropy kernel, per image (needs an image-in run: profiles/bench_image_in.py style)
    x = workload.images(B, H, W, 1000).to(dev) if hasattr(workload, "images") else torch.rand(B, 3, H, W, device=dev)
    for _ in range(3):
        cg.ops.entropy_route(x, c, m)
    torch.cuda.synchronize()
    b_ = np.zeros(1024 * 8, np.uint64)
    ctypes.CDLL(cg._lib.LIB_PATH).cgic_trace_route(b_.ctypes.data_as(ctypes.c_void_p))
    r = b_.reshape(1024, 8)[:B].astype(np.float64)
    k0 = r[:, 0].min()
    for k in range(6):
        print(f"  route stamp {k}: min {(r[:,k].min()-k0)/1e3:6.2f} median {(np.median(r[:,k])-k0)/1e3:6.2f} max {(r[:,k].max()-k0)/1e3:6.2f}")
    d = r[:, 1:6] - r[:, 0:5]
    print("  per image, median us per phase (start, coarse select, coarse mask, medium select, medium mask):", np.round(np.median(d, 0) / 1e3, 2))
