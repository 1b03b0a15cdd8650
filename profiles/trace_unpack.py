#!/usr/bin/env python
"""Per-CTA phase timestamps of unpack_decode_kernel from a -DCGIC_TRACE build
(CGIC_B200_LIB=build/variants/lib_trace.so python profiles/trace_unpack.py)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, workload
import cgic_b200 as cg
B, H, W, c, m = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD]
if len(sys.argv) > 2:
    B = int(sys.argv[2])
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
flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
for _ in range(3):
    idx, zq, sq = cg.ops.vq_assign(z, prepared)
    packed, sizes = cg.ops.pack(idx, mc, mm, mf, mode, table, h, w)
    flush.zero_()
    out = cg.ops.unpack(packed, sizes, mode, table, cb, h, w)
torch.cuda.synchronize()
buf = np.zeros(1024 * 8, np.uint64)
rc = cg._lib.lib()._handle and ctypes.CDLL(cg._lib.LIB_PATH).cgic_trace_unpack(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
st = buf.reshape(1024, 8)[: 4 * B].astype(np.float64)
t0 = st[:, 0][st[:, 0] > 0].min()
names = ["start", "tables", "chunk", "staged", "A done", "B done", "end", "A0 done"]
# linear block id = k * B + b with k -> stream (medium, fine, coarse, masks): see unpack_decode_cta
for s_id, label in ((2, "coarse idx"), (0, "medium idx"), (1, "fine idx"), (3, "mask CTA")):
    rows = st[s_id * B:(s_id + 1) * B]
    print(label)
    if label != 'mask CTA':
        rel = (rows[:, :8] - t0) / 1e3
        for k, n in enumerate(names):
            print(f"   {n:8s} min/median/max {rel[:,k].min():7.2f} {np.median(rel[:,k]):7.2f} {rel[:,k].max():7.2f}")
    else:
        print(f"   start min/median/max {((rows[:,0]-t0)/1e3).min():7.2f} {np.median((rows[:,0]-t0)/1e3):7.2f} {((rows[:,0]-t0)/1e3).max():7.2f}")
