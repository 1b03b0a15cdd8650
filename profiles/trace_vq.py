#!/usr/bin/env python
"""Reads the per-CTA phase timestamps of a -DCGIC_VQ_TRACE build (CGIC_B200_LIB=build/variants/lib_trace.so)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, workload
import cgic_b200 as cg
B, H, W, c, m = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
dev = torch.device("cuda", 0)
cbk, counts = workload.codebook_and_counts()
cb = cbk.to(dev)
e16, e8 = workload.entropy_maps(B, H, W, 1000)
mc, mm, mf, _, mode = cg.ops.router(e16.to(dev), e8.to(dev), c, m, per_image=True)
hc, hm, hf = (t.to(dev) for t in workload.heads(B, H, W, cbk, 1000))
z = cg.ops.mask_mix(hc, hm, hf, mc, mm, mf)
flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
pc = cg.ops.Codebook(cb)
for _ in range(3):
    flush.zero_()
    cg.ops.vq_assign(z, pc)
torch.cuda.synchronize()
ws = list(cg.ops._ws_cache.values())[0]
raw = ws[256:].view(torch.int64)[512:512 + 300 * 8].cpu().numpy().reshape(300, 8)
raw = raw[raw[:, 0] > 0]
smid = raw[:, 7]
st = raw[:, :7].astype(np.float64)
t0 = st[:, 0].min()
st = (st - t0) / 1e3
names = ["start", "pdl_wait", "loaded", "classified", "searched", "finalized", "end"]
print("per-CTA phase END times (us since first CTA start): min / median / max over the CTAs")
for k, n in enumerate(names):
    print(f"  {n:11s} {st[:,k].min():7.2f} {np.median(st[:,k]):7.2f} {st[:,k].max():7.2f}")
d = np.diff(st, axis=1)
print("phase durations (us): median / max")
for k, n in enumerate(names[1:]):
    print(f"  {n:11s} {np.median(d[:,k]):7.2f} {d[:,k].max():7.2f}")

import collections
per_sm = collections.Counter(smid.tolist())
print("CTAs per SM histogram:", collections.Counter(per_sm.values()), "distinct SMs:", len(per_sm))
srch = d[:, 3]
for n_on_sm in sorted(set(per_sm.values())):
    sel = np.array([per_sm[s_] == n_on_sm for s_ in smid.tolist()])
    print(f"  SMs with {n_on_sm} CTA(s): search us min/median/max = {srch[sel].min():.1f} {np.median(srch[sel]):.1f} {srch[sel].max():.1f}; end max {st[sel,6].max():.1f}")
print("search duration deciles:", np.percentile(srch, [0, 10, 25, 50, 75, 90, 100]).round(1))
