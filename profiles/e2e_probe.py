#!/usr/bin/env python
"""e2e variants of the host-buffer round trip (wall clock per call): roundtrip_host vs the pinned-arena
call with 1..8 image ranges, graph replay or eager (CGIC_SESSION_NO_GRAPH=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, workload
import cgic_b200 as cg
B, H, W, c, m = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
h, w = H // 4, W // 4
dev = torch.device("cuda", 0)
cbk, counts = workload.codebook_and_counts()
table = cg.ops.HuffTable(counts.numpy(), workload.lexicographic_order()).upload()
e16, e8 = workload.entropy_maps(B, H, W, 1000)
mc, mm, mf, _, mode = cg.ops.router(e16.to(dev), e8.to(dev), c, m, per_image=True)
hc, hm, hf = (t.to(dev) for t in workload.heads(B, H, W, cbk, 1000))
z = cg.ops.mask_mix(hc, hm, hf, mc, mm, mf)
sess = cg.ops.Session(B, h, w, mode, table, cbk)
zh = z.cpu().pin_memory(); mh = [t.cpu().pin_memory() for t in (mc, mm, mf)]
def timeit(fn, n=200):
    for _ in range(10): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return 1e6 * (time.perf_counter() - t0) / n
parts_list = tuple(int(x) for x in sys.argv[1].split(",")) if len(sys.argv) > 1 else (1, 2, 4, 8)
if len(sys.argv) <= 1:
    print("roundtrip_host us:", round(timeit(lambda: sess.roundtrip(zh, *mh)), 1))
for parts in parts_list:
    views = sess.arena(parts)
    for v in views:
        r = v["images"]; v["z"].copy_(zh[r.start:r.stop])
        for name, src in zip(("m_c", "m_m", "m_f"), mh): v[name].copy_(src[r.start:r.stop])
        for name, src in zip(("m_c8", "m_m8", "m_f8"), mh): v[name].copy_(src[r.start:r.stop].to(torch.uint8))
    print("arena parts", parts, "graph" if not os.environ.get("CGIC_SESSION_NO_GRAPH") else "eager", "us:", round(timeit(sess.roundtrip_arena), 1),
          "| narrow wire us:", round(timeit(lambda: sess.roundtrip_arena(narrow=True)), 1),
          "| decoded tensors left on the device us:", round(timeit(lambda: sess.roundtrip_arena(decoded_on_device=True)), 1),
          "| + byte masks in us:", round(timeit(lambda: sess.roundtrip_arena(decoded_on_device=True, narrow=True)), 1))
# two round trips in flight: alternate the two arena sets (submit / wait)
for parts in parts_list:
    sets = [sess.arena(parts, slot=k) for k in (0, 1)]
    for views in sets:
        for v in views:
            r = v["images"]; v["z"].copy_(zh[r.start:r.stop])
            for name, src in zip(("m_c", "m_m", "m_f"), mh): v[name].copy_(src[r.start:r.stop])
            for name, src in zip(("m_c8", "m_m8", "m_f8"), mh): v[name].copy_(src[r.start:r.stop].to(torch.uint8))
    def piped(n, **kw):
        sess.submit_arena(0, **kw)
        for i in range(1, n):
            sess.submit_arena(i & 1, **kw)
            sess.wait_arena((i - 1) & 1)
        sess.wait_arena((n - 1) & 1)
    out = []
    for kw in ({}, dict(narrow=True), dict(decoded_on_device=True), dict(decoded_on_device=True, narrow=True)):
        piped(20, **kw)
        t0 = time.perf_counter(); piped(400, **kw); out.append(round(1e6 * (time.perf_counter() - t0) / 400, 1))
    print("two in flight, parts", parts, "us per round trip: full", out[0], "| narrow wire", out[1], "| decoded on device", out[2], "| + byte masks in", out[3])
# raw copies of the arena-sized buffers through torch for reference
a = torch.empty(5570560, dtype=torch.uint8).pin_memory(); d = torch.empty(9878016, dtype=torch.uint8, device=dev); b = torch.empty(9878016, dtype=torch.uint8).pin_memory(); da = torch.empty_like(a, device=dev)
def cp():
    da.copy_(a, non_blocking=True); b.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("torch H2D 5.57MB + D2H 9.88MB serial us:", round(timeit(cp), 1))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def cp2():
    with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2): b.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
print("torch H2D 5.57MB + D2H 9.88MB on two streams us:", round(timeit(cp2), 1))
def h2d():
    da.copy_(a, non_blocking=True); torch.cuda.synchronize()
def d2h():
    b.copy_(d, non_blocking=True); torch.cuda.synchronize()
print("torch H2D 5.57MB alone us:", round(timeit(h2d), 1), "| D2H 9.88MB alone us:", round(timeit(d2h), 1))
