#!/usr/bin/env python
"""SpatialNorm (f4): the fused op against the reference's eager expression on the GPU, decoder-sized feature maps.
Algorithmic bytes: 8 B per element (f read once + new_f written; the second read of f is an L2 hit when the map fits)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgic_b200 as cg

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)

def timed(fn, iters=20):
    tot = 0.0
    for it in range(iters + 3):
        flush.zero_()
        a, b = ev(), ev()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if it >= 3: tot += a.elapsed_time(b)
    return 1e3 * tot / iters

for B, Cc, H, W, hz in [(1, 512, 64, 64, 64), (8, 512, 64, 64, 64), (8, 256, 128, 128, 64), (8, 128, 256, 256, 64), (24, 128, 512, 768, 128)]:
    if B * Cc * H * W * 4 > (6 << 30): continue
    m = cg.Normalize(Cc, 4, False).to(dev).eval()
    f = torch.randn(B, Cc, H, W, device=dev)
    zq = torch.randn(B, 4, hz, hz * W // H, device=dev)
    with torch.no_grad():
        t_fused = timed(lambda: m(f, zq))
        t_eager = timed(lambda: cg.decoder._eager(f, zq, m.norm_layer, m.conv_y, m.conv_b))
    nbytes = 8 * f.numel()
    print(f"[{B},{Cc},{H},{W}] fused {t_fused:8.1f} us = {nbytes / t_fused / 1e3:7.1f} GB/s (8 B/elem)   eager {t_eager:8.1f} us   x{t_eager / t_fused:.1f}")
