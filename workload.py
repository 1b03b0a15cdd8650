"""Seeded synthetic inputs of the hot path, shared by bench.py, smoke() and tests.

codebook ~ U(-1/1024, 1/1024) [1024,4] (= the reference's init, quantize.py:26); counters
floor(-ln(U)*1000) (KAT5 recipe); entropy maps ~ U[0,1); latent heads = codebook[randint] +
1e-4*N(0,1) (absolute: about a third of the mean spacing between neighbouring codes, so a latent
is often nearer to another code than to the one it was drawn around, and ~16 % of the latents
fall outside the codebook's bounding box -- still inside the search index's padded grid) mixed per granularity mask exactly like
vqvae_blocks.py:364-366, so z has the coarse / medium block structure of a real encoder output.  Everything comes from CPU generators
with fixed seeds, so every rank / run / box sees the same numbers.
"""
from __future__ import annotations

import numpy as np
import torch

K = 1024


def codebook_and_counts(seed: int = 1234):
    g = torch.Generator().manual_seed(seed)
    counts = (-torch.log(torch.rand(K, generator=g)) * 1000).floor().to(torch.int64)
    codebook = (torch.rand(K, 4, generator=g) * 2 - 1) / K
    return codebook.contiguous(), counts


def lexicographic_order(k: int = K):
    """Iteration order of the model's counter ParameterDict (sorted decimal strings)."""
    return sorted(range(k), key=str)


def _image_generator(seed: int, image: int) -> torch.Generator:
    return torch.Generator().manual_seed(seed * 1_000_003 + image)


def entropy_maps(B: int, H: int, W: int, seed: int, first_image: int = 0):
    """(e16 [B,H/16,W/16], e8 [B,H/8,W/8]); image i depends on (seed, first_image + i) only, so any
    sub-batch (a rank's shard, the CPU baseline's sample) sees the same images."""
    e16, e8 = [], []
    for i in range(B):
        g = _image_generator(seed, first_image + i)
        e16.append(torch.rand(H // 16, W // 16, generator=g))
        e8.append(torch.rand(H // 8, W // 8, generator=g))
    return torch.stack(e16), torch.stack(e8)


def heads(B: int, H: int, W: int, codebook: torch.Tensor, seed: int, first_image: int = 0):
    """Three latent heads [B,4,H/16,W/16], [B,4,H/8,W/8], [B,4,H/4,W/4] near codebook entries."""
    out = [[], [], []]
    for i in range(B):
        g = _image_generator(seed + 1, first_image + i)
        for lvl, div in enumerate((16, 8, 4)):
            h, w = H // div, W // div
            idx = torch.randint(0, codebook.shape[0], (h * w,), generator=g)
            v = codebook[idx] + 1e-4 * torch.randn(h * w, 4, generator=g)          # SURVEY.md 8(d), literally
            out[lvl].append(v.view(h, w, 4).permute(2, 0, 1))
    return [torch.stack(v).contiguous() for v in out]


def mix(hc, hm, hf, mc, mm, mf):
    """vqvae_blocks.py:364-366 in torch (setup code for benches; works on any device)."""
    up = lambda t, r: t.repeat_interleave(r, -1).repeat_interleave(r, -2)
    return up(hc, 4) * up(mc.float(), 4) + up(hm, 2) * up(mm.float(), 2) + hf * mf


def expected_counts(H: int, W: int, c: float, m: float):
    """(n_c, n_m, n_f) of SURVEY.md 8 for distinct entropies in mode 0."""
    n16, n8, n4 = (H // 16) * (W // 16), (H // 8) * (W // 8), (H // 4) * (W // 4)
    k_c = round(n16 * c)
    n_c = max(k_c - 1, 0)
    k_m = round(4 * n16 * c + n8 * m)
    n_m = k_m - 1 - 4 * n_c
    return n_c, n_m, n4 - 16 * n_c - 4 * n_m
