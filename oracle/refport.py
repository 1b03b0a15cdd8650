"""Reference-shaped CPU port of the hot path (Python + torch CPU ops + pure-Python bit strings).

TEST INFRASTRUCTURE ONLY.  This is the "port" that bench.py times as `cpu_baseline` and as the
`--impl reference` arm: it keeps the reference's own cost structure (torch CPU kernels for the
distance matrix, per-symbol Python string concatenation for the coder, per-bit dictionary
lookups for the decoder, files on disk between pack and unpack) so the number is what a user
of the reference sees on the host cores.  The fast, definitional restatement is
cgic_oracle.c; tests cross-check the two on small cases.

The real reference cannot travel to the GPU box (/root/reference is absent there), hence a
port; tests/golden/make_golden.py checks this port against the imported reference here.
"""
from __future__ import annotations

import heapq
import os

import torch


# a1  quantize.py:69-98 -- same op sequence on torch CPU tensors
def vq_forward(z_nchw: torch.Tensor, codebook: torch.Tensor, beta: float = 0.25):
    zl = z_nchw.permute(0, 2, 3, 1).contiguous()
    flat = zl.view(-1, codebook.shape[1])
    dist = (flat ** 2).sum(dim=1, keepdim=True) + (codebook ** 2).sum(dim=1) \
        - 2 * torch.einsum("bd,dn->bn", flat, codebook.t())
    idx = torch.argmin(dist, dim=1)
    e = codebook[idx].view(zl.shape)
    loss = torch.mean((e - zl) ** 2) + beta * torch.mean((e - zl) ** 2)
    zq = (zl + (e - zl)).permute(0, 3, 1, 2).contiguous()
    return zq, loss, idx


class _Node:
    __slots__ = ("sym", "freq", "lo", "hi")

    def __init__(self, sym, freq, lo=None, hi=None):
        self.sym, self.freq, self.lo, self.hi = sym, freq, lo, hi

    def __lt__(self, other):  # indices_coding.py:26-27 -- frequency only
        return self.freq < other.freq


# a8  indices_coding.py:10-17,46-75
def huffman_codes(freq, order=None) -> dict:
    """freq[s]: int()-truncated counter of symbol s; order: symbols in the iteration order of the
    reference's `frequency` mapping (None: 0..K-1) -> {symbol: '0101..'}"""
    pq = []
    for s in (range(len(freq)) if order is None else order):
        heapq.heappush(pq, _Node(int(s), int(freq[s])))
    while len(pq) > 1:
        a = heapq.heappop(pq)
        b = heapq.heappop(pq)
        heapq.heappush(pq, _Node(None, a.freq + b.freq, a, b))
    table = {}
    todo = [(heapq.heappop(pq), "")]
    while todo:
        node, prefix = todo.pop()
        if node.sym is not None:
            table[node.sym] = prefix
        else:
            todo.append((node.hi, prefix + "1"))
            todo.append((node.lo, prefix + "0"))
    return table


def _frame(bits: str) -> bytes:
    """indices_coding.py:91-110 / mask_coding.py:22-38: pad 1..8 zero bits, 8-bit pad header."""
    pad = 8 - len(bits) % 8
    text = format(pad, "08b") + bits + "0" * pad
    return bytes(int(text[i:i + 8], 2) for i in range(0, len(text), 8))


def _unframe(data: bytes) -> str:
    """indices_coding.py:131-138,153-168: bytes -> bit string without header and padding."""
    text = ""
    for byte in data:
        text += bin(byte)[2:].rjust(8, "0")
    pad = int(text[:8], 2)
    return text[8:][:-pad]


# a9  indices_coding.py:113-126
def huff_compress(table: dict, symbols, path: str) -> str:
    with open(path, "wb") as f:
        seq = symbols.tolist()
        if seq:
            bits = ""
            for s in seq:
                bits += table[s]
            f.write(_frame(bits))
    return path


# a10 indices_coding.py:140-168
def huff_decompress(reverse: dict, path: str):
    with open(path, "rb") as f:
        data = f.read()
    if not data:
        return None
    out, cur = [], ""
    for bit in _unframe(data):
        cur += bit
        if cur in reverse:
            out.append(reverse[cur])
            cur = ""
    return out


_BIN = {0: "0", 1: "1"}
_BIN_REV = {"0": 0, "1": 1}


# a11 mask_coding.py:40-55, 81-96
def bits_compress(values, path: str) -> str:
    return huff_compress(_BIN, values, path)


def bits_decompress(path: str):
    return huff_decompress(_BIN_REV, path)


# a7, a12, a13, a14 for mode 0: model.py:217-233, 269-293, 391-392 (one image, B == 1)
def roundtrip_mode0(z_nchw, codebook, masks, table, reverse, workdir: str):
    """VQ -> select -> 5 files -> read back -> re-assemble -> gather.  Returns
    (ind [1,h,w], bpp, ind_decompress [1,h,w], quant [1,4,h,w], sizes[5])."""
    zq, loss, idx = vq_forward(z_nchw, codebook)
    h, w = zq.shape[-2:]
    ind = idx.view(-1, h, w)
    mc, mm, mf = masks
    sel = (ind[:, ::4, ::4][mc[0] == 1], ind[:, ::2, ::2][mm[0] == 1], ind[mf[0] == 1])
    names = ("indices_coarse", "indices_medium", "indices_fine", "mask_coarse", "mask_medium")
    paths = [os.path.join(workdir, n + ".bin") for n in names]
    for s in range(3):
        huff_compress(table, sel[s], paths[s])
    bits_compress(mc.flatten(), paths[3])
    bits_compress(mm.flatten(), paths[4])
    sizes = [os.path.getsize(p) for p in paths]
    bpp = sum(sizes) * 8 / (16 * h * w)

    dec = [huff_decompress(reverse, p) for p in paths[:3]]
    c = torch.tensor(bits_decompress(paths[3])).view(1, h // 4, w // 4)
    m = torch.tensor(bits_decompress(paths[4])).view(1, h // 2, w // 2)
    up = lambda t, r: t.repeat_interleave(r, dim=-1).repeat_interleave(r, dim=-2)
    f = 1 - up(m, 2) - up(c, 4)
    if dec[0] is None:
        c = torch.zeros_like(c)
    else:
        c[c == 1] = torch.tensor(dec[0])
    if dec[1] is None:
        m = torch.zeros_like(m)
    else:
        m[m == 1] = torch.tensor(dec[1])
    f[f == 1] = torch.tensor(dec[2])
    ind_dec = f + up(m, 2) + up(c, 4)
    quant = codebook[ind_dec.flatten()].view(1, h, w, -1).permute(0, 3, 1, 2)
    return ind, bpp, ind_dec, quant, sizes
