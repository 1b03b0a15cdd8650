"""ctypes binding of the C restatement (oracle/cgic_oracle.c) with numpy in/out.

TEST INFRASTRUCTURE ONLY -- the checker, never the product.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
Nothing under control-gic_b200/ does.

Parity pinning: see the header of cgic_oracle.c -- pinned against fixtures generated from the
imported reference (tests/golden/make_golden.py) and the KATs of SURVEY.md 8(c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcgic_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "cgic_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_huff_encode.restype = C.c_int64
        _lib.orc_huff_decode.restype = C.c_int64
        _lib.orc_bits_encode.restype = C.c_int64
        _lib.orc_bits_decode.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


# ---------------------------------------------------------------------------------------------
# a1  VectorQuantize2.forward (quantize.py:69-98)
# ---------------------------------------------------------------------------------------------
def vq_assign(z_nchw: np.ndarray, codebook: np.ndarray, beta: float = 0.25):
    """-> (z_q NCHW f32, loss f32, idx int64 flat [B*h*w]) for e_dim == 4."""
    z = _f32(z_nchw)
    cb = _f32(codebook)
    B, c, h, w = z.shape
    assert c == 4 and cb.shape[1] == 4
    idx = np.empty(B * h * w, np.int64)
    zq = np.empty_like(z)
    sq = C.c_double(0.0)
    rc = lib().orc_vq_assign(_p(z, C.c_float), B, h, w, _p(cb, C.c_float), cb.shape[0],
                             _p(idx, C.c_int64), _p(zq, C.c_float), C.byref(sq))
    assert rc == 0, rc
    m = np.float32(sq.value / z.size)
    loss = np.float32(m + np.float32(beta) * m)
    return zq, loss, idx


# ---------------------------------------------------------------------------------------------
# a8  HuffmanCoding table (indices_coding.py:10-17,46-75)
# ---------------------------------------------------------------------------------------------
@dataclass
class HuffTable:
    K: int
    lengths: np.ndarray  # int32 [K]
    codes_buf: np.ndarray  # uint8 [K, K] '0'/'1' chars, NUL terminated
    left: np.ndarray
    right: np.ndarray
    root: int

    def code(self, s: int) -> str:
        return bytes(self.codes_buf[s, : self.lengths[s]]).decode()

    @property
    def codes(self) -> dict:
        return {s: self.code(s) for s in range(self.K)}

    @property
    def max_len(self) -> int:
        return int(self.lengths.max())


def lexicographic_order(K: int) -> np.ndarray:
    """Iteration order of the reference model's counter ParameterDict: torch builds it from a
    plain dict with sorted(items()), i.e. decimal key strings in lexicographic order
    (quantize.py:28, consumed by indices_coding.py:46-49 via inference.py:137-139)."""
    return np.asarray(sorted(range(K), key=str), np.int32)


def huff_build(freq, order=None) -> HuffTable:
    """freq[s] = int count of symbol s; order[i] = symbol pushed i-th (None: 0..K-1)."""
    f = _i64(freq)
    K = f.shape[0]
    order_p = None if order is None else _p(_i32(order), C.c_int32)
    lengths = np.zeros(K, np.int32)
    buf = np.zeros((K, K), np.uint8)
    left = np.zeros(2 * K, np.int32)
    right = np.zeros(2 * K, np.int32)
    root = C.c_int32(0)
    rc = lib().orc_huff_build(_p(f, C.c_int64), order_p, K, _p(lengths, C.c_int32), _p(buf, C.c_char),
                              _p(left, C.c_int32), _p(right, C.c_int32), C.byref(root))
    assert rc == 0, rc
    return HuffTable(K, lengths, buf, left, right, int(root.value))


def huff_capacity(t: HuffTable, n: int) -> int:
    return (n * t.max_len) // 8 + 2


# a9  HuffmanCoding.compress (indices_coding.py:113-126)
def huff_encode(t: HuffTable, symbols) -> bytes:
    s = _i64(symbols).ravel()
    cap = huff_capacity(t, s.size) + 8
    out = np.zeros(cap, np.uint8)
    n = lib().orc_huff_encode(_p(s, C.c_int64), C.c_int64(s.size), t.K, _p(t.lengths, C.c_int32),
                              _p(t.codes_buf, C.c_char), _p(out, C.c_uint8), C.c_int64(cap))
    assert n >= 0, n
    return out[:n].tobytes()


# a10 HuffmanCoding.decompress_string (indices_coding.py:153-168) -> list or None
def huff_decode(t: HuffTable, data: bytes):
    if len(data) == 0:
        return None
    a = np.frombuffer(data, np.uint8)
    cap = len(data) * 8 + 8
    out = np.zeros(cap, np.int64)
    n = lib().orc_huff_decode(_p(a, C.c_uint8), C.c_int64(len(data)), t.K, _p(t.left, C.c_int32),
                              _p(t.right, C.c_int32), t.root, _p(out, C.c_int64), C.c_int64(cap))
    assert n >= 0, n
    return out[:n].tolist()


# a11 BinaryCoding (mask_coding.py:40-55, 81-96)
def bits_encode(values) -> bytes:
    v = _i32(values).ravel()
    cap = v.size // 8 + 2
    out = np.zeros(cap, np.uint8)
    n = lib().orc_bits_encode(_p(v, C.c_int32), C.c_int64(v.size), _p(out, C.c_uint8), C.c_int64(cap))
    assert n >= 0, n
    return out[:n].tobytes()


def bits_decode(data: bytes):
    if len(data) == 0:
        return None
    a = np.frombuffer(data, np.uint8)
    cap = len(data) * 8
    out = np.zeros(cap, np.int64)
    n = lib().orc_bits_decode(_p(a, C.c_uint8), C.c_int64(len(data)), _p(out, C.c_int64), C.c_int64(cap))
    assert n >= 0, n
    return out[:n].tolist()


# ---------------------------------------------------------------------------------------------
# a7 + a12  selection + 5-stream pack, one image (model.py:217-260)
# ---------------------------------------------------------------------------------------------
STREAM_NAMES = ("indices_coarse", "indices_medium", "indices_fine", "mask_coarse", "mask_medium")
STREAMS_BY_MODE = ((1, 1, 1, 1, 1), (0, 1, 1, 0, 1), (1, 0, 1, 1, 0), (1, 1, 0, 1, 0),
                   (1, 0, 0, 0, 0), (0, 1, 0, 0, 0), (0, 0, 1, 0, 0))


def slot_layout(t: HuffTable, h: int, w: int):
    """Worst-case byte capacity of the five stream slots of one image, 16-byte aligned."""
    n4, n8, n16 = h * w, (h // 2) * (w // 2), (h // 4) * (w // 4)
    caps = [huff_capacity(t, n16), huff_capacity(t, n8), huff_capacity(t, n4), n16 // 8 + 2, n8 // 8 + 2]
    caps = [(c + 15) // 16 * 16 for c in caps]
    offs = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.int64)
    return offs, np.asarray(caps, np.int64), int(sum(caps))


def pack_image(t: HuffTable, ind_hw, mc, mm, mf, mode: int):
    """-> list of 5 bytes objects (b'' for absent/empty streams)."""
    ind = _i64(ind_hw)
    h, w = ind.shape
    offs, caps, total = slot_layout(t, h, w)
    out = np.zeros(total, np.uint8)
    sizes = np.zeros(5, np.int32)
    mc_, mm_, mf_ = _i32(mc).ravel(), _i32(mm).ravel(), _i32(mf).ravel()
    rc = lib().orc_pack_image(_p(ind, C.c_int64), _p(mc_, C.c_int32), _p(mm_, C.c_int32), _p(mf_, C.c_int32),
                              h, w, mode, t.K, _p(t.lengths, C.c_int32), _p(t.codes_buf, C.c_char),
                              _p(out, C.c_uint8), _p(offs, C.c_int64), _p(caps, C.c_int64), _p(sizes, C.c_int32))
    assert rc == 0, rc
    return [out[offs[s]: offs[s] + sizes[s]].tobytes() for s in range(5)]


def bpp_of(streams, H: int, W: int) -> float:
    """model.py:233 -- sum of file sizes * 8 / num_pixels, in Python floats."""
    return sum(len(s) for s in streams) * 8 / (H * W)


# a13 + a14  unpack + re-assembly + gather, one image (model.py:269-392)
def unpack_image(t: HuffTable, streams, h: int, w: int, mode: int, codebook):
    offs, caps, total = slot_layout(t, h, w)
    buf = np.zeros(total, np.uint8)
    sizes = np.zeros(5, np.int32)
    for s in range(5):
        sizes[s] = len(streams[s])
        buf[offs[s]: offs[s] + sizes[s]] = np.frombuffer(streams[s], np.uint8)
    cb = _f32(codebook)
    mc = np.zeros((h // 4, w // 4), np.int64)
    mm = np.zeros((h // 2, w // 2), np.int64)
    mf = np.zeros((h, w), np.int64)
    ind = np.zeros((h, w), np.int64)
    quant = np.zeros((4, h, w), np.float32)
    rc = lib().orc_unpack_image(_p(buf, C.c_uint8), _p(offs, C.c_int64), _p(sizes, C.c_int32), h, w, mode, t.K,
                                _p(t.left, C.c_int32), _p(t.right, C.c_int32), t.root, _p(cb, C.c_float),
                                _p(mc, C.c_int64), _p(mm, C.c_int64), _p(mf, C.c_int64), _p(ind, C.c_int64),
                                _p(quant, C.c_float))
    assert rc == 0, rc
    return mc, mm, mf, ind, quant


# ---------------------------------------------------------------------------------------------
# a5  router (RouterTriple.py:15-96)
# ---------------------------------------------------------------------------------------------
def router_mode(coarse_ratio: float, medium_ratio: float) -> int:
    """Mode from which ratios are EXACTLY 0.0 in Python doubles (RouterTriple.py:8-13,19,36,72)."""
    fine = 1 - coarse_ratio - medium_ratio
    zeros = (fine == 0) + (medium_ratio == 0) + (coarse_ratio == 0)
    if zeros == 0:
        return 0
    if zeros == 1:
        return 1 if coarse_ratio == 0 else (2 if medium_ratio == 0 else 3)
    return 4 if coarse_ratio != 0 else (5 if medium_ratio != 0 else 6)


def router_ranks(coarse_ratio: float, medium_ratio: float, n16: int, n8: int, mode: int):
    """k_coarse, k_medium with Python round() on doubles (RouterTriple.py:23,30,42,54,66)."""
    k_c = round(n16 * coarse_ratio)
    if mode == 0:
        k_m = round(4 * n16 * coarse_ratio + n8 * medium_ratio)
    else:
        k_m = round(n8 * medium_ratio)
    return int(k_c), int(k_m)


def router(e16, e8, coarse_ratio: float, medium_ratio: float):
    """-> (mc, mm, mf int32 [B,1,.,.], mode); thresholds across the whole batch like the reference."""
    a16, a8 = _f32(e16), _f32(e8)
    B, h16, w16 = a16.shape
    mode = router_mode(coarse_ratio, medium_ratio)
    k_c, k_m = router_ranks(coarse_ratio, medium_ratio, a16.size, a8.size, mode)
    mc = np.zeros((B, 1, h16, w16), np.int32)
    mm = np.zeros((B, 1, 2 * h16, 2 * w16), np.int32)
    mf = np.zeros((B, 1, 4 * h16, 4 * w16), np.int32)
    rc = lib().orc_router(_p(a16, C.c_float), _p(a8, C.c_float), B, h16, w16, mode, C.c_int64(k_c), C.c_int64(k_m),
                          _p(mc, C.c_int32), _p(mm, C.c_int32), _p(mf, C.c_int32))
    assert rc == 0, rc
    return mc, mm, mf, mode


# a6  mask-mix (vqvae_blocks.py:361-366)
def mask_mix(hc, hm, hf, mc, mm, mf):
    hc, hm, hf = _f32(hc), _f32(hm), _f32(hf)
    B, Cc, h, w = hf.shape
    out = np.empty_like(hf)
    rc = lib().orc_mask_mix(_p(hc, C.c_float), _p(hm, C.c_float), _p(hf, C.c_float), _p(_i32(mc), C.c_int32),
                            _p(_i32(mm), C.c_int32), _p(_i32(mf), C.c_int32), B, Cc, h, w, _p(out, C.c_float))
    assert rc == 0, rc
    return out


# a4  Entropy (model.py:440-483)
def linspace_bins() -> np.ndarray:
    """torch.linspace(-1, 1, 32) in fp32 (model.py:480): start + i*step below the midpoint,
    end - (steps-1-i)*step from it on, step = (end-start)/(steps-1) in fp32."""
    step = np.float32(2.0) / np.float32(31.0)
    i = np.arange(32)
    lo = (np.float32(-1.0) + step * i.astype(np.float32)).astype(np.float32)
    hi = (np.float32(1.0) - step * (31 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < 16, lo, hi).astype(np.float32)


def entropy(x, psize: int, bins=None):
    a = _f32(x)
    B, c, H, W = a.shape
    assert c == 3
    bins = linspace_bins() if bins is None else _f32(bins)
    out = np.zeros((B, H // psize, W // psize), np.float32)
    rc = lib().orc_entropy(_p(a, C.c_float), B, H, W, psize, _p(bins, C.c_float), _p(out, C.c_float))
    assert rc == 0, rc
    return out


# f4  SpatialNorm (decoder.py:34-53), numpy restatement; statistics in float64
def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    """torch's nearest rule (F.interpolate(mode="nearest"), decoder.py:49): min(floor(dst * fp32(in / out)), in - 1)."""
    scale = np.float32(in_size) / np.float32(out_size)
    return np.minimum(np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64), in_size - 1)


def spatial_norm(f, zq, gn_weight, gn_bias, wy, by, wb, bb, groups: int, eps: float):
    """decoder.py:47-53 without the optional 3x3 conv: GroupNorm(f) * conv_y(nearest(zq)) + conv_b(nearest(zq)).
    f [B,C,H,W], zq [B,Cz,hz,wz]; wy / wb [C,Cz]; vectors [C] or None."""
    f = np.asarray(f, np.float64)
    zq = np.asarray(zq, np.float64)
    B, Cc, H, W = f.shape
    zu = zq[:, :, nearest_index(H, zq.shape[2])][:, :, :, nearest_index(W, zq.shape[3])]
    g = f.reshape(B, groups, -1)
    mean = g.mean(axis=2, keepdims=True)
    var = g.var(axis=2, keepdims=True)                                   # biased (decoder.py:52 -> nn.GroupNorm)
    n = ((g - mean) / np.sqrt(var + eps)).reshape(B, Cc, H, W)
    one = np.ones(Cc)
    zero = np.zeros(Cc)
    n = n * (one if gn_weight is None else np.asarray(gn_weight, np.float64)).reshape(1, Cc, 1, 1) \
        + (zero if gn_bias is None else np.asarray(gn_bias, np.float64)).reshape(1, Cc, 1, 1)
    cy = np.einsum("ck,bkhw->bchw", np.asarray(wy, np.float64).reshape(Cc, -1), zu) + (zero if by is None else np.asarray(by, np.float64)).reshape(1, Cc, 1, 1)
    cb = np.einsum("ck,bkhw->bchw", np.asarray(wb, np.float64).reshape(Cc, -1), zu) + (zero if bb is None else np.asarray(bb, np.float64)).reshape(1, Cc, 1, 1)
    return (n * cy + cb).astype(np.float32)
