"""Tensor-level wrappers over the C-ABI (include/cgic_b200.h).

PyTorch is only plumbing here: it owns device memory and the stream; every op below enqueues
hand-written sm_100a kernels from libcgic_b200.so on torch's current stream and returns torch
tensors.  Inputs must be CUDA tensors -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib

STREAM_NAMES = ("indices_coarse", "indices_medium", "indices_fine", "mask_coarse", "mask_medium")
_STREAM_BITS = (0x1F, 0x16, 0x0D, 0x0B, 0x01, 0x02, 0x04)


def stream_present(mode: int, s: int) -> bool:
    """Which of the five files exist per compression mode (CGIC/models/model.py:225-260)."""
    return bool((_STREAM_BITS[mode] >> s) & 1)


def _cuda(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def tune(key: str, value: int) -> None:
    """Dispatch knobs (cgic_tune): "fused_decode_ctas" in (-1, 0, 1, 2, 4), "fused_encode" in (0, 1), "pack_image" in (-1, 0, 1).  Which kernels serve a
    call changes, the results do not."""
    check(lib().cgic_tune(key.encode(), int(value)), "cgic_tune")


def _on_tensor_device(fn):
    """Runs `fn` with the device of its first CUDA tensor argument current, so that the launch, torch's current stream
    and the per-device set-up inside the library (shared-memory opt-ins, uploaded tables) all refer to the device the
    data lives on -- not to whatever device happens to be current in the caller."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            dev = a.device if isinstance(a, torch.Tensor) else getattr(a, "device", None)
            if isinstance(dev, torch.device) and dev.type == "cuda":
                if dev.index is not None and dev.index != torch.cuda.current_device():
                    with torch.cuda.device(dev):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------------
# a8  Huffman table
# --------------------------------------------------------------------------------------------
class HuffTable:
    """Owns a `cgic_table*`.  freq[s] = int count of symbol s, order = push order (None: 0..K-1)."""

    def __init__(self, freq: Sequence[int], order: Optional[Sequence[int]] = None):
        f = np.ascontiguousarray(np.asarray(freq, dtype=np.int64))
        o = None if order is None else np.ascontiguousarray(np.asarray(order, dtype=np.int32))
        if o is not None and o.shape != f.shape:
            raise ValueError("order must list every symbol once")
        handle = C.c_void_p()
        check(lib().cgic_huff_build(f.ctypes.data, None if o is None else o.ctypes.data, int(f.shape[0]), C.byref(handle)),
              "cgic_huff_build")
        self._h = handle
        self._free = lib().cgic_huff_free          # bound now: module globals may be gone at interpreter exit
        self.K = int(f.shape[0])
        self.max_len = lib().cgic_huff_max_len(self._h)
        self._uploaded_on = set()
        self._layouts = {}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._free(h)

    @property
    def handle(self):
        return self._h

    def code(self, sym: int) -> str:
        buf = C.create_string_buffer(self.max_len + 1)
        n = lib().cgic_huff_code(self._h, int(sym), buf, self.max_len + 1)
        if n < 0:
            check(n, "cgic_huff_code")
        return buf.value.decode()

    def codes(self) -> dict:
        return {s: self.code(s) for s in range(self.K)}

    def lengths(self) -> np.ndarray:
        return np.asarray([lib().cgic_huff_code_len(self._h, s) for s in range(self.K)], np.int32)

    def upload(self) -> "HuffTable":
        dev = torch.cuda.current_device()
        if dev not in self._uploaded_on:          # one immutable copy per device
            check(lib().cgic_huff_upload(self._h), "cgic_huff_upload")
            self._uploaded_on.add(dev)
        return self

    def stream_capacity(self, n_symbols: int) -> int:
        return int(lib().cgic_huff_stream_capacity(self._h, int(n_symbols)))

    def layout(self, h: int, w: int) -> Tuple[np.ndarray, np.ndarray, int]:
        """(slot_off[5], slot_cap[5], image_stride) of the packed-image layout for an h x w token grid."""
        key = (h, w)
        if key not in self._layouts:
            off = np.zeros(5, np.int64)
            cap = np.zeros(5, np.int64)
            stride = C.c_int64(0)
            check(lib().cgic_pack_layout(self._h, h, w, off.ctypes.data, cap.ctypes.data, C.addressof(stride)), "cgic_pack_layout")
            self._layouts[key] = (off, cap, int(stride.value))
        return self._layouts[key]


# --------------------------------------------------------------------------------------------
# a1  VQ
# --------------------------------------------------------------------------------------------
_ws_cache = {}
_WS_PER_OP = 8          # scratch buffers kept per (op, device): one per recently used stream
_warned_alias = False


def _workspace(tag: str, nbytes: int, device: torch.device, geom=()) -> torch.Tensor:
    """Scratch buffer per (tag, geometry, device, stream), zero-filled when it is created (the kernels leave it clean);
    allocation happens off the hot path after warm-up.  The geometry is part of the key because the carve-up of a
    workspace depends on it: the regions one geometry needs zeroed would otherwise lie over another one's scratch data.

    Under CUDA-graph capture a miss would put the zero-fill INTO the graph: a ~2 us fill kernel in front of the kernel on
    every replay, which also severs the programmatic (PDL) edge to its predecessor.  So while capturing, a buffer of the
    same tag that a warm-up run created on another stream of this device is reused instead (warm up before capturing, as
    torch asks anyway; do not replay the graph concurrently with eager calls of the same op on that other stream)."""
    tag = (tag,) + tuple(geom)
    key = (tag, device.index, _stream())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            for (t, d, st), b in reversed(list(_ws_cache.items())):
                if t == tag and d == device.index and b.numel() >= nbytes:
                    global _warned_alias
                    if not _warned_alias:
                        _warned_alias = True
                        warnings.warn(f"cgic_b200.ops: CUDA-graph capture of '{tag[0]}' reuses the scratch buffer of stream {st:#x}; "
                                      "do not run that stream's eager calls concurrently with replays of this graph "
                                      "(warm up on the capture stream to give the graph a buffer of its own)")
                    _ws_cache[key] = b
                    return b
        buf = torch.zeros(max(nbytes, 256), dtype=torch.uint8, device=device)   # cgic_vq_assign wants its ticket zeroed once
        _ws_cache[key] = buf
        # bounded: streams come and go (torch.cuda.Stream() per request, graph captures); keep the newest few per op and device
        mine = [k for k in _ws_cache if k[0][0] == tag[0] and k[1] == device.index]
        for k in mine[:-_WS_PER_OP]:
            del _ws_cache[k]
    return buf


class Codebook:
    """Owns a `cgic_codebook*`: the device copy of `embedding.weight` (quantize.py:25-26) plus the cell
    index cgic_vq_assign_indexed searches.  update() rebuilds it on the current stream (no sync)."""

    def __init__(self, weight: torch.Tensor):
        w = _cuda(weight.detach(), torch.float32, "codebook")
        if w.dim() != 2 or w.shape[1] != 4:
            raise ValueError(f"Codebook supports e_dim == 4 (got {tuple(w.shape)})")
        handle = C.c_void_p()
        with torch.cuda.device(w.device):
            check(lib().cgic_codebook_create(int(w.shape[0]), C.byref(handle)), "cgic_codebook_create")
        self._h = handle
        self._free = lib().cgic_codebook_free
        self.K = int(w.shape[0])
        self.device = w.device
        self.update(w)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._free(h)

    @property
    def handle(self):
        return self._h

    def update(self, weight: torch.Tensor) -> "Codebook":
        w = _cuda(weight.detach(), torch.float32, "codebook")
        if tuple(w.shape) != (self.K, 4) or w.device != self.device:
            raise ValueError(f"codebook {tuple(w.shape)} on {w.device} does not match the prepared [{self.K}, 4] on {self.device}")
        with torch.cuda.device(self.device):
            check(lib().cgic_codebook_update(self._h, w.data_ptr(), _stream()), "cgic_codebook_update")
        return self

    def check(self, weight: torch.Tensor) -> None:
        """Enqueue the staleness guard against the live weights (no sync; see cgic_codebook_check)."""
        w = _cuda(weight.detach(), torch.float32, "codebook")
        with torch.cuda.device(self.device):
            check(lib().cgic_codebook_check(self._h, w.data_ptr(), _stream()), "cgic_codebook_check")

    def is_stale(self) -> bool:
        """True once a check() since the last update() found the weights changed (host flag, no sync)."""
        return lib().cgic_codebook_is_stale(self._h) == 1

    def stats(self) -> dict:
        """{valid, cells, max_list, overflow_cells} (synchronises)."""
        out = np.zeros(4, np.int32)
        check(lib().cgic_codebook_stats_host(self._h, out.ctypes.data), "cgic_codebook_stats_host")
        return dict(valid=int(out[0]), cells=int(out[1]), max_list=int(out[2]), overflow_cells=int(out[3]))


@_on_tensor_device
def vq_assign(z: torch.Tensor, codebook, want_zq: bool = True, want_sqerr: bool = True):
    """quantize.py:69-98 -> (idx int64 [B*h*w], z_q fp32 NCHW or None, sqerr float64[1] or None).
    `codebook`: a [K,4] fp32 CUDA tensor (exhaustive search) or a prepared `Codebook` (indexed search,
    identical results)."""
    z = _cuda(z, torch.float32, "z")
    indexed = isinstance(codebook, Codebook)
    cb = None if indexed else _cuda(codebook, torch.float32, "codebook")
    if z.dim() != 4 or z.shape[1] != 4 or (cb is not None and (cb.dim() != 2 or cb.shape[1] != 4)):
        raise ValueError(f"vq_assign supports e_dim == 4 (z {tuple(z.shape)})")
    B, _, h, w = z.shape
    n = B * h * w
    idx = torch.empty(n, dtype=torch.int64, device=z.device)
    zq = torch.empty_like(z) if want_zq else None
    sq = torch.empty(1, dtype=torch.float64, device=z.device) if want_sqerr else None
    nbytes = lib().cgic_vq_workspace_bytes(n)
    ws = _workspace("vq", nbytes, z.device, (n,))
    if indexed:
        check(lib().cgic_vq_assign_indexed(z.data_ptr(), B, h, w, codebook.handle, idx.data_ptr(), _p(zq), _p(sq),
                                           ws.data_ptr(), ws.numel(), _stream()), "cgic_vq_assign_indexed")
    else:
        check(lib().cgic_vq_assign(z.data_ptr(), B, h, w, cb.data_ptr(), cb.shape[0], idx.data_ptr(), _p(zq), _p(sq),
                                   ws.data_ptr(), ws.numel(), _stream()), "cgic_vq_assign")
    return idx, zq, sq


@_on_tensor_device
def vq_count(idx: torch.Tensor, counters: torch.Tensor) -> None:
    """quantize.py:79-81: counters[idx[i]] += 1, counters fp32 [K] (in place)."""
    idx = _cuda(idx, torch.int64, "idx")
    if not (counters.is_cuda and counters.dtype == torch.float32 and counters.is_contiguous()):
        raise TypeError("counters must be a contiguous fp32 CUDA tensor")
    check(lib().cgic_vq_count(idx.data_ptr(), idx.numel(), counters.data_ptr(), counters.numel(), _stream()), "cgic_vq_count")


# --------------------------------------------------------------------------------------------
# a4 entropy, a5 router, a6 mask-mix
# --------------------------------------------------------------------------------------------
_BINS = None


def linspace_bins() -> np.ndarray:
    """The 32 fp32 values of torch.linspace(-1, 1, 32) (model.py:480), computed by torch on the host."""
    global _BINS
    if _BINS is None:
        _BINS = np.ascontiguousarray(torch.linspace(-1, 1, 32).numpy().astype(np.float32))
    return _BINS


@_on_tensor_device
def entropy_maps(x: torch.Tensor, want8: bool = True, want16: bool = True):
    """model.py:440-483 for patch sizes 8 and 16 in one pass -> (e8 [B,H/8,W/8], e16 [B,H/16,W/16])."""
    x = _cuda(x, torch.float32, "x")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("entropy_maps expects [B,3,H,W]")
    B, _, H, W = x.shape
    e8 = torch.empty(B, H // 8, W // 8, dtype=torch.float32, device=x.device) if want8 else None
    e16 = torch.empty(B, H // 16, W // 16, dtype=torch.float32, device=x.device) if want16 else None
    check(lib().cgic_entropy_maps(x.data_ptr(), B, H, W, linspace_bins().ctypes.data, _p(e8), _p(e16), _stream()),
          "cgic_entropy_maps")
    return e8, e16


def router_mode(coarse_ratio: float, medium_ratio: float) -> int:
    """Mode from which ratios are exactly 0.0 in Python doubles (RouterTriple.py:8-13,19,36,72)."""
    fine_ratio = 1 - coarse_ratio - medium_ratio
    zeros = (fine_ratio == 0) + (medium_ratio == 0) + (coarse_ratio == 0)
    if zeros == 0:
        return 0
    if zeros == 1:
        return 1 if coarse_ratio == 0 else (2 if medium_ratio == 0 else 3)
    return 4 if coarse_ratio != 0 else (5 if medium_ratio != 0 else 6)


def router_ranks(coarse_ratio: float, medium_ratio: float, n16: int, n8: int, mode: int) -> Tuple[int, int]:
    """k_coarse, k_medium: Python round() (banker's) on double products (RouterTriple.py:23,30,42,54,66)."""
    k_c = round(n16 * coarse_ratio)
    k_m = round(4 * n16 * coarse_ratio + n8 * medium_ratio) if mode == 0 else round(n8 * medium_ratio)
    return int(k_c), int(k_m)


@_on_tensor_device
def router(e16: torch.Tensor, e8: torch.Tensor, coarse_ratio: float, medium_ratio: float, per_image: bool = False,
           want_gate: bool = False):
    """RouterTriple.py:15-96 -> (m_c, m_m, m_f int32 [B,1,.,.], gate fp32 [B,1,h,3w] or None, mode)."""
    e16 = _cuda(e16, torch.float32, "e16")
    e8 = _cuda(e8, torch.float32, "e8")
    B, h16, w16 = e16.shape
    if tuple(e8.shape) != (B, 2 * h16, 2 * w16):
        raise ValueError(f"e8 {tuple(e8.shape)} does not match e16 {tuple(e16.shape)}")
    mode = router_mode(coarse_ratio, medium_ratio)
    nimg = 1 if per_image else B
    k_c, k_m = router_ranks(coarse_ratio, medium_ratio, nimg * h16 * w16, nimg * 4 * h16 * w16, mode)
    dev = e16.device
    m_c = torch.empty(B, 1, h16, w16, dtype=torch.int32, device=dev)
    m_m = torch.empty(B, 1, 2 * h16, 2 * w16, dtype=torch.int32, device=dev)
    m_f = torch.empty(B, 1, 4 * h16, 4 * w16, dtype=torch.int32, device=dev)
    gate = torch.empty(B, 1, 4 * h16, 12 * w16, dtype=torch.float32, device=dev) if want_gate else None
    check(lib().cgic_router(e16.data_ptr(), e8.data_ptr(), B, h16, w16, mode, k_c, k_m, int(per_image), m_c.data_ptr(),
                            m_m.data_ptr(), m_f.data_ptr(), _p(gate), None, 0, _stream()), "cgic_router")
    return m_c, m_m, m_f, gate, mode


ENTROPY_RTOL, ENTROPY_ATOL = 2e-5, 1e-6   # float tolerance of Entropy against the reference (tests)


@_on_tensor_device
def entropy_route(x: torch.Tensor, coarse_ratio: float, medium_ratio: float, rtol: float = ENTROPY_RTOL, atol: float = ENTROPY_ATOL):
    """f1, first launch: model.py:440-483 for both patch sizes AND RouterTriple.py:15-96 with per-image thresholds, the image
    read once -> (e8 [B,H/8,W/8], e16 [B,H/16,W/16], m_c, m_m int32 [B,1,.,.], near int32 [B,2], mode).  near[b] = how many
    coarse / medium entropies of image b lie within rtol*|thr| + atol of their threshold."""
    x = _cuda(x, torch.float32, "x")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError("entropy_route expects [B,3,H,W]")
    B, _, H, W = x.shape
    dev = x.device
    h16, w16 = H // 16, W // 16
    mode = router_mode(coarse_ratio, medium_ratio)
    k_c, k_m = router_ranks(coarse_ratio, medium_ratio, h16 * w16, 4 * h16 * w16, mode)
    e8 = torch.empty(B, H // 8, W // 8, dtype=torch.float32, device=dev)
    e16 = torch.empty(B, h16, w16, dtype=torch.float32, device=dev)
    m_c = torch.empty(B, 1, h16, w16, dtype=torch.int32, device=dev)
    m_m = torch.empty(B, 1, 2 * h16, 2 * w16, dtype=torch.int32, device=dev)
    near = torch.empty(B, 2, dtype=torch.int32, device=dev)
    ws = _workspace("entropy_route", lib().cgic_entropy_route_workspace_bytes(B), dev, (B,))
    check(lib().cgic_entropy_route(x.data_ptr(), B, H, W, linspace_bins().ctypes.data, e8.data_ptr(), e16.data_ptr(), mode, k_c, k_m,
                                   float(rtol), float(atol), m_c.data_ptr(), m_m.data_ptr(), near.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
          "cgic_entropy_route")
    return e8, e16, m_c, m_m, near, mode


@_on_tensor_device
def route_mix(h_c, h_m, h_f, m_c, m_m, mode: int, want_gate: bool = False):
    """f1, second launch: fine mask (RouterTriple.py:34) + gate + mask-mix (vqvae_blocks.py:361-366) -> (m_f int32 [B,1,h,w],
    gate fp32 [B,1,h,3w] or None, h fp32 [B,C,h,w])."""
    h_c, h_m, h_f = (_cuda(t, torch.float32, n) for t, n in ((h_c, "h_c"), (h_m, "h_m"), (h_f, "h_f")))
    m_c, m_m = (_cuda(t, torch.int32, n) for t, n in ((m_c, "m_c"), (m_m, "m_m")))
    B, Cc, h, w = h_f.shape
    dev = h_f.device
    m_f = torch.empty(B, 1, h, w, dtype=torch.int32, device=dev)
    gate = torch.empty(B, 1, h, 3 * w, dtype=torch.float32, device=dev) if want_gate else None
    out = torch.empty_like(h_f)
    check(lib().cgic_route_mix(h_c.data_ptr(), h_m.data_ptr(), h_f.data_ptr(), m_c.data_ptr(), m_m.data_ptr(), mode, B, Cc, h, w,
                               m_f.data_ptr(), _p(gate), out.data_ptr(), _stream()), "cgic_route_mix")
    return m_f, gate, out


@_on_tensor_device
def mask_mix(h_c, h_m, h_f, m_c, m_m, m_f) -> torch.Tensor:
    """vqvae_blocks.py:361-366: up4(h_c)*up4(m_c) + up2(h_m)*up2(m_m) + h_f*m_f."""
    h_c, h_m, h_f = (_cuda(t, torch.float32, n) for t, n in ((h_c, "h_c"), (h_m, "h_m"), (h_f, "h_f")))
    m_c, m_m, m_f = (_cuda(t, torch.int32, n) for t, n in ((m_c, "m_c"), (m_m, "m_m"), (m_f, "m_f")))
    B, Cc, h, w = h_f.shape
    out = torch.empty_like(h_f)
    check(lib().cgic_mask_mix(h_c.data_ptr(), h_m.data_ptr(), h_f.data_ptr(), m_c.data_ptr(), m_m.data_ptr(), m_f.data_ptr(),
                              B, Cc, h, w, out.data_ptr(), _stream()), "cgic_mask_mix")
    return out


@_on_tensor_device
def decoder_merge(h: torch.Tensor, other: torch.Tensor, masks, level: int) -> torch.Tensor:
    """decoder.py:373-382, the mask-gated merge at the decoder's entry: level 2: h*up2(mask[0]) + other*mask[1];
    level 3: h*up4(mask[0]) + h*up2(mask[1]) + other*mask[2].  masks: [B,1,.,.] int32 / int64 / float32, one dtype."""
    h = _cuda(h, torch.float32, "h")
    other = _cuda(other, torch.float32, "other")
    if h.shape != other.shape or h.dim() != 4 or level not in (2, 3):
        raise ValueError("decoder_merge: h and other must be [B,C,hh,ww] tensors of one shape, level 2 or 3")
    dt = masks[0].dtype
    elem = {torch.int32: 4, torch.int64: 8, torch.float32: -4}.get(dt)
    if elem is None or any(m.dtype != dt for m in masks[:level]):
        raise TypeError("decoder_merge: masks must all be int32, int64 or float32")
    ms = [_cuda(m, dt, "mask") for m in masks[:level]]
    B, Cc, hh, ww = h.shape
    div = 2 if level == 2 else 4
    want = [(B, 1, hh // div, ww // div), (B, 1, hh // (div // 2), ww // (div // 2))] + ([(B, 1, hh, ww)] if level == 3 else [])
    if [tuple(m.shape) for m in ms] != want:
        raise ValueError(f"decoder_merge: mask shapes {[tuple(m.shape) for m in ms]} do not match {want}")
    out = torch.empty_like(h)
    check(lib().cgic_decoder_merge(h.data_ptr(), other.data_ptr(), ms[0].data_ptr(), ms[1].data_ptr(), ms[2].data_ptr() if level == 3 else None,
                                   elem, level, B, Cc, hh, ww, out.data_ptr(), _stream()), "cgic_decoder_merge")
    return out


@_on_tensor_device
def spatial_norm(f: torch.Tensor, zq: torch.Tensor, gn_weight: Optional[torch.Tensor], gn_bias: Optional[torch.Tensor],
                 wy: torch.Tensor, by: Optional[torch.Tensor], wb: torch.Tensor, bb: Optional[torch.Tensor],
                 groups: int, eps: float) -> torch.Tensor:
    """decoder.py:47-53 (SpatialNorm.forward without the optional 3x3 conv): GroupNorm(f) * conv_y(nearest(zq)) + conv_b(nearest(zq)).
    f [B,C,H,W], zq [B,Cz,hz,wz] fp32; wy / wb [C,Cz] or [C,Cz,1,1]; by / bb / gn_weight / gn_bias [C] or None."""
    f = _cuda(f, torch.float32, "f")
    zq = _cuda(zq, torch.float32, "zq")
    if f.dim() != 4 or zq.dim() != 4 or zq.shape[0] != f.shape[0]:
        raise ValueError(f"spatial_norm: f {tuple(f.shape)} / zq {tuple(zq.shape)} must be [B,C,H,W] / [B,Cz,hz,wz]")
    B, Cc, H, W = f.shape
    _, Cz, hz, wz = zq.shape
    wy = _cuda(wy.reshape(wy.shape[0], -1), torch.float32, "wy")
    wb = _cuda(wb.reshape(wb.shape[0], -1), torch.float32, "wb")
    if tuple(wy.shape) != (Cc, Cz) or tuple(wb.shape) != (Cc, Cz):
        raise ValueError(f"spatial_norm: 1x1 conv weights must be [{Cc},{Cz}], got {tuple(wy.shape)} / {tuple(wb.shape)}")
    vecs = [None if t is None else _cuda(t.reshape(-1), torch.float32, n) for t, n in ((gn_weight, "gn_weight"), (gn_bias, "gn_bias"), (by, "by"), (bb, "bb"))]
    if any(t is not None and t.numel() != Cc for t in vecs):
        raise ValueError(f"spatial_norm: per-channel vectors must have {Cc} elements")
    out = torch.empty_like(f)
    ws = _workspace("spatial_norm", lib().cgic_spatial_norm_workspace_bytes(B, groups), f.device, (B, groups))
    check(lib().cgic_spatial_norm(f.data_ptr(), zq.data_ptr(), _p(vecs[0]), _p(vecs[1]), wy.data_ptr(), _p(vecs[2]), wb.data_ptr(), _p(vecs[3]),
                                  B, Cc, H, W, Cz, hz, wz, groups, float(eps), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
          "cgic_spatial_norm")
    return out


# --------------------------------------------------------------------------------------------
# a7/a9/a11/a12 pack and a10/a13/a14 unpack (batched, B independent images)
# --------------------------------------------------------------------------------------------
@_on_tensor_device
def pack(idx: torch.Tensor, m_c, m_m, m_f, mode: int, table: HuffTable, h: int, w: int):
    """model.py:217-260 for B images -> (bytes uint8 [B, image_stride], sizes int32 [B,5])."""
    idx = _cuda(idx, torch.int64, "idx")
    m_c, m_m, m_f = (_cuda(t, torch.int32, n) for t, n in ((m_c, "m_c"), (m_m, "m_m"), (m_f, "m_f")))
    B = idx.numel() // (h * w)
    table.upload()
    _, _, stride = table.layout(h, w)
    out = torch.empty(B, stride, dtype=torch.uint8, device=idx.device)
    sizes = torch.empty(B, 5, dtype=torch.int32, device=idx.device)
    ws = _workspace("pack", lib().cgic_pack_workspace_bytes(B, h, w), idx.device, (B, h, w))
    check(lib().cgic_pack_ws(idx.data_ptr(), m_c.data_ptr(), m_m.data_ptr(), m_f.data_ptr(), B, h, w, mode, table.handle,
                             out.data_ptr(), sizes.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "cgic_pack_ws")
    return out, sizes


@_on_tensor_device
def encode(z: torch.Tensor, codebook: "Codebook", m_c, m_m, m_f, mode: int, table: HuffTable, want_zq: bool = True, want_sqerr: bool = True):
    """The encoder half in one call (cgic_encode): quantize.py:69-98 + model.py:217-260 for B images ->
    (idx int64 [B*h*w], z_q fp32 NCHW or None, sqerr float64[1] or None, bytes uint8 [B, image_stride], sizes int32 [B,5]);
    identical to vq_assign(z, codebook) followed by pack(idx, ...).  One launch on token grids of at most 4096 cells."""
    if not isinstance(codebook, Codebook):
        raise TypeError("encode needs a prepared Codebook (ops.Codebook(weight))")
    z = _cuda(z, torch.float32, "z")
    m_c, m_m, m_f = (_cuda(t, torch.int32, n) for t, n in ((m_c, "m_c"), (m_m, "m_m"), (m_f, "m_f")))
    if z.dim() != 4 or z.shape[1] != 4:
        raise ValueError(f"encode supports e_dim == 4 (z {tuple(z.shape)})")
    B, _, h, w = z.shape
    table.upload()
    _, _, stride = table.layout(h, w)
    dev = z.device
    idx = torch.empty(B * h * w, dtype=torch.int64, device=dev)
    zq = torch.empty_like(z) if want_zq else None
    sq = torch.empty(1, dtype=torch.float64, device=dev) if want_sqerr else None
    out = torch.empty(B, stride, dtype=torch.uint8, device=dev)
    sizes = torch.empty(B, 5, dtype=torch.int32, device=dev)
    ws = _workspace("encode", lib().cgic_encode_workspace_bytes(B, h, w), dev, (B, h, w))
    check(lib().cgic_encode(z.data_ptr(), m_c.data_ptr(), m_m.data_ptr(), m_f.data_ptr(), B, h, w, mode, codebook.handle, table.handle,
                            idx.data_ptr(), _p(zq), _p(sq), out.data_ptr(), sizes.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
          "cgic_encode")
    return idx, zq, sq, out, sizes


def exhaustive_count(tag: str = "encode") -> int:
    """Running count (since the workspace was created) of latents the indexed search handed to its exhaustive path, summed
    over this op's workspaces (synchronises).  tag: "encode" or "vq"."""
    return int(sum(int(b[8:12].view(torch.int32).item()) for k, b in _ws_cache.items() if k[0][0] == tag))


@_on_tensor_device
def unpack(bytes_: torch.Tensor, sizes: torch.Tensor, mode: int, table: HuffTable, codebook: torch.Tensor, h: int, w: int):
    """model.py:269-392 for B images -> (mc, mm, mf int64, ind int64 [B,h,w], quant fp32 [B,4,h,w], status int32 [B])."""
    bytes_ = _cuda(bytes_, torch.uint8, "bytes")
    sizes = _cuda(sizes, torch.int32, "sizes")
    cb = _cuda(codebook, torch.float32, "codebook")
    B = sizes.shape[0]
    table.upload()
    _, _, stride = table.layout(h, w)
    if bytes_.numel() != B * stride:
        raise ValueError(f"bytes has {bytes_.numel()} elements, expected {B}*{stride}")
    dev = bytes_.device
    mc = torch.empty(B, h // 4, w // 4, dtype=torch.int64, device=dev)
    mm = torch.empty(B, h // 2, w // 2, dtype=torch.int64, device=dev)
    mf = torch.empty(B, h, w, dtype=torch.int64, device=dev)
    ind = torch.empty(B, h, w, dtype=torch.int64, device=dev)
    quant = torch.empty(B, 4, h, w, dtype=torch.float32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    nbytes = lib().cgic_unpack_workspace_bytes(B, h, w)
    ws = _workspace("unpack", nbytes, dev, (B, h, w))
    check(lib().cgic_unpack(bytes_.data_ptr(), sizes.data_ptr(), B, h, w, mode, table.handle, cb.data_ptr(), mc.data_ptr(),
                            mm.data_ptr(), mf.data_ptr(), ind.data_ptr(), quant.data_ptr(), status.data_ptr(), ws.data_ptr(),
                            ws.numel(), _stream()), "cgic_unpack")
    return mc, mm, mf, ind, quant, status


# --------------------------------------------------------------------------------------------
# single-stream codec ops (a9, a10, a11)
# --------------------------------------------------------------------------------------------
@_on_tensor_device
def huff_encode(symbols: torch.Tensor, table: HuffTable) -> bytes:
    """indices_coding.py:113-126 payload: 1-D integer CUDA tensor -> bytes (b'' if empty)."""
    s = _cuda(symbols.reshape(-1), symbols.dtype, "symbols")
    if s.dtype != torch.int64:
        s = s.to(torch.int64)
    n = s.numel()
    if n == 0:
        return b""
    table.upload()
    cap = (table.stream_capacity(n) + 15) // 16 * 16
    out = torch.empty(cap, dtype=torch.uint8, device=s.device)
    size = torch.empty(1, dtype=torch.int32, device=s.device)
    check(lib().cgic_huff_encode(s.data_ptr(), n, table.handle, out.data_ptr(), cap, size.data_ptr(), _stream()), "cgic_huff_encode")
    nb = int(size.item())
    if nb < 0:
        raise KeyError("symbol outside the code table")  # the reference raises KeyError from self.codes[character]
    return bytes(out[:nb].cpu().numpy())


def huff_decode(data: bytes, table: HuffTable, device) -> Optional[list]:
    """indices_coding.py:153-168: bytes -> list of symbols, None for the empty file."""
    if len(data) == 0:
        return None
    table.upload()
    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
    cap = len(data) * 8
    out = torch.empty(cap, dtype=torch.int32, device=device)
    cnt = torch.empty(1, dtype=torch.int32, device=device)
    check(lib().cgic_huff_decode(buf.data_ptr(), len(data), table.handle, out.data_ptr(), cap, cnt.data_ptr(), _stream()),
          "cgic_huff_decode")
    n = int(cnt.item())
    if n < 0:
        raise RuntimeError(f"cgic_huff_decode: status {n}")
    return out[:n].cpu().tolist()


@_on_tensor_device
def bits_encode(values: torch.Tensor) -> bytes:
    """mask_coding.py:40-55 payload."""
    v = _cuda(values.reshape(-1), values.dtype, "values")
    if v.dtype != torch.int32:
        v = v.to(torch.int32)
    n = v.numel()
    if n == 0:
        return b""
    if v.data_ptr() % 16:
        v = v.clone()
    cap = n // 8 + 2
    out = torch.empty(cap, dtype=torch.uint8, device=v.device)
    size = torch.empty(1, dtype=torch.int32, device=v.device)
    check(lib().cgic_bits_encode(v.data_ptr(), n, out.data_ptr(), cap, size.data_ptr(), _stream()), "cgic_bits_encode")
    return bytes(out[: int(size.item())].cpu().numpy())


def bits_decode(data: bytes, device) -> Optional[list]:
    """mask_coding.py:81-96."""
    if len(data) == 0:
        return None
    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
    cap = len(data) * 8
    out = torch.empty(cap, dtype=torch.int32, device=device)
    cnt = torch.empty(1, dtype=torch.int32, device=device)
    check(lib().cgic_bits_decode(buf.data_ptr(), len(data), out.data_ptr(), cap, cnt.data_ptr(), _stream()), "cgic_bits_decode")
    return out[: int(cnt.item())].cpu().tolist()


# --------------------------------------------------------------------------------------------
# host-buffer session (the e2e call)
# --------------------------------------------------------------------------------------------
class Session:
    """cgic_session: device arena + stream created once; compress/decompress take HOST tensors
    (pin them for full PCIe speed), copy in, run the kernels, copy out, synchronise."""

    def __init__(self, B: int, h: int, w: int, mode: int, table: HuffTable, codebook: torch.Tensor):
        cb = np.ascontiguousarray(codebook.detach().cpu().numpy().astype(np.float32))
        handle = C.c_void_p()
        table.upload()
        check(lib().cgic_session_create(B, h, w, mode, table.handle, cb.ctypes.data, cb.shape[0], C.byref(handle)),
              "cgic_session_create")
        self._s = handle
        self._destroy = lib().cgic_session_destroy
        self._table = table
        self.B, self.h, self.w, self.mode = B, h, w, mode
        self.image_stride = int(lib().cgic_session_image_stride(self._s))
        pin = dict(pin_memory=True)
        n4 = B * h * w
        self.bytes = torch.empty(B, self.image_stride, dtype=torch.uint8, **pin)
        self.sizes = torch.empty(B, 5, dtype=torch.int32, **pin)
        self.idx = torch.empty(n4, dtype=torch.int64, **pin)
        self.mc = torch.empty(B, h // 4, w // 4, dtype=torch.int64, **pin)
        self.mm = torch.empty(B, h // 2, w // 2, dtype=torch.int64, **pin)
        self.mf = torch.empty(B, h, w, dtype=torch.int64, **pin)
        self.ind = torch.empty(B, h, w, dtype=torch.int64, **pin)
        self.quant = torch.empty(B, 4, h, w, dtype=torch.float32, **pin)
        self.status = torch.empty(B, dtype=torch.int32, **pin)

    def set_pipeline(self, parts: int) -> "Session":
        """Number of image ranges (streams) the batch is pipelined in; default min(B, 4)."""
        check(lib().cgic_session_set_pipeline(self._s, int(parts)), "cgic_session_set_pipeline")
        return self

    def close(self):
        s, self._s = getattr(self, "_s", None), None
        if s:
            self._destroy(s)

    __del__ = close

    @staticmethod
    def _host(t: torch.Tensor, dtype, name):
        if t.is_cuda or t.dtype != dtype or not t.is_contiguous():
            raise TypeError(f"{name}: expected a contiguous host tensor of {dtype}")
        return t.data_ptr()

    def compress(self, z, m_c, m_m, m_f, want_idx: bool = False):
        """host z fp32 [B,4,h,w] + int32 masks -> (bytes [B,stride] uint8, sizes [B,5] int32[, idx]) (pinned host)."""
        check(lib().cgic_session_compress_host(self._s, self._host(z, torch.float32, "z"), self._host(m_c, torch.int32, "m_c"),
                                               self._host(m_m, torch.int32, "m_m"), self._host(m_f, torch.int32, "m_f"),
                                               self.bytes.data_ptr(), self.sizes.data_ptr(),
                                               self.idx.data_ptr() if want_idx else None, None, None),
              "cgic_session_compress_host")
        return (self.bytes, self.sizes, self.idx) if want_idx else (self.bytes, self.sizes)

    def decompress(self, bytes_, sizes):
        """host bytes/sizes -> (mc, mm, mf, ind int64, quant fp32, status) (pinned host)."""
        check(lib().cgic_session_decompress_host(self._s, self._host(bytes_, torch.uint8, "bytes"),
                                                 self._host(sizes, torch.int32, "sizes"), self.mc.data_ptr(), self.mm.data_ptr(),
                                                 self.mf.data_ptr(), self.ind.data_ptr(), self.quant.data_ptr(),
                                                 self.status.data_ptr()), "cgic_session_decompress_host")
        return self.mc, self.mm, self.mf, self.ind, self.quant, self.status

    # ---- pinned-arena round trip: buffers owned by the session, one copy per direction and image range,
    #      ranges pipelined, replayed as a CUDA graph (cgic_session_arena / cgic_session_roundtrip_arena)
    _ARENA = dict(z=(0, torch.float32), m_c=(1, torch.int32), m_m=(2, torch.int32), m_f=(3, torch.int32), bytes=(4, torch.uint8),
                  sizes=(5, torch.int32), status=(6, torch.int32), sqerr=(7, torch.float64), ind=(8, torch.int64), quant=(9, torch.float32),
                  mc=(10, torch.int64), mm=(11, torch.int64), mf=(12, torch.int64), idx=(13, torch.int64), zq=(14, torch.float32),
                  # narrow wire (roundtrip_arena(narrow=True)): masks in as bytes; int16 indices and byte masks out
                  m_c8=(15, torch.uint8), m_m8=(16, torch.uint8), m_f8=(17, torch.uint8), ind16=(18, torch.int16),
                  mc8=(19, torch.uint8), mm8=(20, torch.uint8), mf8=(21, torch.uint8))

    def arena(self, parts: int = 8, slot: int = 0):
        """Fixes the number of image ranges and returns a list (one entry per range) of dicts of host tensor
        VIEWS into the session's pinned arenas: inputs z, m_c, m_m, m_f (fill them before roundtrip_arena) and
        outputs bytes, sizes, status, ind, quant, mc, mm, mf, idx, zq (valid after it), plus `images` = range.
        Narrow wire: inputs m_c8, m_m8, m_f8 (uint8) replace m_c, m_m, m_f; outputs ind16, mc8, mm8, mf8 replace
        ind, mc, mm, mf.  slot 1 is the second, independent arena set of submit_arena / wait_arena (same `parts`)."""
        check(lib().cgic_session_arena(self._s, int(parts)), "cgic_session_arena")
        h, w = self.h, self.w
        shapes = dict(z=lambda n: (n, 4, h, w), m_c=lambda n: (n, 1, h // 4, w // 4), m_m=lambda n: (n, 1, h // 2, w // 2),
                      m_f=lambda n: (n, 1, h, w), bytes=lambda n: (n, self.image_stride), sizes=lambda n: (n, 5), status=lambda n: (n,),
                      sqerr=lambda n: (1,), ind=lambda n: (n, h, w), quant=lambda n: (n, 4, h, w), mc=lambda n: (n, h // 4, w // 4),
                      mm=lambda n: (n, h // 2, w // 2), mf=lambda n: (n, h, w), idx=lambda n: (n * h * w,), zq=lambda n: (n, 4, h, w))
        shapes.update(m_c8=shapes["m_c"], m_m8=shapes["m_m"], m_f8=shapes["m_f"], ind16=shapes["ind"], mc8=shapes["mc"], mm8=shapes["mm"],
                      mf8=shapes["mf"])
        views = []
        p = 0
        while True:
            ptr, b0, nb = C.c_void_p(), C.c_int(), C.c_int()
            rc = lib().cgic_session_arena_slot_tensor(self._s, int(slot), 0, p, C.byref(ptr), C.byref(b0), C.byref(nb))
            if rc != 0:
                if p == 0:
                    check(rc, "cgic_session_arena_slot_tensor")
                break
            d = {"images": range(b0.value, b0.value + nb.value)}
            for name, (what, dtype) in self._ARENA.items():
                check(lib().cgic_session_arena_slot_tensor(self._s, int(slot), what, p, C.byref(ptr), None, None), "cgic_session_arena_slot_tensor")
                shape = shapes[name](nb.value)
                nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
                buf = (C.c_uint8 * nbytes).from_address(ptr.value)
                d[name] = torch.frombuffer(buf, dtype=dtype).view(*shape)
            views.append(d)
            p += 1
        if slot == 0:
            self._arena_views = views
        return views

    @staticmethod
    def _arena_flags(want_idx, want_zq, decoded_on_device, narrow) -> int:
        return (4 if decoded_on_device else int(want_idx) | (int(want_zq) << 1)) | (8 if narrow else 0)

    def roundtrip_arena(self, want_idx: bool = False, want_zq: bool = False, decoded_on_device: bool = False, narrow: bool = False) -> float:
        """CGIC.compress on the arena contents; returns sum((e - z)^2) over the batch.  decoded_on_device: the decoded
        tensors (ind, quant, mc, mm, mf) stay in HBM (fetch them with device_tensor); only bytes / sizes / status return.
        narrow: the masks are taken from m_c8 / m_m8 / m_f8 and the decoded tensors come back as ind16 / mc8 / mm8 / mf8
        (+ quant); the int64 tensors of the reference stay on the device."""
        sq = C.c_double()
        check(lib().cgic_session_roundtrip_arena(self._s, self._arena_flags(want_idx, want_zq, decoded_on_device, narrow), C.byref(sq)),
              "cgic_session_roundtrip_arena")
        return sq.value

    def submit_arena(self, slot: int, want_idx: bool = False, want_zq: bool = False, decoded_on_device: bool = False, narrow: bool = False) -> None:
        """Enqueues the round trip of arena set `slot` (0 or 1) and returns at once; wait_arena(slot) blocks until its
        results are in that set's host tensors.  Alternate the slots to overlap one batch's D2H with the next one's H2D."""
        check(lib().cgic_session_roundtrip_arena_submit(self._s, int(slot), self._arena_flags(want_idx, want_zq, decoded_on_device, narrow)),
              "cgic_session_roundtrip_arena_submit")

    def wait_arena(self, slot: int) -> float:
        sq = C.c_double()
        check(lib().cgic_session_roundtrip_arena_wait(self._s, int(slot), C.byref(sq)), "cgic_session_roundtrip_arena_wait")
        return sq.value

    def device_tensor(self, name: str, slot: int = 0) -> torch.Tensor:
        """Output tensor `name` (ind, quant, mc, mm, mf, bytes, sizes, status, idx, zq, ind16, mc8, mm8, mf8) of the last arena
        round trip of `slot` for the whole batch, gathered device-to-device into one CUDA tensor."""
        what, dtype = self._ARENA[name]
        B, h, w = self.B, self.h, self.w
        shape = dict(bytes=(B, self.image_stride), sizes=(B, 5), status=(B,), ind=(B, h, w), quant=(B, 4, h, w), mc=(B, h // 4, w // 4),
                     mm=(B, h // 2, w // 2), mf=(B, h, w), idx=(B * h * w,), zq=(B, 4, h, w), ind16=(B, h, w), mc8=(B, h // 4, w // 4),
                     mm8=(B, h // 2, w // 2), mf8=(B, h, w))[name]
        out = torch.empty(shape, dtype=dtype, device="cuda")
        check(lib().cgic_session_arena_slot_gather_device(self._s, int(slot), what, out.data_ptr()), "cgic_session_arena_slot_gather_device")
        return out

    def roundtrip(self, z, m_c, m_m, m_f, want_idx: bool = False):
        """CGIC.compress in one call: host z + masks -> (bytes, sizes, mc, mm, mf, ind, quant, status[, idx]) (pinned host)."""
        check(lib().cgic_session_roundtrip_host(self._s, self._host(z, torch.float32, "z"), self._host(m_c, torch.int32, "m_c"),
                                                self._host(m_m, torch.int32, "m_m"), self._host(m_f, torch.int32, "m_f"),
                                                self.bytes.data_ptr(), self.sizes.data_ptr(),
                                                self.idx.data_ptr() if want_idx else None, None, None,
                                                self.mc.data_ptr(), self.mm.data_ptr(), self.mf.data_ptr(), self.ind.data_ptr(),
                                                self.quant.data_ptr(), self.status.data_ptr()), "cgic_session_roundtrip_host")
        out = (self.bytes, self.sizes, self.mc, self.mm, self.mf, self.ind, self.quant, self.status)
        return out + (self.idx,) if want_idx else out
