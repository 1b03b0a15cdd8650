"""ctypes loader for libcgic_b200.so -- the C-ABI declared in include/cgic_b200.h.

There is no CPU fallback and no alternative backend: if the shared library is missing or fails
to load, importing the ops raises.  Build it with control-gic_b200/csrc/build.sh (or
`python -c "import __graft_entry__ as g; g.build()"`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CGIC_B200_LIB") or os.path.join(_HERE, "libcgic_b200.so")   # env override: kernel-tuning experiments only

OK, EINVAL, ENOMEM, ESPACE, ECUDA, EFORMAT = 0, -1, -2, -3, -4, -5

c_void_p, c_int, c_int64, c_size_t = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

# name -> (restype, argtypes).  Data pointers are passed as raw addresses (c_void_p).
_SIGNATURES = {
    "cgic_abi_version": (c_int, []),
    "cgic_last_error": (C.c_char_p, []),
    "cgic_tune": (c_int, [C.c_char_p, c_int]),
    "cgic_prof_enable": (c_int, [c_int]),
    "cgic_prof_report": (c_int, [C.c_char_p, c_int]),
    "cgic_huff_build": (c_int, [c_void_p, c_void_p, c_int, C.POINTER(c_void_p)]),
    "cgic_huff_free": (None, [c_void_p]),
    "cgic_huff_num_symbols": (c_int, [c_void_p]),
    "cgic_huff_max_len": (c_int, [c_void_p]),
    "cgic_huff_code_len": (c_int, [c_void_p, c_int]),
    "cgic_huff_code": (c_int, [c_void_p, c_int, C.c_char_p, c_int]),
    "cgic_huff_upload": (c_int, [c_void_p]),
    "cgic_vq_workspace_bytes": (c_size_t, [c_int64]),
    "cgic_vq_assign": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_size_t, c_void_p]),
    "cgic_codebook_create": (c_int, [c_int, C.POINTER(c_void_p)]),
    "cgic_codebook_free": (None, [c_void_p]),
    "cgic_codebook_update": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cgic_codebook_stats_host": (c_int, [c_void_p, c_void_p]),
    "cgic_codebook_check": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cgic_codebook_is_stale": (c_int, [c_void_p]),
    "cgic_vq_assign_indexed": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    "cgic_vq_count": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "cgic_entropy_maps": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cgic_entropy_route_workspace_bytes": (c_size_t, [c_int]),
    "cgic_entropy_route": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, C.c_float, C.c_float,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cgic_route_mix": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "cgic_router_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cgic_router": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cgic_mask_mix": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p]),
    "cgic_decoder_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p]),
    "cgic_spatial_norm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cgic_spatial_norm": (c_int, [c_void_p] * 8 + [c_int] * 8 + [C.c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cgic_pack_layout": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cgic_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                          c_void_p]),
    "cgic_pack_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cgic_pack_ws": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_void_p]),
    "cgic_encode_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cgic_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cgic_unpack_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "cgic_unpack": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cgic_huff_stream_capacity": (c_int64, [c_void_p, c_int64]),
    "cgic_huff_encode": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "cgic_huff_decode": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "cgic_bits_encode": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "cgic_bits_decode": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "cgic_session_create": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, C.POINTER(c_void_p)]),
    "cgic_session_destroy": (None, [c_void_p]),
    "cgic_session_image_stride": (c_int64, [c_void_p]),
    "cgic_session_set_pipeline": (c_int, [c_void_p, c_int]),
    "cgic_session_compress_host": (c_int, [c_void_p] * 10),
    "cgic_session_decompress_host": (c_int, [c_void_p] * 9),
    "cgic_session_roundtrip_host": (c_int, [c_void_p] * 16),
    "cgic_session_arena": (c_int, [c_void_p, c_int]),
    "cgic_session_arena_tensor": (c_int, [c_void_p, c_int, c_int, C.POINTER(c_void_p), C.POINTER(c_int), C.POINTER(c_int)]),
    "cgic_session_roundtrip_arena": (c_int, [c_void_p, c_int, c_void_p]),
    "cgic_session_arena_gather_device": (c_int, [c_void_p, c_int, c_void_p]),
    "cgic_session_arena_slot_tensor": (c_int, [c_void_p, c_int, c_int, c_int, C.POINTER(c_void_p), C.POINTER(c_int), C.POINTER(c_int)]),
    "cgic_session_roundtrip_arena_submit": (c_int, [c_void_p, c_int, c_int]),
    "cgic_session_roundtrip_arena_wait": (c_int, [c_void_p, c_int, c_void_p]),
    "cgic_session_arena_slot_gather_device": (c_int, [c_void_p, c_int, c_int, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class CgicError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str):
        self.code = code
        super().__init__(f"{where} failed with code {code}: {detail}")


def lib() -> C.CDLL:
    """The loaded library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run control-gic_b200/csrc/build.sh).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.cgic_abi_version() != 1:
            raise RuntimeError(f"ABI version mismatch: {handle.cgic_abi_version()} != 1")
        _lib = handle
    return _lib


def check(rc: int, where: str) -> None:
    if rc != OK:
        raise CgicError(rc, where, lib().cgic_last_error().decode(errors="replace"))
