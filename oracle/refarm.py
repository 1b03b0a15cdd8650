"""The timed CPU arm of bench.py: the UNMODIFIED reference's own classes on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product never does).

The reference is pure Python, so a copy of its tree travels to the GPU box in the git-ignored `baseline/_ref/`
(`__graft_entry__.build()` makes the copy whenever /root/reference is present).  When that copy is importable this
module drives the reference's `VectorQuantize2.forward` (CGIC/modules/vqvae/quantize.py:69-98),
`HuffmanCoding.compress / decompress_string` (CGIC/tools/indices_coding.py:113-168) and `BinaryCoding.compress /
decompress_string` (CGIC/tools/mask_coding.py:40-96) -- kind "reference".  The lines of `CGIC.compress` that sit
between those calls (selection model.py:217-221, file sizes / bpp :226-233, re-assembly :278-293, gather :391-392) are
inline in a method that also runs the CNNs, so they are restated here one for one.  Without the copy the
reference-shaped port of oracle/refport.py is used -- kind "port".
"""
from __future__ import annotations

import os
import sys
import types

import torch
from torch import nn

from . import refport

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.environ.get("CGIC_REFERENCE_COPY", os.path.join(ROOT, "baseline", "_ref"))

_ref = None


def reference_modules():
    """(VectorQuantize2, HuffmanCoding, BinaryCoding) of the unmodified reference, or None when baseline/_ref is absent.
    Shims: a stub `pytorch_lightning` (not installed here) and a no-op `ParameterDict.cuda`, so that the
    quantiser's counters stay on the host like everything else this arm touches."""
    global _ref
    if _ref is None:
        _ref = False
        if os.path.isdir(os.path.join(REF_DIR, "CGIC")):
            try:
                if "pytorch_lightning" not in sys.modules:
                    pl = types.ModuleType("pytorch_lightning")
                    pl.LightningModule = nn.Module
                    pl.LightningDataModule = object
                    sys.modules["pytorch_lightning"] = pl
                sys.path.insert(0, REF_DIR)
                from CGIC.modules.vqvae.quantize import VectorQuantize2
                from CGIC.tools.indices_coding import HuffmanCoding
                from CGIC.tools.mask_coding import BinaryCoding
                _ref = (VectorQuantize2, HuffmanCoding, BinaryCoding)
            except Exception as e:  # an incomplete copy: fall back to the port, say why
                print(f"[refarm] baseline/_ref is not importable ({type(e).__name__}: {e}); using the port", file=sys.stderr)
            finally:
                if sys.path and sys.path[0] == REF_DIR:
                    sys.path.pop(0)
    return _ref or None


class Arm:
    """One image at a time, B == 1 (the only batch size CGIC.compress supports).  roundtrip(z, masks, workdir) ->
    (ind [1,h,w], bpp, ind_decompress [1,h,w], quant [1,4,h,w], sizes[5]); the five files stay in workdir."""

    def __init__(self, codebook: torch.Tensor, counts, order):
        mods = reference_modules()
        self.kind = "reference" if mods else "port"
        self.codebook = codebook
        if mods:
            VQ, Huff, Bits = mods
            cuda_shim = nn.ParameterDict.cuda
            nn.ParameterDict.cuda = lambda self_, device=None: self_
            try:
                self.vq = VQ(codebook.shape[0], codebook.shape[1], 0.25).eval()
            finally:
                nn.ParameterDict.cuda = cuda_shim
            with torch.no_grad():
                self.vq.embedding.weight.copy_(codebook)
                for i, c in enumerate(counts):
                    self.vq.embedding_counter[str(i)].data.fill_(float(c))
            assert [int(k) for k in self.vq.embedding_counter.keys()] == [int(s) for s in order]
            self.h_indices = Huff(self.vq.embedding_counter)       # inference.py:137-139
            self.h_mask = Bits()
        else:
            self.table = refport.huffman_codes([int(c) for c in counts], [int(s) for s in order])
            self.reverse = {v: k for k, v in self.table.items()}

    NAMES = ("indices_coarse", "indices_medium", "indices_fine", "mask_coarse", "mask_medium")

    def roundtrip(self, z_nchw: torch.Tensor, masks, workdir: str):
        if self.kind == "port":
            return refport.roundtrip_mode0(z_nchw, self.codebook, masks, self.table, self.reverse, workdir)
        with torch.no_grad():
            quant, emb_loss, ind = self.vq(z_nchw)                                        # quantize.py:69-98
        h, w = quant.shape[-2:]
        grain_mask = masks
        # model.py:217-221
        ind = ind.view(-1, h, w)
        ind_coarse = ind[:, ::4, ::4][grain_mask[0][0] == 1]
        ind_medium = ind[:, ::2, ::2][grain_mask[1][0] == 1]
        ind_fine = ind[grain_mask[2][0] == 1]
        paths = [os.path.join(workdir, n + ".bin") for n in self.NAMES]
        # model.py:226-233 (mode 0)
        self.h_indices.compress(ind_coarse, paths[0])
        self.h_indices.compress(ind_medium, paths[1])
        self.h_indices.compress(ind_fine, paths[2])
        self.h_mask.compress(grain_mask[0].flatten(), paths[3])
        self.h_mask.compress(grain_mask[1].flatten(), paths[4])
        sizes = [os.path.getsize(p) for p in paths]
        bpp = sum(sizes) * 8 / (16 * h * w)
        # model.py:269-293
        ind_coarse_d = self.h_indices.decompress_string(paths[0])
        ind_medium_d = self.h_indices.decompress_string(paths[1])
        ind_fine_d = self.h_indices.decompress_string(paths[2])
        mask_coarse = torch.tensor(self.h_mask.decompress_string(paths[3])).view(1, h // 4, w // 4)
        mask_medium = torch.tensor(self.h_mask.decompress_string(paths[4])).view(1, h // 2, w // 2)
        mask_fine = 1 - mask_medium.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2) \
            - mask_coarse.repeat_interleave(4, dim=-1).repeat_interleave(4, dim=-2)
        if ind_coarse_d is None:
            mask_coarse = torch.zeros_like(mask_coarse)
        else:
            mask_coarse[mask_coarse == 1] = torch.tensor(ind_coarse_d)
        if ind_medium_d is None:
            mask_medium = torch.zeros_like(mask_medium)
        else:
            mask_medium[mask_medium == 1] = torch.tensor(ind_medium_d)
        mask_fine[mask_fine == 1] = torch.tensor(ind_fine_d)
        ind_decompress = mask_fine + mask_medium.repeat_interleave(2, dim=-1).repeat_interleave(2, dim=-2) \
            + mask_coarse.repeat_interleave(4, dim=-1).repeat_interleave(4, dim=-2)
        # model.py:391-392
        with torch.no_grad():
            quant_d = self.vq.embedding(ind_decompress.flatten()).view(1, h, w, -1).permute(0, 3, 1, 2)
        return ind, bpp, ind_decompress, quant_d, sizes

    def files(self, workdir: str):
        """The five files the last roundtrip() left in workdir, as bytes (b'' for an absent / empty file)."""
        out = []
        for n in self.NAMES:
            p = os.path.join(workdir, n + ".bin")
            out.append(open(p, "rb").read() if os.path.exists(p) else b"")
        return out
