"""TripleGrainFixedEntropyRouter -- drop-in for CGIC/modules/vqvae/RouterTriple.py:7-96.

forward(x_entropy_p16, x_entropy_p8) -> (mask, gate, [coarse, medium, fine ratio], mode) with
mask = three int32 tensors [B,1,.,.] and gate fp32 [B,1,h,3w], thresholds taken over the whole
batch exactly like the reference.  `per_image=True` (an extension) thresholds every image on
its own, i.e. B independent B == 1 calls in one launch.
"""
from __future__ import annotations

from torch import nn

from . import ops


class TripleGrainFixedEntropyRouter(nn.Module):
    def __init__(self, coarse_grain_ratio, medium_grain_ratio, per_image: bool = False):
        super().__init__()
        self.coarse_grain_ratio = coarse_grain_ratio
        self.medium_grain_ratio = medium_grain_ratio
        self.fine_grain_ratio = 1 - coarse_grain_ratio - medium_grain_ratio           # RouterTriple.py:13
        self.per_image = per_image

    def forward(self, x_entropy_p16, x_entropy_p8, want_gate: bool = True):
        m_c, m_m, m_f, gate, mode = ops.router(x_entropy_p16, x_entropy_p8, self.coarse_grain_ratio,
                                               self.medium_grain_ratio, per_image=self.per_image, want_gate=want_gate)
        return [m_c, m_m, m_f], gate, [self.coarse_grain_ratio, self.medium_grain_ratio, self.fine_grain_ratio], mode
