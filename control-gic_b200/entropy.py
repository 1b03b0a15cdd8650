"""Entropy -- drop-in for CGIC/models/model.py:433-483 (per-patch soft-histogram entropy).

forward(inputs [B,3,H,W]) -> [B, H/p, W/p] for p in {8, 16}.  `entropy_pair` computes both patch
sizes in one pass over the image (what CGIC.encode needs, model.py:100-101).
"""
from __future__ import annotations

from torch import nn

from . import ops


class Entropy(nn.Sequential):
    def __init__(self, patch_size):
        super().__init__()
        if patch_size not in (8, 16):
            raise ValueError("the B200 entropy kernel implements patch sizes 8 and 16 (entropy_patch_size of the reference configs)")
        self.psize = patch_size

    def forward(self, inputs):
        e8, e16 = ops.entropy_maps(inputs, want8=self.psize == 8, want16=self.psize == 16)
        return e8 if self.psize == 8 else e16


def entropy_pair(inputs):
    """(x_entropy_p8, x_entropy_p16) from one read of the image."""
    return ops.entropy_maps(inputs)
