#!/usr/bin/env python
"""Golden vectors for SpatialNorm, generated from the UNMODIFIED reference class
(CGIC/modules/vqvae/decoder.py:34-56) on CPU.  Build container only (needs /root/reference).

    python tests/golden/make_spatial_norm_golden.py     # rewrites tests/golden/spatial_norm.npz
Cases: (name, B, C, H, W, Cz, hz, wz, add_conv, offset) -- integer and non-integer up-sampling factors, W % 4 != 0,
a feature map with a large mean (offset) and the add_conv variant.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CGIC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from CGIC.modules.vqvae.decoder import Normalize  # noqa: E402

CASES = [
    ("up4", 2, 32, 16, 16, 4, 4, 4, False, 0.0),
    ("up1", 1, 64, 8, 12, 4, 8, 12, False, 0.0),
    ("ragged", 2, 32, 12, 10, 4, 5, 3, False, 0.0),
    ("offset", 1, 32, 8, 8, 4, 2, 2, False, 100.0),
    ("add_conv", 1, 32, 8, 8, 4, 4, 4, True, 0.0),
]


def main():
    out = {}
    for name, B, Cc, H, W, Cz, hz, wz, add_conv, offset in CASES:
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        m = Normalize(Cc, Cz, add_conv)
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
        f = torch.randn(B, Cc, H, W, generator=g) * 1.7 + offset
        zq = torch.randn(B, Cz, hz, wz, generator=g)
        with torch.no_grad():
            y = m(f, zq)
        out[f"{name}.f"] = f.numpy()
        out[f"{name}.zq"] = zq.numpy()
        out[f"{name}.out"] = y.numpy()
        for k, v in m.state_dict().items():
            out[f"{name}.sd.{k}"] = v.numpy()
    out["cases"] = np.array([c[0] for c in CASES])
    out["add_conv"] = np.array([c[8] for c in CASES])
    np.savez_compressed(os.path.join(HERE, "spatial_norm.npz"), **out)
    print("wrote spatial_norm.npz:", {c[0]: out[c[0] + ".out"].shape for c in CASES})


if __name__ == "__main__":
    main()
