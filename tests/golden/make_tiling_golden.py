#!/usr/bin/env python
"""Golden vectors for the tiling driver, generated from the UNMODIFIED reference
(inference_high_resolution.py:112-173, get_parser of both CLIs).  Build container only (needs
/root/reference); shims: stub `pytorch_lightning` / `omegaconf` modules (neither is installed).

    python tests/golden/make_tiling_golden.py     # rewrites tests/golden/tiling_kats.json
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CGIC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
_pl = types.ModuleType("pytorch_lightning")
_pl.LightningModule = nn.Module
_pl.LightningDataModule = object
sys.modules["pytorch_lightning"] = _pl
_oc = types.ModuleType("omegaconf")
_oc.OmegaConf = object
sys.modules["omegaconf"] = _oc
nn.ParameterDict.cuda = lambda self, device=None: self

import inference as ref_inf  # noqa: E402
import inference_high_resolution as ref_hr  # noqa: E402

SHAPES = [(1356, 2040), (1344, 2032), (768, 768), (512, 768), (2048, 1536), (100, 3000), (769, 767), (16, 16), (1537, 800)]


def main():
    out = {"shapes": []}
    for H, W in SHAPES:
        pad, unpad = ref_hr.compute_padding(H, W, min_div=2 ** 4)
        Hp, Wp = H + pad[2] + pad[3], W + pad[0] + pad[1]
        x = torch.zeros(1, 3, Hp, Wp)
        h_list, w_list, th, tw = ref_hr.nonoverlapping_grid_indices(x)
        out["shapes"].append(dict(H=H, W=W, pad=list(pad), unpad=list(unpad), h_list=h_list, w_list=w_list, tile_h=th, tile_w=tw))
    out["weights"] = []
    for tw, th in [(768, 768), (496, 576), (16, 32)]:
        wts = ref_hr._gaussian_weights(tw, th, 1, "cpu")
        out["weights"].append(dict(tile_w=tw, tile_h=th, shape=list(wts.shape), dtype=str(wts.dtype),
                                   sha256=hashlib.sha256(np.ascontiguousarray(wts.numpy()).tobytes()).hexdigest(),
                                   corner=float(wts[0, 0, 0, 0]), centre=float(wts[0, 0, th // 2, tw // 2])))
    out["parser"] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(ref_inf.get_parser().parse_args([])).items()}
    out["parser_hr"] = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(ref_hr.get_parser().parse_args([])).items()}
    with open(os.path.join(HERE, "tiling_kats.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote tiling_kats.json")


if __name__ == "__main__":
    main()
