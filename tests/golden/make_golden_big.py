#!/usr/bin/env python
"""Golden fixtures at BASELINE.json's own sizes, generated from the UNMODIFIED reference.

    python tests/golden/make_golden_big.py        # writes tests/golden/big_*.npz

Same method as make_golden.py (whose shims and model builder it imports): the reference's
`CGIC.compress` (CGIC/models/model.py:206-401) runs on one image, ch=32 CNNs, and forward hooks
record the hot path's tensors.  Cases:

    c1_256         256x256, ratio (0.1, 0.8)         BASELINE configs[0] (and one image of configs[1])
    c3_r0.3-0.6    512x768 (Kodak shape), (0.3, 0.6)  BASELINE configs[2]
    c3_r0.1-0.8    512x768, (0.1, 0.8)
    c3_r0.05-0.05  512x768, (0.05, 0.05)
    c5_768x496     768x496 tile, (0.1, 0.8)           a tile of BASELINE configs[4] (2032x1344)
    c5_576x496     576x496 tile, (0.1, 0.8)

To keep the fixtures small only the path's inputs at each boundary and compact outputs are stored:
entropy maps (router input), z (VQ input = quant_conv output), masks as packed bits, indices as
int16, the five files byte for byte, sha256 of z_q and of quant_decompress, loss, bpp.  c1_256 also
stores the image and the three encoder heads, so Entropy and the mask-mix are pinned at 256x256.
While generating, the oracle is asserted equal to the reference on every case.
"""
import hashlib
import os
import sys
import tempfile
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the shims, imports the reference)
from make_golden import BinaryCoding, HuffmanCoding, STREAMS, orc  # noqa: E402


def sha(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def big_case(tag, H, W, c_ratio, m_ratio, counts, full=False):
    model = mg.build_model(c_ratio, m_ratio, counts)
    g = torch.Generator().manual_seed(zlib.crc32(("big_" + tag).encode()))
    x = torch.rand(1, 3, H, W, generator=g)
    cap = {}
    hooks = [
        model.entropy_calculation_p8.register_forward_hook(lambda m, i, o: cap.__setitem__("e8", o.clone())),
        model.entropy_calculation_p16.register_forward_hook(lambda m, i, o: cap.__setitem__("e16", o.clone())),
        model.encoder.conv_out_coarse.register_forward_hook(lambda m, i, o: cap.__setitem__("hc", o.clone())),
        model.encoder.conv_out.register_forward_hook(lambda m, i, o: cap.__setitem__("hm", o.clone())),
        model.encoder.conv_out_fine.register_forward_hook(lambda m, i, o: cap.__setitem__("hf", o.clone())),
        model.encoder.register_forward_hook(lambda m, i, o: cap.update(h=o["h"].clone(), mask=[t.clone() for t in o["mask"]],
                                                                       mode=o["compression_mode"])),
        model.quantize.register_forward_hook(lambda m, i, o: cap.update(z=i[0].clone(), zq=o[0].clone(), loss=o[1].clone(), ind=o[2].clone())),
        model.quantize.embedding.register_forward_hook(lambda m, i, o: cap.setdefault("emb_calls", []).append(i[0].clone())),
        model.decoder.register_forward_hook(lambda m, i, o: cap.update(quant_dec=i[1].clone(), mask_dec=[t.clone() for t in i[2]])),
    ]
    hs = HuffmanCoding(model.quantize.embedding_counter)
    hb = BinaryCoding()
    d = tempfile.mkdtemp()
    with torch.no_grad():
        dec, bpp, _ = model.compress(x, d, hs, hb, False)
    for hk in hooks:
        hk.remove()
    files = {n: (open(os.path.join(d, n + ".bin"), "rb").read() if os.path.exists(os.path.join(d, n + ".bin")) else b"") for n in STREAMS}
    mode = cap["mode"]
    h, w = H // 4, W // 4
    ind_dec = cap["emb_calls"][-1].view(1, h, w)
    E = model.quantize.embedding.weight.detach()

    # ---- the oracle against the reference on this case
    order = [int(k) for k in model.quantize.embedding_counter.keys()]
    t = orc.huff_build(np.asarray(counts, np.int64), order)
    assert t.codes == hs.codes
    ozq, oloss, oidx = orc.vq_assign(cap["z"].numpy(), E.numpy())
    assert np.array_equal(oidx, cap["ind"].numpy()), tag
    assert np.array_equal(ozq.view(np.uint32), cap["zq"].numpy().view(np.uint32)), tag
    omix = orc.mask_mix(cap["hc"].numpy(), cap["hm"].numpy(), cap["hf"].numpy(), *(m.numpy() for m in cap["mask"]))
    assert np.array_equal(omix.view(np.uint32), cap["h"].numpy().view(np.uint32)), tag
    omc, omm, omf, omode = orc.router(cap["e16"].numpy(), cap["e8"].numpy(), c_ratio, m_ratio)
    assert omode == mode
    for o, r in zip((omc, omm, omf), cap["mask"]):
        assert np.array_equal(o, r.numpy()), tag
    streams = orc.pack_image(t, cap["ind"].view(h, w).numpy(), omc[0, 0], omm[0, 0], omf[0, 0], mode)
    for s, n in enumerate(STREAMS):
        assert streams[s] == files[n], (tag, n)
    assert orc.bpp_of(streams, H, W) == bpp
    umc, umm, umf, uind, uq = orc.unpack_image(t, streams, h, w, mode, E.numpy())
    assert np.array_equal(uind, ind_dec[0].numpy()) and np.array_equal(uq, cap["quant_dec"][0].numpy()), tag
    for o, r in zip((umc, umm, umf), cap["mask_dec"]):
        assert np.array_equal(o, r[0, 0].numpy().astype(np.int64)), tag

    out = dict(shape=np.asarray([H, W], np.int32), e8=cap["e8"].numpy(), e16=cap["e16"].numpy(), z=cap["z"].numpy(),
               codebook=E.numpy(), counts=np.asarray(counts, np.int64), order=np.asarray(order, np.int32),
               ratios=np.asarray([c_ratio, m_ratio], np.float64), mode=np.int32(mode),
               ind=cap["ind"].numpy().astype(np.int16), zq_sha=sha(cap["zq"].numpy()), loss=np.float32(cap["loss"]),
               ind_dec=ind_dec.numpy().astype(np.int16), quant_dec_sha=sha(cap["quant_dec"].numpy()), bpp=np.float64(bpp))
    for lvl in range(3):
        out[f"mask{lvl}_bits"] = np.packbits(cap["mask"][lvl].numpy().astype(np.uint8).ravel())
        out[f"mask_dec{lvl}_bits"] = np.packbits(cap["mask_dec"][lvl].numpy().astype(np.uint8).ravel())
    for n in STREAMS:
        out["file_" + n] = np.frombuffer(files[n], np.uint8)
    if full:
        out.update(x=x.numpy(), hc=cap["hc"].numpy(), hm=cap["hm"].numpy(), hf=cap["hf"].numpy(), h_sha=sha(cap["h"].numpy()))
    n = [int(cap["mask"][lvl].sum()) for lvl in range(3)]
    print(f"  big {tag}: {H}x{W} ratio=({c_ratio},{m_ratio}) mode={mode} sizes={[len(files[s]) for s in STREAMS]} bpp={bpp:.5f} n={n}",
          flush=True)
    return out


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    _, cnt5 = mg.kat5_counts()
    cnt5 = cnt5.long().tolist()
    cases = [
        ("c1_256", 256, 256, 0.1, 0.8, True),
        ("c3_r0.3-0.6", 512, 768, 0.3, 0.6, False),
        ("c3_r0.1-0.8", 512, 768, 0.1, 0.8, False),
        ("c3_r0.05-0.05", 512, 768, 0.05, 0.05, False),
        ("c5_768x496", 768, 496, 0.1, 0.8, False),
        ("c5_576x496", 576, 496, 0.1, 0.8, False),
    ]
    only = sys.argv[1:]
    for tag, H, W, c, m, full in cases:
        if only and tag not in only:
            continue
        np.savez_compressed(os.path.join(HERE, f"big_{tag}.npz"), **big_case(tag, H, W, c, m, cnt5, full))
    print("done")


if __name__ == "__main__":
    main()
