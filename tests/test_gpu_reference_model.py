"""The drop-in boundary with the REAL reference CNNs: `cgic_b200.CGIC(**yaml_params)` builds the reference's own Encoder /
Decoder (ch = 32 so that it runs in a blink) through `ReferenceEncoderHeads`, and `compress()` / `inference.main()` run end
to end on the GPU.  The reference package comes from the git-ignored copy baseline/_ref (made by __graft_entry__.build(),
shipped by gpurun); without it the tests skip.  Parity: everything between the CNNs is checked against the oracle on the
tensors observed at the boundary (entropy maps, heads, quant_conv output), the files byte for byte."""
import os
import sys

import numpy as np
import pytest
import torch
import yaml

from conftest import ROOT, STREAMS

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "CGIC")), reason="baseline/_ref (the reference copy) is not present")


def _params(c_ratio=0.1, m_ratio=0.8):
    cfg = yaml.safe_load(open(os.path.join(REF, "configs", "config_inference.yaml")))["model"]["params"]
    cfg.update(ckpt_path=None, lossconfig=None, ema_decay=None)
    cfg["ddconfig"]["ch"] = 32
    cfg["ddconfig"]["router_config"]["params"] = dict(coarse_grain_ratio=c_ratio, medium_grain_ratio=m_ratio)
    return cfg


def _model(cg, c_ratio=0.1, m_ratio=0.8, seed=0):
    torch.manual_seed(seed)
    model = cg.CGIC(**_params(c_ratio, m_ratio)).cuda().eval()          # no encoder= / decoder=: built from ddconfig like the reference
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        model.quantize.embedding.weight.copy_((torch.randn(1024, 4, generator=g) * 0.5).cuda())   # spread the codes over O(1) latents
        cnt = (-torch.log(torch.rand(1024, generator=g)) * 1000).floor()
        for i in range(1024):
            model.quantize.embedding_counter[str(i)].data.fill_(float(cnt[i]))
    return model


@needs_ref
@pytest.mark.parametrize("H,W,c,m", [(64, 64, 0.1, 0.8), (96, 128, 0.3, 0.6), (64, 64, 0.0, 0.0)])
def test_yaml_built_model_compress_against_oracle(tmp_path, H, W, c, m):
    import cgic_b200 as cg
    from oracle import oracle as orc
    model = _model(cg, c, m)
    assert type(model.encoder).__name__ == "ReferenceEncoderHeads" and type(model.encoder.encoder).__module__ == "CGIC.modules.vqvae.vqvae_blocks"
    assert type(model.decoder).__module__ == "CGIC.modules.vqvae.decoder"
    x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(H + W)).cuda()
    cap = {}
    hooks = [model.encoder.encoder.conv_out_coarse.register_forward_hook(lambda mod, i, o: cap.__setitem__("hc", o.detach().clone())),
             model.encoder.encoder.conv_out.register_forward_hook(lambda mod, i, o: cap.__setitem__("hm", o.detach().clone())),
             model.encoder.encoder.conv_out_fine.register_forward_hook(lambda mod, i, o: cap.__setitem__("hf", o.detach().clone())),
             model.quant_conv.register_forward_hook(lambda mod, i, o: cap.update(h=i[0].detach().clone(), z=o.detach().clone())),
             model.decoder.register_forward_hook(lambda mod, i, o: cap.update(quant_dec=i[1].detach().clone(), masks_dec=[t.clone() for t in i[2]]))]
    # the reference's OWN HuffmanCoding object, built the way inference.py:137-139 does, is accepted
    sys.path.insert(0, REF)
    try:
        from CGIC.tools.indices_coding import HuffmanCoding as RefHuffman
        from CGIC.tools.mask_coding import BinaryCoding as RefBinary
    finally:
        sys.path.pop(0)
    h_ref = RefHuffman(model.quantize.embedding_counter)
    with torch.no_grad():
        dec, bpp, pm = model.compress(x, str(tmp_path), h_ref, RefBinary(), False)
    for hk in hooks:
        hk.remove()
    h, w = H // 4, W // 4
    # boundary tensors -> oracle
    e8, e16 = cg.entropy_pair(x)
    omc, omm, omf, mode = orc.router(e16.cpu().numpy(), e8.cpu().numpy(), c, m)
    mix = orc.mask_mix(cap["hc"].cpu().numpy(), cap["hm"].cpu().numpy(), cap["hf"].cpu().numpy(), omc, omm, omf)
    assert np.array_equal(mix.view(np.uint32), cap["h"].cpu().numpy().view(np.uint32))
    E = model.quantize.embedding.weight.detach().cpu().numpy()
    ozq, oloss, oidx = orc.vq_assign(cap["z"].cpu().numpy(), E)
    counts = [int(model.quantize.embedding_counter[str(i)].item()) for i in range(1024)]
    ot = orc.huff_build(np.asarray(counts, np.int64), orc.lexicographic_order(1024))
    assert ot.codes == h_ref.codes
    streams = orc.pack_image(ot, oidx.reshape(h, w), omc[0, 0], omm[0, 0], omf[0, 0], mode)
    for s, n in enumerate(STREAMS):
        fn = tmp_path / (n + ".bin")
        assert (fn.read_bytes() if fn.exists() else b"") == streams[s], n
    assert bpp == orc.bpp_of(streams, H, W) and pm is None
    umc, umm, umf, uind, uq = orc.unpack_image(ot, streams, h, w, mode, E)
    assert np.array_equal(cap["quant_dec"].cpu().numpy()[0], uq)
    for o, r in zip((umc, umm, umf), cap["masks_dec"]):
        assert np.array_equal(r.cpu().numpy()[0, 0].astype(np.int64), o)
    # the decoder really ran on those tensors
    with torch.no_grad():
        again = model.decode(cap["quant_dec"], cap["masks_dec"])
    assert torch.equal(dec, again) and tuple(dec.shape) == (1, 3, H, W)
    assert model.threshold_adjacent is not None and tuple(model.threshold_adjacent.shape) == (1, 2)


@needs_ref
def test_cli_main_writes_the_reference_outputs(tmp_path, monkeypatch):
    """inference.main(): the reference's flags, bpp.txt lines, reconstructed/{k:03d}_{bpp:05f}.png naming, the five files."""
    import cgic_b200 as cg
    from PIL import Image
    img_dir, out_dir = tmp_path / "in", tmp_path / "out"
    img_dir.mkdir()
    rng = np.random.default_rng(1)
    for name, (hh, ww) in {"a.png": (70, 85), "b.png": (64, 64)}.items():
        Image.fromarray(rng.integers(0, 255, (hh, ww, 3), dtype=np.uint8)).save(img_dir / name)
    cfg = {"model": {"target": "CGIC.models.model.CGIC", "params": _params(0.1, 0.8)}}
    cfg_path = tmp_path / "config_inference.yaml"
    cfg_path.write_text(yaml.safe_dump(cfg))
    torch.manual_seed(0)
    avg = cg.inference.main(["-i", str(img_dir), "-o", str(out_dir), "-n", "0", "-r", "0", "2"], config_path=str(cfg_path))
    lines = (out_dir / "bpp.txt").read_text().splitlines()
    assert lines[0].startswith("image: 0 \t bpp: ") and lines[1].startswith("image: 1 \t bpp: ") and lines[2].startswith("Bpp Average: ")
    bpps = [float(l.split("bpp: ")[1]) for l in lines[:2]]
    assert abs(avg - sum(bpps) / 2) < 1e-12
    pngs = sorted(p.name for p in (out_dir / "reconstructed").iterdir())
    assert pngs == [f"{k:03d}_{b:05f}.png" for k, b in enumerate(bpps)]
    assert Image.open(out_dir / "reconstructed" / pngs[0]).size == (80, 64)          # centre crop to multiples of 16 (W, H)
    for n in STREAMS:
        assert (out_dir / (n + ".bin")).exists()
    total = sum((out_dir / (n + ".bin")).stat().st_size for n in STREAMS)
    assert total * 8 / (64 * 64) == bpps[1]                                            # the files on disk are the last image's
