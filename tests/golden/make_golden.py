#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the GPU box and the test-suite read
the fixtures it wrote, never the reference.  Two shims make the reference importable on a
CPU-only host without pytorch_lightning: a stub `pytorch_lightning` module and a
no-op `nn.ParameterDict.cuda`.  The CNN encoder/decoder are instantiated at ch=32 instead of
128 (ddconfig.ch) so that `CGIC.compress` runs in ~0.1 s; the hot path (entropy -> router ->
mask-mix -> quant_conv -> VQ -> select -> Huffman/binary pack -> unpack -> re-assembly ->
gather) is exactly the reference's code, observed through forward hooks.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz, *.json

While generating, it also asserts that oracle/refport.py (the port bench.py times) and
oracle/cgic_oracle.c agree with the reference on every case.
"""
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch
import yaml
import zlib
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CGIC_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

_pl = types.ModuleType("pytorch_lightning")
_pl.LightningModule = nn.Module
_pl.LightningDataModule = object
sys.modules["pytorch_lightning"] = _pl
nn.ParameterDict.cuda = lambda self, device=None: self

from CGIC.models.model import CGIC, Entropy  # noqa: E402
from CGIC.modules.vqvae.quantize import VectorQuantize2  # noqa: E402
from CGIC.modules.vqvae.RouterTriple import TripleGrainFixedEntropyRouter  # noqa: E402
from CGIC.tools.indices_coding import HuffmanCoding  # noqa: E402
from CGIC.tools.mask_coding import BinaryCoding  # noqa: E402

from oracle import oracle as orc  # noqa: E402
from oracle import refport  # noqa: E402

STREAMS = ("indices_coarse", "indices_medium", "indices_fine", "mask_coarse", "mask_medium")


def freq_dict(counts):
    return {str(i): torch.tensor([float(c)]) for i, c in enumerate(counts)}


def digest(codes, K):
    return hashlib.sha256("".join(f"{i}:{codes[i]};" for i in range(K)).encode()).hexdigest()


def kat5_counts():
    g = torch.Generator().manual_seed(1234)
    cnt = (-torch.log(torch.rand(1024, generator=g)) * 1000).floor()
    return g, cnt


# ------------------------------------------------------------------------------------------
def huffman_kats():
    out = {}
    tmp = tempfile.mkdtemp()

    def enc(h, syms):
        p = os.path.join(tmp, "s.bin")
        h.compress(torch.tensor(syms, dtype=torch.int64), p)
        data = open(p, "rb").read()
        back = h.decompress_string(p)
        assert (back is None and not syms) or back == list(syms)
        return data.hex()

    for name, freq, syms in (
        ("kat1", [5, 9, 12, 13, 16, 45, 0, 0], [5, 0, 1, 6, 7, 5, 5, 2, 3, 4]),
        ("kat1b", [5, 9, 12, 13, 16, 45, 0, 0], [5] * 8),
        ("kat1c", [5, 9, 12, 13, 16, 45, 0, 0], []),
        ("kat2", [1] * 8, [0, 1, 2, 3, 4, 5, 6, 7]),
        ("kat3", [0] * 16, list(range(16))),
        ("kat_desc", list(range(40, 0, -1)), [0, 39, 20, 5, 5, 7]),
        ("kat_pow2", [2 ** i for i in range(20)], [0, 1, 2, 19, 18, 0]),
    ):
        h = HuffmanCoding(freq_dict(freq))
        t = orc.huff_build(freq)
        assert t.codes == h.codes, name
        assert refport.huffman_codes(freq) == h.codes, name
        out[name] = dict(freq=freq, codes={str(k): v for k, v in h.codes.items()}, symbols=syms, bytes=enc(h, syms))
        assert orc.huff_encode(t, syms).hex() == out[name]["bytes"], name

    g, cnt = kat5_counts()
    h5 = HuffmanCoding(freq_dict(cnt.tolist()))
    idx = torch.randint(0, 1024, (4096,), generator=g)
    p = os.path.join(tmp, "k5.bin")
    h5.compress(idx, p)
    data = open(p, "rb").read()
    assert h5.decompress_string(p) == idx.tolist()
    lens = [len(h5.codes[i]) for i in range(1024)]
    out["kat5"] = dict(first_counts=cnt[:5].tolist(), first_idx=idx[:5].tolist(), code_digest=digest(h5.codes, 1024),
                       min_len=min(lens), max_len=max(lens), stream_len=len(data),
                       stream_sha256=hashlib.sha256(data).hexdigest())
    t5 = orc.huff_build(cnt.long().numpy())
    assert digest(t5.codes, 1024) == out["kat5"]["code_digest"]
    assert orc.huff_encode(t5, idx.numpy()) == data
    assert orc.huff_decode(t5, data) == idx.tolist()

    h6 = HuffmanCoding(freq_dict([0] * 1024))
    lens = [len(h6.codes[i]) for i in range(1024)]
    h6.compress(idx, p)
    data6 = open(p, "rb").read()
    out["kat6"] = dict(code_digest=digest(h6.codes, 1024), min_len=min(lens), max_len=max(lens),
                       stream_len=len(data6), stream_sha256=hashlib.sha256(data6).hexdigest())
    t6 = orc.huff_build([0] * 1024)
    assert digest(t6.codes, 1024) == out["kat6"]["code_digest"]
    assert orc.huff_encode(t6, idx.numpy()) == data6

    b = BinaryCoding()
    for name, bits in (("kat4", [1, 0, 1, 1, 0, 0, 0, 1, 1]), ("kat4b", [1, 0, 1, 1, 0, 0, 0, 1]), ("kat4c", [])):
        b.compress(torch.tensor(bits, dtype=torch.int32), p)
        data = open(p, "rb").read()
        back = b.decompress_string(p)
        assert (back is None and not bits) or back == bits
        out[name] = dict(bits=bits, bytes=data.hex())
        assert orc.bits_encode(bits) == data
    return out


# ------------------------------------------------------------------------------------------
def vq_cases():
    torch.manual_seed(0)
    vq = VectorQuantize2(1024, 4, 0.25).eval()
    E = vq.embedding.weight.detach().clone()
    out = {"codebook": E.numpy()}
    g = torch.Generator().manual_seed(7)
    near = E[torch.randint(0, 1024, (2 * 48 * 32,), generator=g)] + 1e-4 * torch.randn(2 * 48 * 32, 4, generator=g)
    cases = {
        "randn": torch.randn(2, 4, 64, 64, generator=g),          # adversarial: exact fp32 ties at the minimum
        "small": torch.randn(1, 4, 32, 48, generator=g) * 1e-3,
        "near": near.view(2, 48, 32, 4).permute(0, 3, 1, 2).contiguous(),
        "zeros": torch.zeros(1, 4, 8, 8),
        "big": torch.randn(1, 4, 8, 8, generator=g) * 1e3,
    }
    # blocky latent: constant over 4x4 / 2x2 blocks like the encoder's mask-mix output
    c = torch.randn(1, 4, 4, 4, generator=g) * 1e-3
    m = torch.randn(1, 4, 8, 8, generator=g) * 1e-3
    f = torch.randn(1, 4, 16, 16, generator=g) * 1e-3
    up = lambda t, r: t.repeat_interleave(r, -1).repeat_interleave(r, -2)
    sel = torch.rand(1, 1, 4, 4, generator=g) < 0.4
    sel2 = (torch.rand(1, 1, 8, 8, generator=g) < 0.5) & ~up(sel, 2)
    cases["blocky"] = torch.where(up(sel, 4), up(c, 4), torch.where(up(sel2, 2), up(m, 2), f))
    for name, z in cases.items():
        with torch.no_grad():
            zq, loss, idx = vq(z)
            flat = z.permute(0, 2, 3, 1).reshape(-1, 4)
            d = flat.pow(2).sum(1, keepdim=True) + E.pow(2).sum(1) - 2 * flat @ E.t()
        ties = int(((d == d.min(1, keepdim=True).values).sum(1) > 1).sum())
        ozq, oloss, oidx = orc.vq_assign(z.numpy(), E.numpy())
        assert np.array_equal(oidx, idx.numpy()), name
        assert np.array_equal(ozq.view(np.uint32), zq.numpy().view(np.uint32)), name
        assert abs(float(oloss) - float(loss)) <= 1e-5 * abs(float(loss)) + 1e-12, (name, oloss, loss)
        pzq, ploss, pidx = refport.vq_forward(z, E)
        assert torch.equal(pidx, idx) and torch.equal(pzq, zq)
        out[f"{name}_z"] = z.numpy()
        out[f"{name}_idx"] = idx.numpy().astype(np.int16)
        out[f"{name}_zq_sha"] = np.frombuffer(hashlib.sha256(zq.numpy().tobytes()).digest(), np.uint8)
        out[f"{name}_loss"] = np.float32(loss)
        out[f"{name}_ties"] = np.int32(ties)
        print(f"  vq {name}: N={idx.numel()} ties={ties} loss={float(loss):.6g}")
    return out


# ------------------------------------------------------------------------------------------
def router_cases():
    out = {}
    torch.manual_seed(0)
    e16 = torch.rand(2, 16, 16)
    e8 = torch.rand(2, 32, 32)
    out["e16"], out["e8"] = e16.numpy(), e8.numpy()
    ratios = [(0.1, 0.8), (0.3, 0.6), (0.05, 0.05), (0.0, 0.5), (0.5, 0.0), (0.2, 0.8), (0.1, 0.9), (0.5, 0.5),
              (0.3, 0.7), (1.0, 0.0), (0.0, 1.0), (0.0, 0.0), (0.001, 0.001), (0.1, 0.4)]
    modes = []
    for i, (c, m) in enumerate(ratios):
        r = TripleGrainFixedEntropyRouter(c, m)
        for tag, a, b in (("b1", e16[:1], e8[:1]), ("b2", e16, e8)):
            mask, gate, rr, mode = r(a, b)
            omc, omm, omf, omode = orc.router(a.numpy(), b.numpy(), c, m)
            assert omode == mode
            assert np.array_equal(omc, mask[0].numpy()) and np.array_equal(omm, mask[1].numpy()) \
                and np.array_equal(omf, mask[2].numpy()), (c, m, tag)
            for lvl in range(3):
                out[f"r{i}_{tag}_m{lvl}"] = np.packbits(mask[lvl].numpy().astype(np.uint8).ravel())
            out[f"r{i}_{tag}_gate_sha"] = np.frombuffer(hashlib.sha256(gate.numpy().tobytes()).digest(), np.uint8)
        modes.append(mode)
    # ties: a constant map selects nothing (quirk Q4)
    r = TripleGrainFixedEntropyRouter(0.1, 0.8)
    mask, _, _, _ = r(torch.full((1, 4, 4), 0.5), torch.full((1, 8, 8), 0.5))
    assert int(mask[0].sum()) == 0 and int(mask[1].sum()) == 0
    omc, omm, omf, _ = orc.router(np.full((1, 4, 4), 0.5, np.float32), np.full((1, 8, 8), 0.5, np.float32), 0.1, 0.8)
    assert omc.sum() == 0 and omm.sum() == 0 and omf.all()
    out["ratios"] = np.asarray(ratios, np.float64)
    out["modes"] = np.asarray(modes, np.int32)
    return out


# ------------------------------------------------------------------------------------------
def build_model(c_ratio, m_ratio, counts):
    cfg = yaml.safe_load(open(os.path.join(REF, "configs/config_inference.yaml")))["model"]["params"]
    cfg.update(ckpt_path=None, lossconfig=None, ema_decay=None)
    cfg["ddconfig"]["ch"] = 32
    cfg["ddconfig"]["router_config"]["params"] = dict(coarse_grain_ratio=c_ratio, medium_grain_ratio=m_ratio)
    torch.manual_seed(0)
    model = CGIC(**cfg).eval()
    # an untrained codebook U(+-1/1024) against O(1) latents degenerates to a handful of codes;
    # widen it so the indices exercise the whole table (the codebook is an input of the path).
    g = torch.Generator().manual_seed(99)
    model.quantize.embedding.weight.data = torch.randn(1024, 4, generator=g) * 0.5
    for i, cnt in enumerate(counts):
        model.quantize.embedding_counter[str(i)].data.fill_(float(cnt))
    return model


def e2e_case(tag, H, W, c_ratio, m_ratio, counts, image="rand"):
    model = build_model(c_ratio, m_ratio, counts)
    g = torch.Generator().manual_seed(zlib.crc32(tag.encode()))
    x = torch.rand(1, 3, H, W, generator=g)
    if image == "flat":      # large constant areas -> entropy ties (quirk Q4)
        x[:, :, : H // 2] = 0.25
    cap = {}
    hooks = [
        model.entropy_calculation_p8.register_forward_hook(lambda m, i, o: cap.__setitem__("e8", o.clone())),
        model.entropy_calculation_p16.register_forward_hook(lambda m, i, o: cap.__setitem__("e16", o.clone())),
        model.encoder.conv_out_coarse.register_forward_hook(lambda m, i, o: cap.__setitem__("hc", o.clone())),
        model.encoder.conv_out.register_forward_hook(lambda m, i, o: cap.__setitem__("hm", o.clone())),
        model.encoder.conv_out_fine.register_forward_hook(lambda m, i, o: cap.__setitem__("hf", o.clone())),
        model.encoder.register_forward_hook(lambda m, i, o: cap.update(h=o["h"].clone(), mask=[t.clone() for t in o["mask"]],
                                                                       mode=o["compression_mode"], gidx=o["indices"].clone())),
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
    files = {n: (open(os.path.join(d, n + ".bin"), "rb").read() if os.path.exists(os.path.join(d, n + ".bin")) else b"")
             for n in STREAMS}
    mode = cap["mode"]
    ind_dec = cap["emb_calls"][-1].view(1, H // 4, W // 4)
    E = model.quantize.embedding.weight.detach()

    # ---- cross-check the oracle and the port against the reference on this case
    order = [int(k) for k in model.quantize.embedding_counter.keys()]
    assert order == orc.lexicographic_order(1024).tolist()      # ParameterDict: sorted key strings
    t = orc.huff_build(np.asarray(counts, np.int64), order)
    assert t.codes == hs.codes
    ozq, oloss, oidx = orc.vq_assign(cap["z"].numpy(), E.numpy())
    assert np.array_equal(oidx, cap["ind"].numpy()), tag
    assert np.array_equal(ozq.view(np.uint32), cap["zq"].numpy().view(np.uint32)), tag
    omix = orc.mask_mix(cap["hc"].numpy(), cap["hm"].numpy(), cap["hf"].numpy(), cap["mask"][0].numpy(),
                        cap["mask"][1].numpy(), cap["mask"][2].numpy())
    assert np.array_equal(omix.view(np.uint32), cap["h"].numpy().view(np.uint32)), tag
    omc, omm, omf, omode = orc.router(cap["e16"].numpy(), cap["e8"].numpy(), c_ratio, m_ratio)
    assert omode == mode and np.array_equal(omc, cap["mask"][0].numpy()) and np.array_equal(omm, cap["mask"][1].numpy()) \
        and np.array_equal(omf, cap["mask"][2].numpy()), tag
    streams = orc.pack_image(t, cap["ind"].view(H // 4, W // 4).numpy(), omc[0, 0], omm[0, 0], omf[0, 0], mode)
    for s, n in enumerate(STREAMS):
        assert streams[s] == files[n], (tag, n, len(streams[s]), len(files[n]))
    assert orc.bpp_of(streams, H, W) == bpp
    umc, umm, umf, uind, uq = orc.unpack_image(t, streams, H // 4, W // 4, mode, E.numpy())
    assert np.array_equal(uind, ind_dec[0].numpy()), tag
    assert np.array_equal(uq, cap["quant_dec"][0].numpy()), tag
    for o, r in zip((umc, umm, umf), cap["mask_dec"]):
        assert np.array_equal(o, r[0, 0].numpy().astype(np.int64)), tag
    for p in (8, 16):
        oe = orc.entropy(x.numpy(), p)
        ref_e = cap["e8" if p == 8 else "e16"].numpy()
        assert np.allclose(oe, ref_e, rtol=2e-5, atol=1e-6), (tag, p, np.abs(oe - ref_e).max())
    if mode == 0:
        table = refport.huffman_codes(counts, order)
        assert table == hs.codes
        rev = {v: k for k, v in table.items()}
        pind, pbpp, pind_dec, pq, _ = refport.roundtrip_mode0(cap["z"], E, cap["mask"], table, rev, tempfile.mkdtemp())
        assert torch.equal(pind.flatten(), cap["ind"]) and pbpp == bpp and torch.equal(pind_dec, ind_dec) \
            and torch.equal(pq, cap["quant_dec"])

    out = dict(x=x.numpy(), e8=cap["e8"].numpy(), e16=cap["e16"].numpy(), hc=cap["hc"].numpy(), hm=cap["hm"].numpy(),
               hf=cap["hf"].numpy(), h=cap["h"].numpy(), z=cap["z"].numpy(), codebook=E.numpy(),
               counts=np.asarray(counts, np.int64), order=np.asarray(order, np.int32),
               code_digest=np.asarray(digest(hs.codes, 1024)), ratios=np.asarray([c_ratio, m_ratio], np.float64), mode=np.int32(mode),
               ind=cap["ind"].numpy().astype(np.int16), zq=cap["zq"].numpy(), loss=np.float32(cap["loss"]),
               ind_dec=ind_dec.numpy().astype(np.int16), quant_dec=cap["quant_dec"].numpy(), bpp=np.float64(bpp),
               gidx_shape=np.asarray(cap["gidx"].shape, np.int32))
    for lvl in range(3):
        out[f"mask{lvl}"] = cap["mask"][lvl].numpy().astype(np.uint8)
        out[f"mask_dec{lvl}"] = cap["mask_dec"][lvl].numpy().astype(np.uint8)
        out[f"mask_dec{lvl}_dtype"] = np.asarray(str(cap["mask_dec"][lvl].dtype))
    out["ind_dec_dtype"] = np.asarray(str(ind_dec.dtype))
    for n in STREAMS:
        out["file_" + n] = np.frombuffer(files[n], np.uint8)
    sizes = [len(files[n]) for n in STREAMS]
    print(f"  e2e {tag}: {H}x{W} ratio=({c_ratio},{m_ratio}) mode={mode} sizes={sizes} bpp={bpp:.5f} "
          f"n=({int(cap['mask'][0].sum())},{int(cap['mask'][1].sum())},{int(cap['mask'][2].sum())})")
    return out


def main():
    torch.set_num_threads(1)
    print("huffman KATs")
    with open(os.path.join(HERE, "huffman_kats.json"), "w") as f:
        json.dump(huffman_kats(), f, indent=1, sort_keys=True)
    print("vq cases")
    np.savez_compressed(os.path.join(HERE, "vq_cases.npz"), **vq_cases())
    print("router cases")
    np.savez_compressed(os.path.join(HERE, "router_cases.npz"), **router_cases())
    print("end-to-end cases (unmodified CGIC.compress, ch=32)")
    _, cnt5 = kat5_counts()
    cnt5 = cnt5.long().tolist()
    zero = [0] * 1024
    cases = [
        ("m0_a", 64, 64, 0.1, 0.8, cnt5, "rand"),
        ("m0_b", 96, 128, 0.3, 0.6, cnt5, "rand"),
        ("m0_c", 64, 96, 0.05, 0.05, cnt5, "rand"),
        ("m0_long", 64, 64, 0.1, 0.8, zero, "rand"),       # untrained counters: codes up to 224 bits
        ("m0_flat", 64, 64, 0.1, 0.8, cnt5, "flat"),       # entropy ties
        ("m0_cfg", 128, 128, 0.1, 0.4, cnt5, "rand"),      # the shipped config_inference.yaml ratios
        ("m1", 64, 64, 0.0, 0.5, cnt5, "rand"),
        ("m2", 64, 64, 0.5, 0.0, cnt5, "rand"),
        ("m3", 64, 64, 0.2, 0.8, cnt5, "rand"),
        ("m4", 64, 64, 1.0, 0.0, cnt5, "rand"),
        ("m5", 64, 64, 0.0, 1.0, cnt5, "rand"),
        ("m6", 64, 64, 0.0, 0.0, cnt5, "rand"),
        ("m6_long", 32, 48, 0.0, 0.0, zero, "rand"),
    ]
    for tag, H, W, c, m, counts, image in cases:
        np.savez_compressed(os.path.join(HERE, f"e2e_{tag}.npz"), **e2e_case(tag, H, W, c, m, counts, image))
    print("done")


if __name__ == "__main__":
    main()
