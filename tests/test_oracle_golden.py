"""The oracle (oracle/cgic_oracle.c via oracle/oracle.py, and the reference-shaped port
oracle/refport.py) against the fixtures generated from the unmodified reference
(tests/golden/make_golden.py) and the KATs of SURVEY.md 8(c).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import STREAMS, big_case_names, e2e_case_names, load_npz, unpack_mask_bits
from oracle import oracle as orc
from oracle import refport


def digest(codes, K):
    return hashlib.sha256("".join(f"{i}:{codes[i]};" for i in range(K)).encode()).hexdigest()


def kat5():
    g = torch.Generator().manual_seed(1234)
    cnt = (-torch.log(torch.rand(1024, generator=g)) * 1000).floor().long().numpy()
    idx = torch.randint(0, 1024, (4096,), generator=g).numpy()
    return cnt, idx


@pytest.mark.parametrize("name", ["kat1", "kat1b", "kat1c", "kat2", "kat3", "kat_desc", "kat_pow2"])
def test_huffman_small_kats(kats, name):
    k = kats[name]
    t = orc.huff_build(k["freq"])
    assert {str(s): c for s, c in t.codes.items()} == k["codes"]
    assert refport.huffman_codes(k["freq"]) == {int(s): c for s, c in k["codes"].items()}
    data = orc.huff_encode(t, k["symbols"])
    assert data.hex() == k["bytes"]
    back = orc.huff_decode(t, data)
    assert back == (k["symbols"] if k["symbols"] else None)


def test_survey_kat1_literal():
    # SURVEY.md 8(c) KAT1, typed in by hand (independent of the json fixture)
    t = orc.huff_build([5, 9, 12, 13, 16, 45, 0, 0])
    assert t.codes == {5: "0", 2: "100", 3: "101", 6: "110000", 7: "110001", 0: "11001", 1: "1101", 4: "111"}
    assert orc.huff_encode(t, [5, 0, 1, 6, 7, 5, 5, 2, 3, 4]).hex() == "076770c49780"
    assert orc.huff_encode(t, [5] * 8).hex() == "080000"
    assert orc.huff_encode(t, []) == b"" and orc.huff_decode(t, b"") is None
    t2 = orc.huff_build([1] * 8)
    assert t2.codes == {0: "000", 2: "001", 7: "010", 1: "011", 6: "100", 5: "101", 4: "110", 3: "111"}
    assert orc.bits_encode([1, 0, 1, 1, 0, 0, 0, 1, 1]).hex() == "07b180"
    assert orc.bits_encode([1, 0, 1, 1, 0, 0, 0, 1]).hex() == "08b100"
    assert orc.bits_decode(bytes.fromhex("07b180")) == [1, 0, 1, 1, 0, 0, 0, 1, 1]


def test_huffman_kat5_kat6(kats):
    cnt, idx = kat5()
    assert cnt[:5].tolist() == [3541, 911, 1347, 1003, 2842] and idx[:5].tolist() == [737, 314, 126, 419, 666]
    t5 = orc.huff_build(cnt)
    assert digest(t5.codes, 1024) == kats["kat5"]["code_digest"] == \
        "c88daeb9f3412d94697a0a7ec56d0ba0c523025003f5a41e01b17bd6be6a21c6"
    data = orc.huff_encode(t5, idx)
    assert len(data) == kats["kat5"]["stream_len"] == 5518
    assert hashlib.sha256(data).hexdigest() == kats["kat5"]["stream_sha256"] == \
        "98807e2c7bdb26e0997b9fd65b3d0582c3c4118ddfe5b5f64d00ee0187efbe91"
    assert orc.huff_decode(t5, data) == idx.tolist()
    t6 = orc.huff_build([0] * 1024)
    assert digest(t6.codes, 1024) == kats["kat6"]["code_digest"] == \
        "e867e46a7333cfe2bff157e72d26fc7f4de79ce9a7ee0890a89e0b5d603e58ec"
    assert t6.max_len == kats["kat6"]["max_len"] == 224
    data6 = orc.huff_encode(t6, idx)
    assert hashlib.sha256(data6).hexdigest() == kats["kat6"]["stream_sha256"]
    assert orc.huff_decode(t6, data6) == idx.tolist()


def test_port_matches_c_on_kat5():
    cnt, idx = kat5()
    order = orc.lexicographic_order(1024)
    assert refport.huffman_codes(cnt.tolist(), order.tolist()) == orc.huff_build(cnt, order).codes


def test_vq_cases():
    g = load_npz("vq_cases.npz")
    cb = g["codebook"]
    for name in ("randn", "small", "near", "zeros", "big", "blocky"):
        zq, loss, idx = orc.vq_assign(g[f"{name}_z"], cb)
        assert np.array_equal(idx, g[f"{name}_idx"].astype(np.int64)), name
        assert hashlib.sha256(zq.tobytes()).digest() == g[f"{name}_zq_sha"].tobytes(), name
        assert np.isclose(loss, g[f"{name}_loss"], rtol=1e-5), name
    assert int(g["randn_ties"]) > 0  # the adversarial set really contains exact fp32 ties


def test_router_cases():
    g = load_npz("router_cases.npz")
    for i, (c, m) in enumerate(g["ratios"]):
        assert orc.router_mode(float(c), float(m)) == int(g["modes"][i])
        for tag, sl in (("b1", slice(0, 1)), ("b2", slice(0, 2))):
            mc, mm, mf, mode = orc.router(g["e16"][sl], g["e8"][sl], float(c), float(m))
            for lvl, arr in enumerate((mc, mm, mf)):
                assert np.array_equal(np.packbits(arr.astype(np.uint8).ravel()), g[f"r{i}_{tag}_m{lvl}"]), (c, m, tag, lvl)
    # KAT7 / KAT8 of SURVEY.md 8(c)
    mc, mm, mf, mode = orc.router(g["e16"][:1], g["e8"][:1], 0.1, 0.8)
    assert (int(mc.sum()), int(mm.sum()), int(mf.sum())) == (25, 821, 412)
    mc, mm, mf, mode = orc.router(g["e16"], g["e8"], 0.1, 0.8)
    assert mc.reshape(2, -1).sum(1).tolist() == [27, 23] and mm.reshape(2, -1).sum(1).tolist() == [806, 836]
    for (c, m), want in {(0, .5): 1, (.5, 0): 2, (.2, .8): 3, (.1, .9): 3, (.5, .5): 3, (.3, .7): 3, (1, 0): 4,
                         (0, 1): 5, (0, 0): 6, (.1, .8): 0}.items():
        assert orc.router_mode(c, m) == want


@pytest.mark.parametrize("tag", e2e_case_names())
def test_e2e_against_reference_run(tag, tmp_path):
    g = load_npz(f"e2e_{tag}.npz")
    H, W = g["x"].shape[-2:]
    h, w = H // 4, W // 4
    mode = int(g["mode"])
    c_ratio, m_ratio = map(float, g["ratios"])
    # a4 entropy (tolerance), a5 router (exact given the reference's entropy maps), a6 mix (exact)
    assert np.allclose(orc.entropy(g["x"], 8), g["e8"], rtol=2e-5, atol=1e-6)
    assert np.allclose(orc.entropy(g["x"], 16), g["e16"], rtol=2e-5, atol=1e-6)
    mc, mm, mf, omode = orc.router(g["e16"], g["e8"], c_ratio, m_ratio)
    assert omode == mode
    for lvl, arr in enumerate((mc, mm, mf)):
        assert np.array_equal(arr.astype(np.uint8), g[f"mask{lvl}"])
    mix = orc.mask_mix(g["hc"], g["hm"], g["hf"], mc, mm, mf)
    assert np.array_equal(mix.view(np.uint32), g["h"].view(np.uint32))
    # a1 VQ
    zq, loss, idx = orc.vq_assign(g["z"], g["codebook"])
    assert np.array_equal(idx, g["ind"].astype(np.int64))
    assert np.array_equal(zq.view(np.uint32), g["zq"].view(np.uint32))
    assert np.isclose(loss, g["loss"], rtol=1e-5)
    # a8 table, a7/a9/a11/a12 pack
    t = orc.huff_build(g["counts"], g["order"])
    assert digest(t.codes, 1024) == str(g["code_digest"])
    streams = orc.pack_image(t, idx.reshape(h, w), mc[0, 0], mm[0, 0], mf[0, 0], mode)
    for s, n in enumerate(STREAMS):
        assert streams[s] == g["file_" + n].tobytes(), n
    assert orc.bpp_of(streams, H, W) == float(g["bpp"])
    # a10/a13/a14 unpack
    umc, umm, umf, uind, uq = orc.unpack_image(t, streams, h, w, mode, g["codebook"])
    assert np.array_equal(uind, g["ind_dec"][0].astype(np.int64))
    assert np.array_equal(uq, g["quant_dec"][0])
    for lvl, arr in enumerate((umc, umm, umf)):
        assert np.array_equal(arr.astype(np.uint8), g[f"mask_dec{lvl}"][0, 0])
    # the reference-shaped port agrees as well (mode 0 is what bench.py times)
    if mode == 0:
        table = refport.huffman_codes(g["counts"].tolist(), g["order"].tolist())
        rev = {v: k for k, v in table.items()}
        masks = [torch.from_numpy(g[f"mask{lvl}"].astype(np.int32)) for lvl in range(3)]
        pind, pbpp, pind_dec, pq, sizes = refport.roundtrip_mode0(torch.from_numpy(g["z"]), torch.from_numpy(g["codebook"]),
                                                                   masks, table, rev, str(tmp_path))
        assert pbpp == float(g["bpp"]) and sizes == [len(s) for s in streams]
        assert np.array_equal(pind_dec.numpy()[0], uind) and np.array_equal(pq.numpy()[0], uq)


@pytest.mark.parametrize("tag", big_case_names())
def test_big_cases_against_reference_run(tag):
    """BASELINE-size fixtures (tests/golden/make_golden_big.py: 256x256, 512x768 at the three config-3 ratios, the
    768x496 / 576x496 tiles of config 5): router, VQ, pack, unpack of the oracle == the unmodified reference's run."""
    g = load_npz(f"big_{tag}.npz")
    H, W = map(int, g["shape"])
    h, w = H // 4, W // 4
    mode = int(g["mode"])
    c_ratio, m_ratio = map(float, g["ratios"])
    mc, mm, mf, omode = orc.router(g["e16"], g["e8"], c_ratio, m_ratio)
    assert omode == mode
    for lvl, arr in enumerate((mc, mm, mf)):
        assert np.array_equal(np.packbits(arr.astype(np.uint8).ravel()), g[f"mask{lvl}_bits"]), lvl
    if "x" in g:   # the 256x256 case also pins Entropy and the mask-mix
        assert np.allclose(orc.entropy(g["x"], 8), g["e8"], rtol=2e-5, atol=1e-6)
        assert np.allclose(orc.entropy(g["x"], 16), g["e16"], rtol=2e-5, atol=1e-6)
        mix = orc.mask_mix(g["hc"], g["hm"], g["hf"], mc, mm, mf)
        assert hashlib.sha256(mix.tobytes()).digest() == g["h_sha"].tobytes()
    zq, loss, idx = orc.vq_assign(g["z"], g["codebook"])
    assert np.array_equal(idx, g["ind"].astype(np.int64))
    assert hashlib.sha256(zq.tobytes()).digest() == g["zq_sha"].tobytes()
    assert np.isclose(loss, g["loss"], rtol=1e-5)
    t = orc.huff_build(g["counts"], g["order"])
    streams = orc.pack_image(t, idx.reshape(h, w), mc[0, 0], mm[0, 0], mf[0, 0], mode)
    for s, n in enumerate(STREAMS):
        assert streams[s] == g["file_" + n].tobytes(), n
    assert orc.bpp_of(streams, H, W) == float(g["bpp"])
    umc, umm, umf, uind, uq = orc.unpack_image(t, streams, h, w, mode, g["codebook"])
    assert np.array_equal(uind, g["ind_dec"][0].astype(np.int64))
    assert hashlib.sha256(np.ascontiguousarray(uq[None]).tobytes()).digest() == g["quant_dec_sha"].tobytes()
    for lvl, arr in enumerate((umc, umm, umf)):
        assert np.array_equal(np.packbits(arr.astype(np.uint8).ravel()), g[f"mask_dec{lvl}_bits"]), lvl


def test_framing_properties():
    """pad in 1..8, size = nbits//8 + 2, round trip for every length 0..40 (quirk Q1/Q5)."""
    rng = np.random.default_rng(0)
    for n in range(0, 41):
        bits = rng.integers(0, 2, n).tolist()
        data = orc.bits_encode(bits)
        if n == 0:
            assert data == b""
            continue
        assert len(data) == n // 8 + 2 and 1 <= data[0] <= 8 and data[0] == 8 - n % 8
        assert orc.bits_decode(data) == bits


# ---- f4 SpatialNorm: the numpy restatement against vectors from the reference class (make_spatial_norm_golden.py) ----
def _sn_cases():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "spatial_norm.npz"))
    for name, add_conv in zip(d["cases"], d["add_conv"]):
        sd = {k[len(name) + 4:]: d[k] for k in d.files if k.startswith(f"{name}.sd.")}
        yield str(name), bool(add_conv), d[f"{name}.f"], d[f"{name}.zq"], sd, d[f"{name}.out"]


def test_spatial_norm_oracle_matches_reference_vectors():
    import torch
    from oracle import oracle as orc
    seen = 0
    for name, add_conv, f, zq, sd, want in _sn_cases():
        if add_conv:   # decoder.py:49-51: the 3x3 conv acts on the up-sampled zq; the restatement then sees zq at full size
            iy, ix = orc.nearest_index(f.shape[2], zq.shape[2]), orc.nearest_index(f.shape[3], zq.shape[3])
            zu = torch.from_numpy(zq[:, :, iy][:, :, :, ix])
            zq = torch.nn.functional.conv2d(zu, torch.from_numpy(sd["conv.weight"]), torch.from_numpy(sd["conv.bias"]), padding=1).numpy()
        got = orc.spatial_norm(f, zq, sd["norm_layer.weight"], sd["norm_layer.bias"], sd["conv_y.weight"], sd["conv_y.bias"],
                               sd["conv_b.weight"], sd["conv_b.bias"], groups=32, eps=1e-6)
        # tolerance: the reference computes in fp32 (statistics, 1x1 convs); the restatement in float64
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5, err_msg=name)
        seen += 1
    assert seen == 5
