"""CPU-side checks of the product: the C-ABI library loads and exports every symbol that
include/cgic_b200.h declares, the host-side Huffman table builder (our code, not the oracle)
reproduces the reference's tables, the host logic (modes, ranks, layouts, sharding) is right,
and nothing in the product imports the oracle.  No compute kernels run here."""
import ctypes
import glob
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, e2e_case_names, load_npz
from oracle import oracle as orc

import cgic_b200
from cgic_b200 import _lib, ops


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cgic_b200.h")).read()
    return re.findall(r"CGIC_API[^;(]*?\b(cgic_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 30 and len(set(names)) == len(names)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().cgic_abi_version() == 1


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "control-gic_b200")
    files = glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True) + glob.glob(os.path.join(pkg, "csrc", "*"))
    assert files
    for f in files:
        if os.path.isfile(f) and not f.endswith(".so"):
            src = open(f, errors="ignore").read()
            assert not re.search(r"^\s*(from|import)\s+oracle|orc_\w+\(|libcgic_oracle", src, re.M), f


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.vq_assign(torch.zeros(1, 4, 4, 4), torch.zeros(8, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.entropy_maps(torch.zeros(1, 3, 16, 16))


def digest(codes, K):
    return hashlib.sha256("".join(f"{i}:{codes[i]};" for i in range(K)).encode()).hexdigest()


@pytest.mark.parametrize("name", ["kat1", "kat2", "kat3", "kat_desc", "kat_pow2"])
def test_host_table_builder_kats(kats, name):
    t = ops.HuffTable(kats[name]["freq"])
    assert {str(s): c for s, c in t.codes().items()} == kats[name]["codes"]
    assert t.max_len == max(len(c) for c in kats[name]["codes"].values())


def test_host_table_builder_large(kats):
    g = torch.Generator().manual_seed(1234)
    cnt = (-torch.log(torch.rand(1024, generator=g)) * 1000).floor().long().numpy()
    assert digest(ops.HuffTable(cnt).codes(), 1024) == kats["kat5"]["code_digest"]
    t6 = ops.HuffTable([0] * 1024)
    assert digest(t6.codes(), 1024) == kats["kat6"]["code_digest"] and t6.max_len == 224
    # the model's ParameterDict order (lexicographic keys), pinned by the reference runs
    for tag in ("m0_a", "m0_long"):
        e = load_npz(f"e2e_{tag}.npz")
        t = ops.HuffTable(e["counts"], e["order"])
        assert digest(t.codes(), 1024) == str(e["code_digest"])
        assert np.array_equal(t.lengths(), orc.huff_build(e["counts"], e["order"]).lengths)


def test_codec_object_host_side():
    vq = cgic_b200.VectorQuantize2(1024, 4, 0.25)
    assert list(vq.embedding_counter.keys())[:5] == ["0", "1", "10", "100", "1000"]
    assert vq.counter_order() == orc.lexicographic_order(1024).tolist()
    e = load_npz("e2e_m0_a.npz")
    for i in range(1024):
        vq.embedding_counter[str(i)].data.fill_(float(e["counts"][i]))
    assert torch.equal(vq.counters_flat(), torch.from_numpy(e["counts"]).float())
    h = cgic_b200.HuffmanCoding(vq.embedding_counter)
    assert digest(h.codes, 1024) == str(e["code_digest"])
    assert all(h.reverse_mapping[c] == s for s, c in h.codes.items())
    b = cgic_b200.BinaryCoding()
    assert b.codes == {0: "0", 1: "1"} and b.reverse_mapping == {"0": 0, "1": 1}
    sd = vq.state_dict()
    assert sd["embedding.weight"].shape == (1024, 4) and sd["embedding_counter.17"].shape == (1,)


def test_error_reporting():
    handle = ctypes.c_void_p()
    f = np.zeros(1, np.int64)
    rc = _lib.lib().cgic_huff_build(f.ctypes.data, None, 1, ctypes.byref(handle))
    assert rc == _lib.EINVAL and b"K=1" in _lib.lib().cgic_last_error()
    with pytest.raises(_lib.CgicError):
        ops.HuffTable([1, 2, 3], [0, 0, 1])                  # order is not a permutation
    with pytest.raises(_lib.CgicError):
        ops.HuffTable([1] * 8).layout(6, 8)                  # token grid must be multiples of 4


def test_layout_matches_oracle():
    for counts in ([0] * 1024, list(range(1024))):
        t, ot = ops.HuffTable(counts), orc.huff_build(counts)
        for h, w in ((16, 16), (64, 64), (128, 192), (144, 124), (192, 192)):
            off, cap, stride = t.layout(h, w)
            ooff, ocap, ostride = orc.slot_layout(ot, h, w)
            assert off.tolist() == ooff.tolist() and cap.tolist() == ocap.tolist() and stride == ostride
            assert all(o % 16 == 0 for o in off) and stride % 16 == 0


def test_router_host_logic():
    for (c, m), want in {(0, .5): 1, (.5, 0): 2, (.2, .8): 3, (.1, .9): 3, (.5, .5): 3, (.3, .7): 3, (1, 0): 4, (0, 1): 5,
                         (0, 0): 6, (.1, .8): 0, (.3, .6): 0, (.05, .05): 0}.items():
        assert ops.router_mode(c, m) == want == orc.router_mode(c, m)
    g = load_npz("router_cases.npz")
    for (c, m), mode in zip(g["ratios"], g["modes"]):
        assert ops.router_mode(float(c), float(m)) == int(mode)
        assert ops.router_ranks(float(c), float(m), 256, 1024, int(mode)) == orc.router_ranks(float(c), float(m), 256, 1024, int(mode))
    assert ops.router_ranks(0.1, 0.8, 256, 1024, 0) == (26, 922)      # round(25.6), round(102.4 + 819.2)
    assert ops.router_ranks(0.5, 0.0, 5, 20, 2)[0] == 2               # banker's rounding of 2.5
    assert np.array_equal(ops.linspace_bins(), orc.linspace_bins())
    for mode, want in enumerate(orc.STREAMS_BY_MODE):
        assert tuple(int(ops.stream_present(mode, s)) for s in range(5)) == want


def test_workload_counts_formula():
    import workload
    assert workload.expected_counts(256, 256, 0.1, 0.8) == (25, 821, 412)
    assert workload.expected_counts(512, 768, 0.3, 0.6) == (460, 3689, 2460)
    assert workload.expected_counts(768, 768, 0.05, 0.05) == (114, 465, 33180)
    e16, e8 = workload.entropy_maps(1, 256, 256, 3)
    mc, mm, mf, _ = orc.router(e16.numpy(), e8.numpy(), 0.1, 0.8)
    assert (int(mc.sum()), int(mm.sum()), int(mf.sum())) == (25, 821, 412)


def test_spatial_norm_module_host_side():
    """State-dict keys / shapes of cgic_b200.Normalize == the reference class's (read from the golden file that
    tests/golden/make_spatial_norm_golden.py wrote from CGIC/modules/vqvae/decoder.py:34-56); no CPU fallback."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "spatial_norm.npz"))
    for name, add_conv in zip(d["cases"], d["add_conv"]):
        ref = {k[len(name) + 4:]: d[k].shape for k in d.files if k.startswith(f"{name}.sd.")}
        m = cgic_b200.Normalize(d[f"{name}.f"].shape[1], d[f"{name}.zq"].shape[1], bool(add_conv))
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == ref
        assert m.norm_layer.num_groups == 32 and m.norm_layer.eps == 1e-6
    with pytest.raises(RuntimeError, match="CUDA"):
        cgic_b200.Normalize(32, 4, False)(torch.zeros(1, 32, 4, 4), torch.zeros(1, 4, 1, 1))
