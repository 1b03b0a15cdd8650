"""Randomised differential test: the whole hot path (router -> mask-mix -> VQ (indexed and exhaustive) -> pack ->
unpack) on the GPU against the oracle, over random image shapes, batch sizes, ratios (all 7 modes), code tables
(including zero counts and very skewed ones), codebooks and latent distributions.  Seeds are fixed: every run
sees the same 40 cases."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RATIOS = [(0.1, 0.8), (0.3, 0.6), (0.05, 0.05), (0.5, 0.25), (0.0, 0.5), (0.5, 0.0), (0.2, 0.8), (1.0, 0.0), (0.0, 1.0), (0.0, 0.0)]


@pytest.fixture(scope="module")
def cg():
    import cgic_b200
    assert torch.cuda.is_available()
    return cgic_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _case(seed):
    g = torch.Generator().manual_seed(90_000 + seed)
    r = lambda n: int(torch.randint(0, n, (1,), generator=g))
    B = 1 + r(3)
    H, W = 16 * (1 + r(12)), 16 * (1 + r(12))
    c, m = RATIOS[seed % len(RATIOS)]
    K = [1024, 1024, 256, 37, 1000][r(5)]
    kind = r(4)
    if kind == 0:
        counts = (-torch.log(torch.rand(K, generator=g)) * 1000).floor().long()
    elif kind == 1:
        counts = torch.zeros(K, dtype=torch.long)                                  # untrained model: very long codes
    elif kind == 2:
        counts = (torch.rand(K, generator=g) ** 8 * 1e6).floor().long()            # very skewed
    else:
        counts = torch.randint(0, 3, (K,), generator=g)                             # many ties
    cb = [(torch.rand(K, 4, generator=g) * 2 - 1) / K, torch.randn(K, 4, generator=g)][r(2)]
    e16 = torch.rand(B, H // 16, W // 16, generator=g)
    e8 = torch.rand(B, H // 8, W // 8, generator=g)
    if r(4) == 0:
        e16[:] = 0.5                                                                # ties in the router
    scale = cb.abs().max()
    heads = []
    for div in (16, 8, 4):
        n = B * (H // div) * (W // div)
        if r(2):
            v = cb[torch.randint(0, K, (n,), generator=g)] + 1e-3 * scale * torch.randn(n, 4, generator=g)
        else:
            v = torch.randn(n, 4, generator=g) * scale
        heads.append(v.view(B, H // div, W // div, 4).permute(0, 3, 1, 2).contiguous())
    return B, H, W, c, m, K, counts, cb.contiguous(), e16, e8, heads


@pytest.mark.parametrize("seed", range(40))
def test_random_case_against_oracle(cg, orc, seed):
    B, H, W, c, m, K, counts, cb, e16, e8, heads = _case(seed)
    h, w = H // 4, W // 4
    dev = "cuda"
    mc, mm, mf, gate, mode = cg.ops.router(e16.to(dev), e8.to(dev), c, m, per_image=True, want_gate=True)
    z = cg.ops.mask_mix(*(t.to(dev) for t in heads), mc, mm, mf)
    cbd = cb.to(dev)
    idx, zq, sq = cg.ops.vq_assign(z, cg.ops.Codebook(cbd))
    idx_x, zq_x, _ = cg.ops.vq_assign(z, cbd)
    assert torch.equal(idx, idx_x) and torch.equal(zq.view(torch.int32), zq_x.view(torch.int32))
    order = sorted(range(K), key=str)
    table = cg.ops.HuffTable(counts.numpy(), order)
    packed, sizes = cg.ops.pack(idx, mc, mm, mf, mode, table, h, w)
    dmc, dmm, dmf, ind, quant, status = cg.ops.unpack(packed, sizes, mode, table, cbd, h, w)
    assert int(status.abs().sum()) == 0
    ot = orc.huff_build(counts.numpy(), order)
    assert table.max_len == ot.max_len
    offs, _, _ = table.layout(h, w)
    for b in range(B):
        omc, omm, omf, omode = orc.router(e16[b:b + 1].numpy(), e8[b:b + 1].numpy(), c, m)
        assert omode == mode
        for got, want in zip((mc, mm, mf), (omc, omm, omf)):
            assert np.array_equal(got[b].cpu().numpy(), want[0]), (seed, "mask")
        oz = orc.mask_mix(*(t[b:b + 1].numpy() for t in heads), omc, omm, omf)
        assert np.array_equal(z[b:b + 1].cpu().numpy().view(np.uint32), oz.view(np.uint32)), (seed, "mix")
        ozq, _, oidx = orc.vq_assign(oz, cb.numpy())
        assert np.array_equal(idx.view(B, -1)[b].cpu().numpy(), oidx), (seed, "vq")
        assert np.array_equal(zq[b:b + 1].cpu().numpy().view(np.uint32), ozq.view(np.uint32))
        streams = orc.pack_image(ot, oidx.reshape(h, w), omc[0, 0], omm[0, 0], omf[0, 0], mode)
        blob, sz = packed[b].cpu().numpy(), sizes[b].cpu().numpy()
        for s in range(5):
            assert blob[offs[s]: offs[s] + sz[s]].tobytes() == streams[s], (seed, "stream", s)
        umc, umm, umf, uind, uq = orc.unpack_image(ot, streams, h, w, mode, cb.numpy())
        assert np.array_equal(ind[b].cpu().numpy(), uind) and np.array_equal(quant[b].cpu().numpy(), uq), (seed, "decode")
        for got, want in zip((dmc, dmm, dmf), (umc, umm, umf)):
            assert np.array_equal(got[b].cpu().numpy(), want)
