"""GPU parity of the indexed VQ search (cgic_codebook_* + cgic_vq_assign_indexed, csrc/codebook.cu):
bit-identical to the oracle's exhaustive restatement of quantize.py:69-98 and to the exhaustive
kernel, for codebooks and latents chosen to stress the index -- near-ties on bisectors, latents on
cell boundaries and outside the grid, clumped / duplicated / degenerate codebooks."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_npz

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5


@pytest.fixture(scope="module")
def cg():
    import cgic_b200
    assert torch.cuda.is_available()
    return cgic_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def codebooks(name, K, g):
    if name == "init":          # the reference's init, quantize.py:26
        return (torch.rand(K, 4, generator=g) * 2 - 1) / K
    if name == "gauss":
        return torch.randn(K, 4, generator=g)
    if name == "clumped":       # a few live codes at scale 0.5, the rest still at their tiny init
        live = K // 5
        return torch.cat([torch.randn(live, 4, generator=g) * 0.5, (torch.rand(K - live, 4, generator=g) * 2 - 1) / 1024])
    if name == "dups":          # duplicated rows: the lowest index must win
        base = torch.randn(K // 4, 4, generator=g)
        return base.repeat(4, 1)[torch.randperm(K // 4 * 4, generator=g)]
    if name == "flat_dim":      # one coordinate identical for every code
        cb = torch.randn(K, 4, generator=g)
        cb[:, 2] = 0.25
        return cb
    if name == "lattice":       # many exact ties between neighbours
        v = torch.arange(K)
        return torch.stack([(v % 6).float(), (v // 6 % 6).float(), (v // 36 % 6).float(), (v // 216).float()], 1) * 0.125
    raise KeyError(name)


def latents(cb, n, g):
    """[n,4] latents exercising every path of the indexed search."""
    K = cb.shape[0]
    span = (cb.max(0).values - cb.min(0).values).clamp_min(1e-6)
    parts = []
    q = n // 8
    pick = lambda m: cb[torch.randint(0, K, (m,), generator=g)]
    parts.append(pick(q))                                                      # exactly on codes
    parts.append(pick(q) + 1e-7 * span * torch.randn(q, 4, generator=g))       # trained-like
    parts.append(pick(q) + 0.02 * span * torch.randn(q, 4, generator=g))       # between codes
    mid = 0.5 * (pick(q) + pick(q))                                            # on bisectors of random pairs
    parts.append(mid)
    a = pick(q)                                                                # on the bisector of a code and its nearest neighbour,
    nn = torch.empty_like(a)                                                   # nudged by a few ulps: real fp32 near-ties
    for s in range(0, q, 8192):
        d = torch.cdist(a[s:s + 8192].double(), cb.double())
        d[d == 0] = float("inf")
        nn[s:s + 8192] = cb[d.argmin(1)]
    m2 = 0.5 * (a + nn)
    parts.append(m2 * (1 + 2.0 ** -23 * torch.randint(-3, 4, (q, 4), generator=g)))
    lo, hi = cb.min(0).values, cb.max(0).values                                # uniform over the bounding box and a little beyond
    parts.append(lo - 0.3 * span + (1.6 * span) * torch.rand(q, 4, generator=g))
    parts.append(torch.randn(q, 4, generator=g) * span * 3)                    # far outside the grid
    rest = n - 7 * q
    edge = lo + span * torch.randint(0, 161, (rest, 4), generator=g) / 160.0   # on fine-bin edges of the grid
    parts.append(edge)
    return torch.cat(parts)[torch.randperm(n, generator=g)].contiguous()


@pytest.mark.parametrize("kind,K", [("init", 1024), ("gauss", 1024), ("clumped", 1024), ("dups", 1024), ("flat_dim", 1024),
                                    ("lattice", 1024), ("gauss", 37), ("init", 2), ("gauss", 1), ("gauss", 4096), ("init", 1000)])
def test_indexed_vs_oracle(cg, orc, kind, K):
    g = torch.Generator().manual_seed(1000 + K + len(kind))
    cb = codebooks(kind, K, g).contiguous()
    K = cb.shape[0]
    B, h, w = 2, 64, 96
    z = latents(cb, B * h * w, g).view(B, h, w, 4).permute(0, 3, 1, 2).contiguous()
    pc = cg.ops.Codebook(cb.cuda())
    idx, zq, sq = cg.ops.vq_assign(z.cuda(), pc)
    ozq, oloss, oidx = orc.vq_assign(z.numpy(), cb.numpy())
    got = idx.cpu().numpy()
    bad = np.flatnonzero(got != oidx)
    assert bad.size == 0, f"{kind}/{K}: {bad.size} indices differ, first at {bad[:5]}: got {got[bad[:5]]} want {oidx[bad[:5]]}; {pc.stats()}"
    assert np.array_equal(zq.cpu().numpy().view(np.uint32), ozq.view(np.uint32))
    assert np.isclose(1.25 * float(sq.item()) / z.numel(), oloss, rtol=LOSS_RTOL)
    st = pc.stats()
    assert st["valid"] == 1 and st["cells"] == 12 ** 4, st


def test_indexed_golden_cases(cg):
    g = load_npz("vq_cases.npz")
    pc = cg.ops.Codebook(torch.from_numpy(g["codebook"]).cuda())
    for name in ("randn", "small", "near", "zeros", "big", "blocky"):
        z = torch.from_numpy(g[f"{name}_z"]).cuda()
        idx, zq, sq = cg.ops.vq_assign(z, pc)
        assert np.array_equal(idx.cpu().numpy(), g[f"{name}_idx"].astype(np.int64)), name
        assert hashlib.sha256(zq.cpu().numpy().tobytes()).digest() == g[f"{name}_zq_sha"].tobytes(), name
        assert np.isclose(1.25 * float(sq.item()) / z.numel(), g[f"{name}_loss"], rtol=LOSS_RTOL), name


@pytest.mark.parametrize("kind", ["init", "gauss", "clumped", "lattice"])
def test_indexed_equals_exhaustive_at_scale(cg, kind):
    """4 M latents per codebook: the two kernels must agree bit for bit (idx, z_q, squared error)."""
    g = torch.Generator().manual_seed(77)
    cb = codebooks(kind, 1024, g).contiguous().cuda()
    B, h, w = 16, 512, 512
    z = latents(cb.cpu(), B * h * w, g).view(B, h, w, 4).permute(0, 3, 1, 2).contiguous().cuda()
    pc = cg.ops.Codebook(cb)
    i1, q1, s1 = cg.ops.vq_assign(z, pc)
    i2, q2, s2 = cg.ops.vq_assign(z, cb)
    assert torch.equal(i1, i2), f"{int((i1 != i2).sum())} of {i1.numel()} differ; {pc.stats()}"
    assert torch.equal(q1.view(torch.int32), q2.view(torch.int32))
    assert np.isclose(float(s1), float(s2), rtol=1e-12)


def test_indexed_bench_workload_stays_in_the_grid(cg):
    """The bench / reference-init codebook: no overflowing cell, lists much shorter than K."""
    import workload
    cbk, _ = workload.codebook_and_counts()
    st = cg.ops.Codebook(cbk.cuda()).stats()
    assert st == dict(valid=1, cells=12 ** 4, max_list=st["max_list"], overflow_cells=0) and 1 <= st["max_list"] <= 63, st


def test_indexed_unusable_codebook_falls_back(cg, orc):
    """Rows the index cannot describe (inf, 1e30): flagged invalid, every latent searched exhaustively."""
    g = torch.Generator().manual_seed(5)
    cb = torch.randn(64, 4, generator=g)
    cb[7, 1] = 1e30
    z = torch.randn(1, 4, 8, 12, generator=g)
    pc = cg.ops.Codebook(cb.cuda())
    assert pc.stats()["valid"] == 0
    i1, q1, _ = cg.ops.vq_assign(z.cuda(), pc)
    i2, q2, _ = cg.ops.vq_assign(z.cuda(), cb.cuda())
    assert torch.equal(i1, i2) and torch.equal(q1.view(torch.int32), q2.view(torch.int32))


def test_indexed_update_and_ragged_shapes(cg, orc):
    """update() after an in-place weight change (what VectorQuantize2 does when `_version` moves);
    w % 4 != 0 takes the generic kernel with the same results."""
    g = torch.Generator().manual_seed(9)
    cb = torch.randn(1024, 4, generator=g).cuda()
    pc = cg.ops.Codebook(cb)
    for shape in ((1, 4, 7, 5), (2, 4, 8, 8), (1, 4, 4, 4), (3, 4, 1, 4)):
        z = torch.randn(*shape, generator=g)
        idx, zq, _ = cg.ops.vq_assign(z.cuda(), pc)
        ozq, _, oidx = orc.vq_assign(z.numpy(), cb.cpu().numpy())
        assert np.array_equal(idx.cpu().numpy(), oidx), shape
        assert np.array_equal(zq.cpu().numpy().view(np.uint32), ozq.view(np.uint32)), shape
    cb.mul_(0.01).add_(0.5)
    pc.update(cb)
    z = (torch.randn(2, 4, 16, 16, generator=g) * 0.01 + 0.5)
    idx, _, _ = cg.ops.vq_assign(z.cuda(), pc)
    _, _, oidx = orc.vq_assign(z.numpy(), cb.cpu().numpy())
    assert np.array_equal(idx.cpu().numpy(), oidx)
    vq = cg.VectorQuantize2(1024, 4, 0.25).cuda().eval()
    with torch.no_grad():
        _, _, i0 = vq(z.cuda())
        vq.embedding.weight.copy_(cb)
        _, _, i1 = vq(z.cuda())
    assert np.array_equal(i1.cpu().numpy(), oidx) and not torch.equal(i0, i1)


def test_module_survives_writes_through_weight_data(cg, orc):
    """`weight.data.copy_()` / `.data.mul_()` (the reference's LitEma.copy_to / restore, ema.py:51,76) do not
    move torch's version counter.  The forward right after such a write must already answer for the LIVE weights (the
    device-side guard switches that call to the exhaustive path); the next one rebuilds the index."""
    g = torch.Generator().manual_seed(13)
    vq = cg.VectorQuantize2(1024, 4, 0.25).cuda().eval()
    z = (torch.rand(2, 4, 16, 24, generator=g) * 2 - 1) / 1024
    new = torch.randn(1024, 4, generator=g) * 3e-4
    with torch.no_grad():
        _, _, i0 = vq(z.cuda())
        v0 = vq.embedding.weight._version
        vq.embedding.weight.data.copy_(new.cuda())
        assert vq.embedding.weight._version == v0                     # the host cannot see the write
        zq1, _, i1 = vq(z.cuda())                                     # guard: exhaustive on the live weights
        torch.cuda.synchronize()
        assert vq._prepared.is_stale()
        zq2, _, i2 = vq(z.cuda())                                     # rebuilt index
        torch.cuda.synchronize()
        assert not vq._prepared.is_stale() and vq._prepared.stats()["valid"] == 1
    ozq, _, oidx = orc.vq_assign(z.numpy(), new.numpy())
    for i, q in ((i1, zq1), (i2, zq2)):
        assert np.array_equal(i.cpu().numpy(), oidx) and np.array_equal(q.cpu().numpy().view(np.uint32), ozq.view(np.uint32))
    assert not torch.equal(i0, i1)
    # invalidate(): the explicit hook; freeze_codebook(): no guard kernel any more
    with torch.no_grad():
        vq.embedding.weight.data.mul_(2.0)
        vq.invalidate()
        _, _, i3 = vq(z.cuda())
    _, _, oidx3 = orc.vq_assign(z.numpy(), (new * 2.0).numpy())
    assert np.array_equal(i3.cpu().numpy(), oidx3)
    vq.freeze_codebook()
    with torch.no_grad():
        _, _, i4 = vq(z.cuda())
    assert torch.equal(i3, i4)
