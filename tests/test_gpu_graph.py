"""The hot path under CUDA-graph capture: the captured step must consist of the library's kernels only.  A workspace
created DURING capture would be zero-filled by a kernel that becomes part of the graph (one extra ~2 us node per op and
per replay, and a full instead of a programmatic edge): ops._workspace reuses the warm-up's buffers while capturing."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_capture_reuses_warmup_workspaces_and_replays_bit_exact():
    import cgic_b200 as cg
    import workload
    dev = torch.device("cuda", 0)
    B, H, W = 4, 64, 64
    h, w = H // 4, W // 4
    cbk, counts = workload.codebook_and_counts()
    table = cg.ops.HuffTable(counts.numpy(), workload.lexicographic_order()).upload()
    cb = cbk.to(dev)
    prepared = cg.ops.Codebook(cb)
    e16, e8 = workload.entropy_maps(B, H, W, 31)
    mc, mm, mf, _, mode = cg.ops.router(e16.to(dev), e8.to(dev), 0.1, 0.8, per_image=True)
    hc, hm, hf = (t.to(dev) for t in workload.heads(B, H, W, cbk, 31))
    z = cg.ops.mask_mix(hc, hm, hf, mc, mm, mf)

    def step():
        idx, zq, sq = cg.ops.vq_assign(z, prepared)
        packed, sizes = cg.ops.pack(idx, mc, mm, mf, mode, table, h, w)
        return (idx, zq, packed, sizes) + tuple(cg.ops.unpack(packed, sizes, mode, table, cb, h, w))

    eager = [t.clone() for t in step()]
    torch.cuda.synchronize()
    before = {k: v.data_ptr() for k, v in cg.ops._ws_cache.items()}
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):                      # torch captures on a stream of its own: every workspace key misses
        out = step()
    new = {k: v.data_ptr() for k, v in cg.ops._ws_cache.items() if k not in before}
    assert new, "the capture stream should have registered its own workspace keys"
    assert set(new.values()) <= set(before.values()), "a workspace was allocated (and zero-filled) inside the capture"
    z.add_(0)                                      # replay twice: the kernels leave the workspaces clean
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(eager, out)):
        if i == 2:      # packed slots: only the first sizes[b, s] bytes of a slot are defined (the rest is torch.empty memory)
            offs, _, _ = table.layout(h, w)
            sz = out[3].cpu()
            for img in range(B):
                for st in range(5):
                    lo, n = int(offs[st]), int(sz[img, st])
                    assert torch.equal(a[img, lo:lo + n], b[img, lo:lo + n]), (img, st)
        else:
            assert torch.equal(a, b), i
    assert int(out[-1].abs().sum()) == 0
