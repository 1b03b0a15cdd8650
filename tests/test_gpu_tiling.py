"""GPU tests of the tiling driver (BASELINE config 5, inference_high_resolution.py path): batched equal-shape
tiles must give exactly what the reference's loop of B == 1 `model.compress` calls gives -- same per-tile
streams / bpp, same blended reconstruction, same bpp.txt line -- and the hot path of the six DIV2K-shape
tiles must agree with the oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cg():
    import cgic_b200
    assert torch.cuda.is_available()
    return cgic_b200


class _Heads(torch.nn.Module):
    """Stand-in for the out-of-scope CNN encoder: three heads from strided pixels, elementwise only
    (so the result for a tile does not depend on which batch it travels in)."""

    def forward_heads(self, x):
        def head(k, s):
            p = x[:, :, ::k, ::k]
            r, g, b = p[:, 0], p[:, 1], p[:, 2]
            return torch.stack([r - 0.5, g - 0.5, b - 0.5, (r + g + b) * s - 0.5], 1) * (1.8 / 1024)
        return head(16, 0.31), head(8, 0.33), head(4, 0.35)


class _Decoder(torch.nn.Module):
    def forward(self, quant2, quant, mask):
        up = F.interpolate(quant[:, :3] * 512 + 0.5, scale_factor=4, mode="nearest")
        return up * mask[2].repeat_interleave(4, -1).repeat_interleave(4, -2).float().clamp(0.5, 1.0)


def _model(cg, c=0.1, m=0.8):
    import workload
    cbk, counts = workload.codebook_and_counts()
    dd = dict(z_channels=4, router_config=dict(params=dict(coarse_grain_ratio=c, medium_grain_ratio=m)))
    model = cg.CGIC(ddconfig=dd, encoder=_Heads(), decoder=_Decoder()).cuda().eval()
    # the 1x1 convs are stock cuDNN (out of scope) and pick batch-size dependent algorithms (TF32 on by default):
    # identities keep this test about the hot path and the driver
    model.quant_conv = torch.nn.Identity()
    model.post_quant_conv = torch.nn.Identity()
    with torch.no_grad():
        model.quantize.embedding.weight.copy_(cbk)
        for i in range(1024):
            model.quantize.embedding_counter[str(i)].fill_(float(counts[i]))
    return model, cg.HuffmanCoding(model.quantize.embedding_counter), cg.BinaryCoding()


def test_compress_tiled_equals_the_reference_loop(cg, tmp_path):
    inf = cg.inference
    model, h_string, h_mask = _model(cg)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 1550, 1000, generator=g).cuda()         # pads to 1552 x 1008: tiles 768x768, 768x240, 768x768, 768x240, 16x768, 16x240
    d_batched, d_loop = tmp_path / "a", tmp_path / "b"
    d_batched.mkdir(), d_loop.mkdir()
    x_rec, bpp_image, tiles = inf.compress_tiled(model, x, h_string, h_mask, str(d_batched))
    # the reference's loop (inference_high_resolution.py:226-257), tile by tile with B == 1
    pad, unpad = inf.compute_padding(1550, 1000, min_div=16)
    xp = F.pad(x, pad, mode="constant", value=0)
    h_list, w_list, th, tw = inf.nonoverlapping_grid_indices(xp)
    rec = torch.zeros(xp.shape, device=x.device)
    contributors = torch.zeros(xp.shape, device=x.device)
    bit_sum, k = 0.0, 0
    for i in range(len(h_list)):
        for j in range(len(w_list)):
            hi, wi, a, b = h_list[i], w_list[j], th[i], tw[j]
            wts = inf.gaussian_weights(b, a, 1, x.device)
            with torch.no_grad():
                t_rec, bpp, _ = model.compress(xp[:, :, hi:hi + a, wi:wi + b], str(d_loop), h_string, h_mask, False)
            rec[:, :, hi:hi + a, wi:wi + b] += t_rec * wts
            contributors[:, :, hi:hi + a, wi:wi + b] += wts
            bit_sum += bpp * b * a
            assert tiles[k]["bpp"] == bpp and (tiles[k]["y"], tiles[k]["x"], tiles[k]["h"], tiles[k]["w"]) == (hi, wi, a, b), k
            k += 1
    rec /= contributors
    rec = F.pad(rec.clamp(0, 1), unpad)
    assert bpp_image == bit_sum / 1000 / 1550
    assert torch.equal(x_rec, rec) and x_rec.shape == x.shape
    for name in cg.ops.STREAM_NAMES:                              # the files left behind are the last tile's
        fa, fb = d_batched / (name + ".bin"), d_loop / (name + ".bin")
        assert fa.exists() == fb.exists() and (not fa.exists() or fa.read_bytes() == fb.read_bytes()), name
    # rank-sharded: two ranks' partial sums add up to the same bits
    parts = [inf.compress_tiled(model, x, h_string, h_mask, None, rank=r, world=2) for r in range(2)]
    got = sum(t["bpp"] * t["h"] * t["w"] for p in parts for t in p[2] if t is not None)
    assert got == bit_sum and all((parts[0][2][i] is None) != (parts[1][2][i] is None) for i in range(len(tiles)))


def test_run_writes_bpp_txt_like_the_reference(cg, tmp_path):
    model, h_string, h_mask = _model(cg)
    g = torch.Generator().manual_seed(4)
    imgs = [torch.rand(1, 3, 64, 96, generator=g), torch.rand(1, 3, 64, 96, generator=g)]
    avg = cg.inference.run(model, imgs, str(tmp_path), h_string, h_mask)
    lines = (tmp_path / "bpp.txt").read_text().split("\n")
    assert len(lines) == 3 and lines[0].startswith("image: 0 \t bpp: ") and lines[2] == f"Bpp Average: {avg}"
    bpps = [float(l.split("bpp: ")[1]) for l in lines[:2]]
    assert avg == sum(bpps) / 2
    avg_hr = cg.inference.run(model, imgs, str(tmp_path / "hr"), h_string, h_mask, high_resolution=True)
    assert avg_hr == avg                                          # a 64x96 image is a single tile


def test_config5_tiles_hot_path_vs_oracle(cg):
    """DIV2K shape 2032 x 1344 as its six tiles (four shape groups), hot path only: every group one launch;
    two tiles also through the oracle end to end; round trip and stream-size identities for all."""
    import workload
    from oracle import oracle as orc
    inf = cg.inference
    cbk, counts = workload.codebook_and_counts()
    order = workload.lexicographic_order()
    table = cg.ops.HuffTable(counts.numpy(), order)
    cb = cbk.cuda()
    prepared = cg.ops.Codebook(cb)
    plan = inf.tile_plan(1344, 2032)
    groups = inf.group_tiles(plan)
    zs, ms, px, keep = [], [], [], []
    for gi, ((th, tw), members) in enumerate(groups.items()):
        n = len(members)
        e16, e8 = workload.entropy_maps(n, th, tw, 50 + gi)
        mc, mm, mf, _, mode = cg.ops.router(e16.cuda(), e8.cuda(), 0.1, 0.8, per_image=True)
        hc, hm, hf = (t.cuda() for t in workload.heads(n, th, tw, cbk, 50 + gi))
        zs.append(cg.ops.mask_mix(hc, hm, hf, mc, mm, mf))
        ms.append((mc, mm, mf))
        px.append((th, tw))
        keep.append((e16, e8))
    bpp_image, outs = inf.tiled_hot_path(zs, ms, mode, table, cb, px, 1344 * 2032, prepared=prepared)
    lens = table.lengths().astype(np.int64)
    bits = 0
    ot = orc.huff_build(counts.numpy(), order)
    for gi, (o, (th, tw)) in enumerate(zip(outs, px)):
        h, w = th // 4, tw // 4
        n = o["sizes"].shape[0]
        assert int(o["status"].abs().sum()) == 0 and torch.equal(o["ind"].view(-1), o["idx"])
        want = workload.expected_counts(th, tw, 0.1, 0.8)
        idx = o["idx"].view(n, h, w).cpu().numpy()
        masks = [m.cpu().numpy() for m in ms[gi]]
        for b in range(n):
            sel = [idx[b, ::4, ::4][masks[0][b, 0] == 1], idx[b, ::2, ::2][masks[1][b, 0] == 1], idx[b][masks[2][b, 0] == 1]]
            assert tuple(len(s) for s in sel) == want
            for s in range(3):
                assert int(o["sizes"][b, s]) == int(lens[sel[s]].sum()) // 8 + 2
            bits += int(o["sizes"][b].sum()) * 8
        if gi in (1, 3):                                          # 768x496 and 576x496: full oracle chain
            e16, e8 = keep[gi]
            omc, omm, omf, omode = orc.router(e16[:1].numpy(), e8[:1].numpy(), 0.1, 0.8)
            ozq, _, oidx = orc.vq_assign(zs[gi][:1].cpu().numpy(), cbk.numpy())
            assert np.array_equal(oidx, o["idx"].view(n, -1)[0].cpu().numpy())
            streams = orc.pack_image(ot, oidx.reshape(h, w), omc[0, 0], omm[0, 0], omf[0, 0], omode)
            offs, _, _ = table.layout(h, w)
            blob = o["bytes"][0].cpu().numpy()
            for s in range(5):
                assert blob[offs[s]: offs[s] + int(o["sizes"][0, s])].tobytes() == streams[s], (gi, s)
    assert bpp_image == bits / (1344 * 2032)
