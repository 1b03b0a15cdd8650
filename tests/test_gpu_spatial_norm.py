"""f4: SpatialNorm (CGIC/modules/vqvae/decoder.py:34-56) -- the fused CUDA op behind cgic_b200.SpatialNorm against
(1) vectors generated from the reference class (tests/golden/spatial_norm.npz), (2) the float64 numpy restatement
(oracle.spatial_norm) on decoder-sized feature maps, (3) torch's eager expression for the gradients.
Floating point: tolerance rtol = atol = 2e-5 against the fp32 reference / float64 oracle (group statistics and the two
1x1 convolutions are summed in a different order than cuDNN / ATen do)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = ATOL = 2e-5


def _golden():
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "spatial_norm.npz"))
    for name, add_conv in zip(d["cases"], d["add_conv"]):
        sd = {k[len(name) + 4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith(f"{name}.sd.")}
        yield str(name), bool(add_conv), d[f"{name}.f"], d[f"{name}.zq"], sd, d[f"{name}.out"]


def test_module_matches_reference_vectors():
    import cgic_b200 as cg
    dev = torch.device("cuda", 0)
    n = 0
    for name, add_conv, f, zq, sd, want in _golden():
        Cc, Cz = f.shape[1], zq.shape[1]
        m = cg.Normalize(Cc, Cz, add_conv)
        m.load_state_dict(sd, strict=True)          # the reference's state-dict keys
        m = m.to(dev).eval()
        with torch.no_grad():
            got = m(torch.from_numpy(f).to(dev), torch.from_numpy(zq).to(dev))
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=ATOL, err_msg=name)
        n += 1
    assert n == 5


@pytest.mark.parametrize("B,Cc,H,W,hz,wz,offset", [
    (2, 512, 64, 64, 16, 16, 0.0),      # the decoder's first blocks at 256 x 256: 16 channels x 64 x 64 per group
    (1, 128, 128, 192, 32, 48, 0.0),    # Kodak shape, later level
    (3, 64, 30, 18, 7, 5, 0.0),         # W % 4 != 0, non-integer factors
    (1, 256, 64, 64, 16, 16, 300.0),    # mean >> spread: the pivoted sums must not cancel
    (1, 32, 4, 4, 1, 1, 0.0),           # slabs smaller than a CTA
    (70, 32, 8, 8, 2, 2, 0.0),          # more slabs than SMs
])
def test_op_matches_oracle(B, Cc, H, W, hz, wz, offset):
    import cgic_b200 as cg
    from oracle import oracle as orc
    g = torch.Generator().manual_seed(B * 1000 + Cc + H)
    f = torch.randn(B, Cc, H, W, generator=g) * 2.0 + offset
    zq = torch.randn(B, 4, hz, wz, generator=g)
    gw, gb, by, bb = (torch.randn(Cc, generator=g) for _ in range(4))
    wy, wb = torch.randn(Cc, 4, 1, 1, generator=g), torch.randn(Cc, 4, 1, 1, generator=g)
    dev = torch.device("cuda", 0)
    got = cg.ops.spatial_norm(f.to(dev), zq.to(dev), gw.to(dev), gb.to(dev), wy.to(dev), by.to(dev), wb.to(dev), bb.to(dev), 32, 1e-6)
    want = orc.spatial_norm(f.numpy(), zq.numpy(), gw.numpy(), gb.numpy(), wy.numpy(), by.numpy(), wb.numpy(), bb.numpy(), 32, 1e-6)
    scale = float(np.abs(want).max())
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=ATOL * max(1.0, scale))
    # no affine / no bias variants
    got2 = cg.ops.spatial_norm(f.to(dev), zq.to(dev), None, None, wy.to(dev), None, wb.to(dev), None, 32, 1e-6)
    want2 = orc.spatial_norm(f.numpy(), zq.numpy(), None, None, wy.numpy(), None, wb.numpy(), None, 32, 1e-6)
    np.testing.assert_allclose(got2.cpu().numpy(), want2, rtol=RTOL, atol=ATOL * max(1.0, float(np.abs(want2).max())))


def test_gradients_match_eager():
    import cgic_b200 as cg
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    m = cg.Normalize(64, 4, False).to(dev)
    f = torch.randn(2, 64, 16, 16, device=dev, requires_grad=True)
    zq = torch.randn(2, 4, 4, 4, device=dev, requires_grad=True)
    # backward differentiates the eager expression, whose 1x1 convolutions cuDNN would run in TF32 by default: fp32 on both sides
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        out = m(f, zq)
        gout = torch.randn_like(out)
        out.backward(gout)
        got = [f.grad.clone(), zq.grad.clone()] + [p.grad.clone() for p in m.parameters()]
        f.grad = zq.grad = None
        m.zero_grad()
        ref = cg.decoder._eager(f, zq, m.norm_layer, m.conv_y, m.conv_b)
        ref.backward(gout)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    want = [f.grad, zq.grad] + [p.grad for p in m.parameters()]
    torch.testing.assert_close(out.detach(), ref.detach(), rtol=1e-4, atol=1e-4)
    for a, b in zip(got, want):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)


def test_errors():
    import cgic_b200 as cg
    dev = torch.device("cuda", 0)
    f = torch.zeros(1, 32, 4, 4, device=dev)
    w = torch.zeros(32, 4, device=dev)
    with pytest.raises(RuntimeError):   # no CPU fallback
        cg.ops.spatial_norm(f.cpu(), torch.zeros(1, 4, 1, 1), None, None, w.cpu(), None, w.cpu(), None, 32, 1e-6)
    with pytest.raises(ValueError):
        cg.ops.spatial_norm(f, torch.zeros(1, 4, 1, 1, device=dev), None, None, torch.zeros(32, 3, device=dev), None, w, None, 32, 1e-6)
    with pytest.raises(cg._lib.CgicError):   # groups must divide C
        cg.ops.spatial_norm(f, torch.zeros(1, 4, 1, 1, device=dev), None, None, w, None, w, None, 5, 1e-6)
    with pytest.raises(cg._lib.CgicError):   # zq channels > 8
        cg.ops.spatial_norm(f, torch.zeros(1, 9, 1, 1, device=dev), None, None, torch.zeros(32, 9, device=dev), None, torch.zeros(32, 9, device=dev), None, 32, 1e-6)
