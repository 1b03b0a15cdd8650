"""Decoder entry -- drop-ins for the two pieces of CGIC/modules/vqvae/decoder.py that touch the decoded latents
and masks; the CNN around them stays the reference's.

`SpatialNorm` (decoder.py:34-53): same constructor, same sub-modules and state-dict keys (`norm_layer.*`, `conv_y.*`,
`conv_b.*`, `conv.*` with add_conv), so reference checkpoints load unchanged.  Under `torch.no_grad()` / inference the
forward is one fused CUDA op (cgic_spatial_norm: group statistics + both 1x1 convolutions evaluated on the fly at zq's
own resolution -- neither the up-sampled zq nor the two [B,C,H,W] convolution outputs exist).  When gradients are
required the same forward runs, and backward differentiates the reference's eager expression (recomputed on the GPU).
`Normalize` (decoder.py:55-56) and `merge` (decoder.py:373-382, the mask-gated merge) complete the entry.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


def _eager(f, zq, norm_layer, conv_y, conv_b):
    zq = torch.nn.functional.interpolate(zq, size=f.shape[-2:], mode="nearest")
    return norm_layer(f) * conv_y(zq) + conv_b(zq)


class _SpatialNormFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, zq, module, *params):
        ctx.module = module
        ctx.save_for_backward(f, zq)
        return module._fused(f.detach(), zq.detach())

    @staticmethod
    def backward(ctx, g):
        f, zq = ctx.saved_tensors
        m = ctx.module
        params = [p for p in m._fused_params()]
        with torch.enable_grad():
            fi = f.detach().requires_grad_(ctx.needs_input_grad[0])
            zi = zq.detach().requires_grad_(ctx.needs_input_grad[1])
            out = _eager(fi, zi, m.norm_layer, m.conv_y, m.conv_b)
            wanted = [t for t in [fi, zi] + params if t.requires_grad]
            grads = iter(torch.autograd.grad(out, wanted, g, allow_unused=True))
        res = [next(grads) if t.requires_grad else None for t in [fi, zi]]
        res.append(None)
        res += [next(grads) if p.requires_grad else None for p in params]
        return tuple(res)


class SpatialNorm(nn.Module):
    def __init__(self, f_channels, zq_channels, norm_layer=nn.GroupNorm, freeze_norm_layer=False, add_conv=False, **norm_layer_params):
        super().__init__()
        self.norm_layer = norm_layer(num_channels=f_channels, **norm_layer_params)
        if freeze_norm_layer:
            for p in self.norm_layer.parameters():     # (the reference iterates `.parameters` without calling it and would raise)
                p.requires_grad = False
        self.add_conv = add_conv
        if self.add_conv:
            self.conv = nn.Conv2d(zq_channels, zq_channels, kernel_size=3, stride=1, padding=1)
        self.conv_y = nn.Conv2d(zq_channels, f_channels, kernel_size=1, stride=1, padding=0)
        self.conv_b = nn.Conv2d(zq_channels, f_channels, kernel_size=1, stride=1, padding=0)

    def _fused_params(self):
        n = self.norm_layer
        return [p for p in (n.weight, n.bias, self.conv_y.weight, self.conv_y.bias, self.conv_b.weight, self.conv_b.bias) if p is not None]

    def _fused(self, f, zq):
        n = self.norm_layer
        return ops.spatial_norm(f, zq, n.weight, n.bias, self.conv_y.weight, self.conv_y.bias, self.conv_b.weight, self.conv_b.bias,
                                n.num_groups, n.eps)

    def forward(self, f, zq):
        if not isinstance(self.norm_layer, nn.GroupNorm):
            raise TypeError("the B200 SpatialNorm kernel implements norm_layer=nn.GroupNorm (the only one Control-GIC uses)")
        if self.add_conv:
            # the 3x3 convolution acts on the UP-SAMPLED zq (decoder.py:49-51): up-sample first, the kernel then reads zq 1:1
            zq = self.conv(torch.nn.functional.interpolate(zq, size=f.shape[-2:], mode="nearest"))
        params = self._fused_params()
        if torch.is_grad_enabled() and (f.requires_grad or zq.requires_grad or any(p.requires_grad for p in params)):
            return _SpatialNormFunction.apply(f, zq, self, *params)
        return self._fused(f, zq)


def Normalize(in_channels, zq_ch, add_conv):
    """decoder.py:55-56"""
    return SpatialNorm(in_channels, zq_ch, norm_layer=nn.GroupNorm, freeze_norm_layer=False, add_conv=add_conv, num_groups=32, eps=1e-6, affine=True)


def merge(h, other, mask, level: int):
    """decoder.py:373-382: level 2: h*up2(mask[0]) + other*mask[1]; level 3: h*up4(mask[0]) + h*up2(mask[1]) + other*mask[2]."""
    return ops.decoder_merge(h, other, mask, level)
