"""VectorQuantize2 -- drop-in for CGIC/modules/vqvae/quantize.py:9-98 backed by cgic_vq_assign_indexed
(prepared codebook, csrc/codebook.cu; cgic_vq_assign is the exhaustive search with identical results).

Same constructor, attributes (`embedding`, `embedding_counter`, n_e, e_dim, beta, legacy) and
state-dict keys (`embedding.weight`, `embedding_counter.<i>` with shape [1]); forward returns
(z_q, loss, z_indices) with z_indices int64 flat [B*h*w] exactly like the reference.
Differences, all deliberate: the counters are not moved to CUDA at construction (the reference's
`.cuda()` at quantize.py:28 makes CPU construction impossible); `remap` / `sane_index_shape`
are accepted and, as in the reference's forward, unused.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class _VQFunction(torch.autograd.Function):
    """Forward = the CUDA kernels; backward = straight-through for z_q (quantize.py:93) plus the
    analytic gradient of the commitment loss (quantize.py:85-90)."""

    @staticmethod
    def forward(ctx, z, weight, beta, legacy, prepared=None):
        idx, zq, sq = ops.vq_assign(z.detach(), prepared if prepared is not None else weight.detach())
        mean = (sq / z.numel()).to(torch.float32)[0]
        loss = mean + beta * mean if legacy else beta * mean + mean
        ctx.save_for_backward(z, weight, idx)
        ctx.beta, ctx.legacy = beta, legacy
        ctx.mark_non_differentiable(idx)
        return zq, loss, idx

    @staticmethod
    def backward(ctx, g_zq, g_loss, _g_idx):
        z, weight, idx = ctx.saved_tensors
        B, C, h, w = z.shape
        e = weight[idx].view(B, h, w, C).permute(0, 3, 1, 2)
        diff = (z - e) * (2.0 / z.numel())
        wz, we = (1.0, ctx.beta) if ctx.legacy else (ctx.beta, 1.0)
        g_z = g_w = None
        if ctx.needs_input_grad[0]:
            g_z = g_zq + g_loss * wz * diff
        if ctx.needs_input_grad[1]:
            g_w = torch.zeros_like(weight).index_add_(0, idx, (-g_loss * we * diff).permute(0, 2, 3, 1).reshape(-1, C))
        return g_z, g_w, None, None, None


class VectorQuantize2(nn.Module):
    def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
        super().__init__()
        if e_dim != 4:
            raise ValueError("the B200 VQ kernel implements e_dim == 4 (the only value Control-GIC uses)")
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        self.legacy = legacy
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)                      # quantize.py:25-26
        # plain dict -> torch sorts the keys: iteration order "0","1","10","100",... (quantize.py:28)
        self.embedding_counter = nn.ParameterDict(
            {str(i): nn.Parameter(torch.zeros(1)) for i in range(n_e)}).requires_grad_(False)
        self.remap = remap
        self.unknown_index = unknown_index
        self.re_embed = n_e
        self.sane_index_shape = sane_index_shape

    # -- counters as one tensor (the reference touches them one .item() at a time) ---------
    def counter_order(self):
        """Symbols in the ParameterDict's iteration order = the Huffman heap push order."""
        return [int(k) for k in self.embedding_counter.keys()]

    def counters_flat(self) -> torch.Tensor:
        """fp32 [n_e] indexed by symbol."""
        vals = torch.cat([p.detach().reshape(1) for p in self.embedding_counter.values()])
        flat = torch.empty_like(vals)
        flat[torch.as_tensor(self.counter_order(), device=vals.device)] = vals
        return flat

    def _flat_counter_storage(self, device) -> torch.Tensor:
        """One fp32 [n_e] tensor (indexed by symbol) that the 1024 `embedding_counter.<i>` parameters are VIEWS of, so
        that the histogram kernel updates all of them in one launch.  `.to()` / `.cuda()` / `load_state_dict` give the
        parameters storage of their own again; the aliasing is then re-established here (metadata only, no kernels
        beyond one gather)."""
        flat = getattr(self, "_counter_flat", None)
        params = self.embedding_counter
        ok = flat is not None and flat.device == device and all(
            p.device == device and p.data_ptr() == flat.data_ptr() + 4 * int(k) for k, p in params.items())
        if not ok:
            flat = self.counters_flat().to(device=device, dtype=torch.float32).contiguous()
            for k, p in params.items():
                p.data = flat[int(k): int(k) + 1]
            self._counter_flat = flat
        return flat

    @torch.no_grad()
    def _bump_counters(self, idx: torch.Tensor) -> None:                                 # quantize.py:79-81
        ops.vq_count(idx, self._flat_counter_storage(idx.device))

    def invalidate(self) -> None:
        """Forget the prepared codebook: call after writing `embedding.weight.data` directly."""
        self._prepared_key = None

    def freeze_codebook(self, frozen: bool = True) -> None:
        """Promise that `embedding.weight` will not be written through `.data` any more (inference): skips the
        per-forward staleness guard (one ~2 us kernel).  invalidate() still forces a rebuild."""
        self._frozen = frozen

    def prepared_codebook(self) -> "ops.Codebook":
        """The search index of `embedding.weight` (ops.Codebook), rebuilt on the current stream whenever the weight
        tensor was replaced or written in place (its `_version` moved).  Writes through `weight.data` (LitEma.copy_to /
        restore of the reference, CGIC/models/ema.py:51,76) move neither: unless freeze_codebook() was called, every
        call also enqueues a device-side comparison of the live weights with the index's copy -- on a mismatch the
        search of THAT call already runs exhaustively on the live weights (same results), and the next call, which sees
        the flag the kernel raised, rebuilds the index."""
        w = self.embedding.weight
        key = (w.data_ptr(), w._version, w.device)
        cb = getattr(self, "_prepared", None)
        if getattr(self, "_prepared_key", None) != key or (cb is not None and cb.is_stale()):
            if cb is None or cb.device != w.device or cb.K != w.shape[0]:
                self._prepared = ops.Codebook(w)
            else:
                cb.update(w)
            self._prepared_key = key
        elif not getattr(self, "_frozen", False):
            cb.check(w)
        return self._prepared

    def forward(self, z):
        w = self.embedding.weight
        # the index costs ~1.5 ms to build: worth it when the weights stand still (inference), not when every
        # optimiser step moves them (training) -- there the exhaustive kernel (~37 us per 262 k latents) is used
        prepared = None if self.training else self.prepared_codebook()
        if torch.is_grad_enabled() and (z.requires_grad or w.requires_grad):
            z_q, loss, idx = _VQFunction.apply(z, w, self.beta, self.legacy, prepared)
        else:
            idx, z_q, sq = ops.vq_assign(z, prepared if prepared is not None else w)
            mean = (sq / z.numel()).to(torch.float32)[0]
            loss = mean + self.beta * mean if self.legacy else self.beta * mean + mean
        if self.training:
            self._bump_counters(idx)
        return z_q, loss, idx
