"""CGIC -- the reference's model-level API (CGIC/models/model.py:22-401) around the B200 hot path.

Keeps `encode(x)`, `decode(quant, mask)`, `compress(input, path, h_indices, h_mask, save_img)`,
`forward(input)` with the reference's signatures and return values, without the Lightning
dependency.  Everything between the CNN encoder heads and the CNN decoder runs in the sm_100a
kernels of libcgic_b200.so:

    Entropy(8)+Entropy(16)  -> cgic_entropy_maps      (model.py:100-101, 433-483)
    router                  -> cgic_router            (vqvae_blocks.py:354-355, RouterTriple.py)
    mask-mix                -> cgic_mask_mix          (vqvae_blocks.py:361-366)
    quant_conv              -> torch 1x1 conv (kept: fusing it would change z's bits)
    quantize                -> cgic_vq_assign         (quantize.py:69-98)
    select + 5-stream pack  -> cgic_pack              (model.py:217-260)
    unpack + re-assembly + gather -> cgic_unpack      (model.py:269-392)

The CNN encoder / decoder are out of scope (stock PyTorch): pass any module as `encoder` that
offers `forward_heads(x) -> (h_coarse, h_medium, h_fine)`, or an unmodified reference `Encoder`
(its three `conv_out*` heads are tapped with forward hooks); `decoder(quant2, quant, mask)` is
called exactly as model.py:114-117 does.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
from torch import nn

from . import ops
from .codec import BinaryCoding, HuffmanCoding
from .entropy import Entropy, entropy_pair
from .quantize import VectorQuantize2 as VectorQuantizer
from .router import TripleGrainFixedEntropyRouter


class ReferenceEncoderHeads(nn.Module):
    """Adapter around an UNMODIFIED reference `Encoder` (CGIC/modules/vqvae/vqvae_blocks.py):
    runs it and taps the outputs of conv_out_coarse / conv_out / conv_out_fine, i.e. the three
    heads right before the reference's own router + mask-mix tail (whose result is discarded)."""

    def __init__(self, encoder: nn.Module):
        super().__init__()
        self.encoder = encoder

    def forward_heads(self, x, x_entropy_p16, x_entropy_p8):
        taps = {}
        hooks = [getattr(self.encoder, name).register_forward_hook(lambda m, i, o, key=key: taps.__setitem__(key, o))
                 for key, name in (("c", "conv_out_coarse"), ("m", "conv_out"), ("f", "conv_out_fine"))]
        try:
            self.encoder(x, x_entropy_p16, x_entropy_p8)
        finally:
            for hk in hooks:
                hk.remove()
        return taps["c"], taps["m"], taps["f"]


def reference_cnns(ddconfig, embed_dim):
    """The out-of-scope CNNs built from the reference package itself -- `Encoder(**ddconfig)`, `Decoder(zq_ch=embed_dim,
    **ddconfig)` exactly as CGIC/models/model.py:42-43 does -- when that package can be imported: already on sys.path, or
    under $CGIC_REFERENCE, or in this repository's git-ignored baseline/_ref copy.  Returns (encoder, decoder) or None.
    `import pytorch_lightning` (which the reference's blocks do not need, only its model.py) is never touched."""
    import importlib
    import sys
    roots = [None, os.environ.get("CGIC_REFERENCE"),
             os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")]
    for root in roots:
        if root is not None and not os.path.isdir(os.path.join(root, "CGIC", "modules", "vqvae")):
            continue
        if root is not None:
            sys.path.insert(0, root)
        try:
            blocks = importlib.import_module("CGIC.modules.vqvae.vqvae_blocks")
            dec = importlib.import_module("CGIC.modules.vqvae.decoder")
            return blocks.Encoder(**ddconfig), dec.Decoder(zq_ch=embed_dim, **ddconfig)
        except ImportError:
            continue
        finally:
            if root is not None and sys.path and sys.path[0] == root:
                sys.path.pop(0)
    return None


def _heads_arity(encoder) -> int:
    """How many positional arguments `encoder.forward_heads` takes: 3 = (x, e16, e8) like the reference Encoder's
    forward, 1 = (x).  Decided once from the signature (no try / except around the CNN: an error raised inside it
    must surface as it is, not trigger a second run of the whole encoder)."""
    import inspect
    if not hasattr(encoder, "forward_heads"):
        raise TypeError("encoder must provide forward_heads(x[, e16, e8]) -> (h_coarse, h_medium, h_fine); wrap a reference "
                        "Encoder in ReferenceEncoderHeads")
    params = [p for p in inspect.signature(encoder.forward_heads).parameters.values()
              if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
    if any(p.kind == p.VAR_POSITIONAL for p in inspect.signature(encoder.forward_heads).parameters.values()):
        return 3
    return 3 if len(params) >= 3 else 1


class CGIC(nn.Module):
    def __init__(self, ddconfig=None, n_embed=1024, embed_dim=4, learning_rate=None, lossconfig=None, ckpt_path=None,
                 ignore_keys=(), image_key="image", colorize_nlabels=None, monitor=None, remap=None,
                 sane_index_shape=False, ema_decay=None, image_size=256, entropy_patch_size=(8, 16),
                 encoder: Optional[nn.Module] = None, decoder: Optional[nn.Module] = None):
        super().__init__()
        ddconfig = dict(ddconfig or {})
        self.image_key = image_key
        if encoder is None or decoder is None:
            # like the reference (model.py:42-43): build the CNNs from ddconfig -- with the reference's own classes, which stay
            # out of scope here (stock PyTorch); explicit `encoder=` / `decoder=` arguments take precedence
            built = reference_cnns(ddconfig, embed_dim) if ddconfig else None
            if built is None:
                raise ValueError("the reference package (CGIC.modules.vqvae) is not importable, so the out-of-scope CNNs cannot be built "
                                 "from ddconfig: pass `encoder` and `decoder` modules, e.g. ReferenceEncoderHeads(Encoder(**ddconfig)) and "
                                 "Decoder(zq_ch=embed_dim, **ddconfig), or put the reference on sys.path / $CGIC_REFERENCE")
            encoder = encoder if encoder is not None else built[0]
            decoder = decoder if decoder is not None else built[1]
        self.encoder = encoder if hasattr(encoder, "forward_heads") else ReferenceEncoderHeads(encoder)
        self.decoder = decoder
        self._heads_arity = _heads_arity(self.encoder)
        if learning_rate is not None:
            self.learning_rate = learning_rate
        self.quantize = VectorQuantizer(n_embed, embed_dim, beta=0.25, remap=remap, sane_index_shape=sane_index_shape)
        z_channels = ddconfig.get("z_channels", embed_dim)
        self.quant_conv = nn.Conv2d(z_channels, embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, z_channels, 1)
        self.entropy_calculation_p8 = Entropy(entropy_patch_size[0]).eval()
        self.entropy_calculation_p16 = Entropy(entropy_patch_size[1]).eval()
        # like vqvae_blocks.py:299,354 the router is re-instantiated from this dict on every
        # forward, so editing router_config["params"] changes the granularity ratio at once
        self.router_config = ddconfig.get("router_config", {"params": {"coarse_grain_ratio": 0.1, "medium_grain_ratio": 0.8}})
        if monitor is not None:
            self.monitor = monitor
        self.use_ema = False
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    def init_from_ckpt(self, path, ignore_keys=()):
        ckpt = torch.load(path, map_location="cpu")
        sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        self.load_state_dict(sd, strict=False)

    # ---------------------------------------------------------------- model.py:99-112
    def _heads(self, x, e16, e8):
        if self._heads_arity == 3:
            return self.encoder.forward_heads(x, e16, e8)
        return self.encoder.forward_heads(x)

    def _router(self, per_image=False):
        p = self.router_config["params"]
        return TripleGrainFixedEntropyRouter(p["coarse_grain_ratio"], p["medium_grain_ratio"], per_image=per_image)

    def encode(self, x, per_image: bool = False):
        """model.py:99-112.  With per-image thresholds (per_image=True, or a single image, where the reference's batch-wide
        thresholds ARE per-image) the tail runs as the two launches of SURVEY 8f f1: entropy maps + routing, then fine mask +
        gate + mask-mix; `self.threshold_adjacent` [B,2] then counts the coarse / medium entropies within the Entropy kernel's
        float tolerance of a threshold (all zero = masks provably those of the reference)."""
        router = self._router(per_image)
        if per_image or x.shape[0] == 1:
            x_entropy_p8, x_entropy_p16, m_c, m_m, near, mode = ops.entropy_route(x, router.coarse_grain_ratio, router.medium_grain_ratio)
            self.threshold_adjacent = near
            h_coarse, h_medium, h_fine = self._heads(x, x_entropy_p16, x_entropy_p8)
            m_f, gate, h = ops.route_mix(h_coarse, h_medium, h_fine, m_c, m_m, mode, want_gate=True)
            grain_mask = [m_c, m_m, m_f]
            ratios = [router.coarse_grain_ratio, router.medium_grain_ratio, router.fine_grain_ratio]
        else:
            x_entropy_p8, x_entropy_p16 = entropy_pair(x)
            self.threshold_adjacent = None
            h_coarse, h_medium, h_fine = self._heads(x, x_entropy_p16, x_entropy_p8)
            grain_mask, gate, ratios, mode = router(x_entropy_p16, x_entropy_p8)
            h = ops.mask_mix(h_coarse, h_medium, h_fine, *grain_mask)
        # vqvae_blocks.py:357-359 (quirk Q6: argmax over the width-concatenated axis; kept as is)
        grain_indices = gate.permute(0, 3, 1, 2).argmax(dim=1)
        h = self.quant_conv(h)
        quant, emb_loss, ind = self.quantize(h)
        return quant, emb_loss, grain_indices, grain_mask, ind, ratios, mode

    def decode(self, quant, mask):
        quant2 = self.post_quant_conv(quant)
        return self.decoder(quant2, quant, mask)

    def forward(self, input):
        quant, diff, grain_indices, grain_mask, _, _, _ = self.encode(input)
        dec = self.decode(quant, grain_mask)
        return dec, diff, grain_indices

    # ---------------------------------------------------------------- model.py:206-401
    @staticmethod
    def _reference_dtypes(mode: int, masks: List[torch.Tensor], ind: torch.Tensor):
        """dtype quirks of grain_mask_decompress / ind_decompress per mode (model.py:282-389)."""
        if mode in (1, 2, 3):
            absent = {1: 0, 2: 1, 3: 2}[mode]
            masks[absent] = masks[absent].to(torch.float32)
        elif mode >= 4:
            masks = [m.to(torch.int32) for m in masks]
            ind = ind.to(torch.int32)
        return masks, ind

    def compress(self, input, path, h_indices: HuffmanCoding, h_mask: BinaryCoding, save_img):
        assert len(input.shape) == 4
        if input.shape[0] != 1:
            raise ValueError("compress() follows the reference and handles one image per call; use compress_batch()")
        h_indices = self._native_coder(h_indices)
        out = self.compress_batch(input, h_indices, per_image=False)
        sizes = out["sizes_host"][0].tolist()
        blob = out["bytes"][0].cpu().numpy()
        offs, _, _ = h_indices.table.layout(*out["grid"])
        for s, name in enumerate(ops.STREAM_NAMES):
            if ops.stream_present(out["mode"], s):
                with open(os.path.join(path, name + ".bin"), "wb") as f:
                    f.write(blob[offs[s]: offs[s] + sizes[s]].tobytes())
        partition_map = None  # drawing (CGIC/modules/draw.py) is out of scope; save_img is accepted and ignored
        return out["dec"], out["bpp"][0], partition_map

    def _native_coder(self, h_indices):
        """A caller that built the REFERENCE's HuffmanCoding (inference.py:137-139 does, from `model.quantize.embedding_counter`)
        gets the native table for the same counters; its codes are checked against the object passed in, so a coder built
        from other frequencies is refused rather than silently replaced."""
        if hasattr(h_indices, "table"):
            return h_indices
        own = getattr(self, "_own_coder", None)
        if own is None or own[0] is not h_indices:
            native = HuffmanCoding(self.quantize.embedding_counter)
            theirs = getattr(h_indices, "codes", None)
            if not isinstance(theirs, dict) or {int(k): v for k, v in theirs.items()} != native.codes:
                raise TypeError("h_indices is neither a cgic_b200 HuffmanCoding nor a coder whose codes are those of this model's "
                                "embedding_counter")
            own = (h_indices, native)
            self._own_coder = own
        return own[1]

    def compress_batch(self, input, h_indices: HuffmanCoding, per_image: bool = True, decode: bool = True):
        """B independent images in one pass (per-image router thresholds when per_image=True):
        returns dict(bytes [B,stride] uint8 cuda, sizes [B,5], sizes_host, bpp list, mode, grid,
        ind, quant, dec, masks, ind_decompress, quant_decompress, masks_decompress)."""
        quant, diff, grain_indices, grain_mask, ind, _, mode = self.encode(input, per_image=per_image)
        B, _, H, W = input.shape
        h, w = quant.shape[-2:]
        h_indices = self._native_coder(h_indices)
        table = h_indices.table
        packed, sizes = ops.pack(ind, *grain_mask, mode, table, h, w)
        out = dict(bytes=packed, sizes=sizes, mode=mode, grid=(h, w), ind=ind, quant=quant, masks=grain_mask, emb_loss=diff)
        if decode:
            mc, mm, mf, ind_dec, quant_dec, status = ops.unpack(packed, sizes, mode, table, self.quantize.embedding.weight, h, w)
            out.update(ind_decompress=ind_dec, quant_decompress=quant_dec)
        sizes_host = sizes.cpu()                                  # the one host sync of the call
        if int(sizes_host.min()) < 0:
            raise KeyError("index outside the Huffman table")
        out["sizes_host"] = sizes_host
        out["bpp"] = [int(sizes_host[b].sum()) * 8 / (H * W) for b in range(B)]               # model.py:233
        if decode:
            if int(status.max()) != 0:
                raise RuntimeError("corrupt stream: symbol count does not match the mask population")
            masks, ind_dec = self._reference_dtypes(mode, [mc, mm, mf], ind_dec)
            if not self.training:                                                              # model.py:394-397
                masks = [m.unsqueeze(1) for m in masks]
            out["masks_decompress"] = masks
            out["ind_decompress"] = ind_dec
            out["dec"] = self.decode(quant_dec, masks)
        return out

    def decompress_files(self, path, H, W, mode, h_indices: HuffmanCoding):
        """Decoder-side entry the reference lacks: the five .bin files of one image -> reconstruction."""
        h, w = H // 4, W // 4
        table = h_indices.table
        offs, caps, stride = table.layout(h, w)
        dev = self.quantize.embedding.weight.device
        blob = torch.zeros(1, stride, dtype=torch.uint8)
        sizes = torch.zeros(1, 5, dtype=torch.int32)
        for s, name in enumerate(ops.STREAM_NAMES):
            fn = os.path.join(path, name + ".bin")
            if ops.stream_present(mode, s) and os.path.exists(fn):
                data = open(fn, "rb").read()
                if len(data) > caps[s]:
                    raise ValueError(f"{fn}: {len(data)} bytes exceed the slot capacity {caps[s]}")
                blob[0, offs[s]: offs[s] + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8)
                sizes[0, s] = len(data)
        mc, mm, mf, ind, quant, status = ops.unpack(blob.to(dev), sizes.to(dev), mode, table, self.quantize.embedding.weight, h, w)
        if int(status.max()) != 0:
            raise RuntimeError("corrupt stream")
        masks, ind = self._reference_dtypes(mode, [mc, mm, mf], ind)
        masks = [m.unsqueeze(1) for m in masks]
        return self.decode(quant, masks), ind, masks

    def get_last_layer(self):
        return self.decoder.conv_out.weight
