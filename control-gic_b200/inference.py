"""Entry points of the reference's `inference.py` and `inference_high_resolution.py` on the B200 path.

Same command-line flags (`-i -b -n -s -o -w -r`, inference.py:112-124 /
inference_high_resolution.py:176-194), the same `configs/config_inference.yaml` keys (PyYAML here;
the reference uses OmegaConf), the same tiling of large images into non-overlapping 768-pixel
tiles (inference_high_resolution.py:112-125, 226-262), the same outputs: `bpp.txt` lines
(`image: {i} \\t bpp: {bpp}`, `Bpp Average: ...`) and the five `.bin` files per call.

What changes is how the work is issued: the reference calls `model.compress` once per tile with
B = 1; here all tiles of an image that share a shape go through ONE launch of the batched hot
path (`CGIC.compress_batch(per_image=True)`: per-tile router thresholds, per-tile streams, i.e.
exactly B independent B == 1 calls), and the tile groups of an image can be spread over ranks
(`dist.shard_range`).  bpp of an image = sum of tile bits / (H * W) as in the reference (:250,256).
The CNN encoder / decoder are out of scope (stock PyTorch) and are passed in by the caller.
"""
from __future__ import annotations

import argparse
import os
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from numpy import exp, pi, sqrt  # numpy scalars, as the reference (inference_high_resolution.py:21): same last bits

from . import ops
from .codec import BinaryCoding, HuffmanCoding

TILE = 768                       # inference_high_resolution.py:114-123


# ------------------------------------------------------------------------------------ CLI / config
def get_parser(high_resolution: bool = False, **parser_kwargs) -> argparse.ArgumentParser:
    """inference.py:112-124; with high_resolution=True the defaults of inference_high_resolution.py:176-194."""
    parser = argparse.ArgumentParser(**parser_kwargs)
    parser.add_argument("-i", "--images_dir", type=str, default="" if high_resolution else "../dataset/Kodak", required=False,
                        help="Path to the root directory where the images are")
    parser.add_argument("-b", "--batch_size", type=int, default=1, help="Number of images in a minibatch")
    parser.add_argument("-n", "--num_workers", type=int, default=1, help="Number of worker threads to load the images")
    parser.add_argument("-s", "--image_size", type=int, default=512, help="Size of the reconstructed image")
    parser.add_argument("-o", "--output_dir", type=str, default="output_reconstruction" if high_resolution else "./output",
                        help="Path to a directory where the outputs will be saved")
    if high_resolution:
        parser.add_argument("-w", "--write_partiton_map", default=False,
                            help="If set, the partition maps will also be saved to the output directory")
    else:
        parser.add_argument("-w", "--write_partiton_map", action="store_true",
                            help="If set, the partition maps will be saved to the output directory")
    parser.add_argument("-r", "--images_range", type=int, nargs=2, default=(0, -1),
                        help="Optional. Two values: starting and ending indices of the images to be loaded (manual sharding).")
    return parser


class _Cfg(dict):
    """dict with attribute access, enough of OmegaConf for `config.model.params`."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return _Cfg(v) if isinstance(v, dict) else v


def load_config(config_path, display: bool = False):
    """inference.py:80-84 (the YAML keys are unchanged; `target:` / `params:` as in CGIC/util.py:18-28)."""
    import yaml
    with open(config_path) as f:
        cfg = yaml.safe_load(f)
    if display:
        print(yaml.dump(cfg))
    return _Cfg(cfg)


def load_model(config, ckpt_path=None, encoder=None, decoder=None):
    """inference.py:87-93.  The CNNs are out of scope: they are built from `ddconfig` with the reference's own classes when
    that package is importable (CGIC.__init__ / model.reference_cnns), or passed in as `encoder` / `decoder`."""
    from .model import CGIC
    params = dict(config["model"]["params"])
    params.pop("ckpt_path", None)
    params.pop("lossconfig", None)                       # training only
    model = CGIC(**params, encoder=encoder, decoder=decoder)
    if ckpt_path is not None:
        sd = torch.load(ckpt_path, map_location="cpu")["state_dict"]
        model.load_state_dict(sd, strict=False)
        print(f"Restored from {ckpt_path}")
    return model.eval()


# ------------------------------------------------------------------------------------ data in / images out
class ImageDataset(torch.utils.data.Dataset):
    """inference.py:34-79: every .jpg / .jpeg / .png below a directory (sorted), optionally the slice `images_range`,
    centre-cropped to multiples of 16 and turned into a [3,H,W] float tensor in [0, 1] (`ToTensor`)."""

    def __init__(self, imagenet_images_dir, target_size: int = 512, images_range: Tuple[int, int] = (0, -1)) -> None:
        super().__init__()
        from pathlib import Path
        self.target_size = target_size
        root = Path(imagenet_images_dir)
        self.image_paths = sorted(p for p in root.glob("**/*") if self._is_image_path(p))
        if images_range[1] > 0:
            self.image_paths = self.image_paths[images_range[0]:images_range[1]]
        print(f"Found {len(self.image_paths)} images to reconstruct")

    def __getitem__(self, index: int) -> torch.Tensor:
        from PIL import Image
        image = self._resize_and_crop(Image.open(self.image_paths[index]))
        a = np.asarray(image)
        if a.ndim == 2:
            a = a[:, :, None]
        # torchvision's ToTensor: HWC uint8 -> CHW float32 / 255
        return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))).to(torch.float32).div(255)

    @staticmethod
    def _resize_and_crop(img):
        """centre crop to (16 * (h // 16), 16 * (w // 16)) with torchvision's rounding of the offsets (inference.py:62-67)"""
        w, h = img.size
        ch, cw = 16 * (h // 16), 16 * (w // 16)
        top, left = int(round((h - ch) / 2.0)), int(round((w - cw) / 2.0))
        return img.crop((left, top, left + cw, top + ch))

    def __len__(self) -> int:
        return len(self.image_paths)

    @staticmethod
    def _is_image_path(path) -> bool:
        ext = path.name[path.name.rfind(".") + 1:]
        return not (path.is_dir() or path.name.startswith(".") or len(ext) == 0 or ext.lower() not in ("jpg", "jpeg", "png"))


def write_images(images: torch.Tensor, output_dir, i: int, batch_size: int, start_offset: int, bpp=None) -> None:
    """inference.py:95-110: `{k:03d}_{bpp:05f}.png` (or `{k:03d}.png`), k = i * batch_size + j + start_offset."""
    from pathlib import Path
    from PIL import Image
    images = (255 * images.permute(0, 2, 3, 1).detach().cpu().numpy()).astype(np.uint8)
    for j, img in enumerate(images):
        k = i * batch_size + j + start_offset
        name = f"{k:03d}_{bpp:05f}.png" if bpp is not None else f"{k:03d}.png"
        Image.fromarray(img).save(Path(output_dir) / name)


# ------------------------------------------------------------------------------------ tiling geometry
def compute_padding(in_h: int, in_w: int, *, out_h=None, out_w=None, min_div: int = 1):
    """inference_high_resolution.py:143-173: centred zero padding up to a multiple of min_div."""
    if out_h is None:
        out_h = (in_h + min_div - 1) // min_div * min_div
    if out_w is None:
        out_w = (in_w + min_div - 1) // min_div * min_div
    if out_h % min_div != 0 or out_w % min_div != 0:
        raise ValueError(f"Padded output height and width are not divisible by min_div={min_div}.")
    left = (out_w - in_w) // 2
    right = out_w - in_w - left
    top = (out_h - in_h) // 2
    bottom = out_h - in_h - top
    return (left, right, top, bottom), (-left, -right, -top, -bottom)


def nonoverlapping_grid_indices(x_padded) -> Tuple[List[int], List[int], List[int], List[int]]:
    """inference_high_resolution.py:112-125 (accepts a tensor or a shape): tile origins and sizes per axis."""
    shape = x_padded.shape if hasattr(x_padded, "shape") else x_padded
    h, w = int(shape[-2]), int(shape[-1])
    h_list = list(range(0, h, TILE))
    w_list = list(range(0, w, TILE))
    tile_h_list = [TILE] * (h // TILE) + ([h % TILE] if h % TILE else [])
    tile_w_list = [TILE] * (w // TILE) + ([w % TILE] if w % TILE else [])
    return h_list, w_list, tile_h_list, tile_w_list


def gaussian_weights(tile_width: int, tile_height: int, nbatches: int, device) -> torch.Tensor:
    """inference_high_resolution.py:127-141 (float64 [nbatches,3,th,tw]); with non-overlapping tiles the
    weights cancel in x_rec / contributors (:253), they are kept so that the arithmetic is the reference's."""
    var = 0.01
    midpoint = (tile_width - 1) / 2
    x_probs = [exp(-(x - midpoint) * (x - midpoint) / (tile_width * tile_width) / (2 * var)) / sqrt(2 * pi * var) for x in range(tile_width)]
    midpoint = tile_height / 2
    y_probs = [exp(-(y - midpoint) * (y - midpoint) / (tile_height * tile_height) / (2 * var)) / sqrt(2 * pi * var) for y in range(tile_height)]
    return torch.tile(torch.tensor(np.outer(y_probs, x_probs), device=device), (nbatches, 3, 1, 1))


def tile_plan(H: int, W: int) -> List[Tuple[int, int, int, int]]:
    """(y, x, tile_h, tile_w) of every tile of a padded H x W image in the reference's visiting order."""
    h_list, w_list, th, tw = nonoverlapping_grid_indices((1, 3, H, W))
    return [(h_list[i], w_list[j], th[i], tw[j]) for i in range(len(h_list)) for j in range(len(w_list))]


def group_tiles(plan: Sequence[Tuple[int, int, int, int]]) -> Dict[Tuple[int, int], List[int]]:
    """tile indices by (tile_h, tile_w): every group is one batched launch of the hot path."""
    groups: Dict[Tuple[int, int], List[int]] = {}
    for i, (_, _, th, tw) in enumerate(plan):
        groups.setdefault((th, tw), []).append(i)
    return groups


# ------------------------------------------------------------------------------------ drivers
@torch.no_grad()
def compress_tiled(model, x: torch.Tensor, h_indices: HuffmanCoding, h_mask: BinaryCoding = None, output_dir=None,
                   rank: int = 0, world: int = 1):
    """inference_high_resolution.py:226-257 for ONE image x [1,3,H,W] in [0,1]: pad to a multiple of 16,
    cut into 768-pixel tiles, compress every tile independently (equal-shape tiles in one launch),
    blend, un-pad.  Returns (x_rec [1,3,H,W], bpp_image, tiles) with tiles = list of dicts
    (y, x, h, w, bpp, sizes [5]) in the reference's visiting order.  With world > 1 this rank handles
    its contiguous share of the tile groups' tiles and the caller reduces bit_sum / x_rec.
    When output_dir is given the five .bin files of the LAST tile are left there, as the
    reference does by overwriting them tile after tile (:246)."""
    assert x.dim() == 4 and x.shape[0] == 1
    H0, W0 = x.shape[-2:]
    pad, unpad = compute_padding(H0, W0, min_div=2 ** 4)
    x_padded = F.pad(x, pad, mode="constant", value=0)
    plan = tile_plan(*x_padded.shape[-2:])
    x_rec = torch.zeros(x_padded.shape, device=x.device)               # fp32 accumulators, as the reference (:230-231)
    contributors = torch.zeros(x_padded.shape, device=x.device)
    tiles: List[dict] = [None] * len(plan)
    bit_sum = 0.0
    from .dist import shard_range
    for (th, tw), members in group_tiles(plan).items():
        lo, hi = shard_range(len(members), rank, world)
        mine = members[lo:hi]
        if not mine:
            continue
        batch = torch.cat([x_padded[:, :, plan[i][0]:plan[i][0] + th, plan[i][1]:plan[i][1] + tw] for i in mine], 0)
        out = model.compress_batch(batch, h_indices, per_image=True)
        wts = gaussian_weights(tw, th, 1, x.device)
        for k, i in enumerate(mine):
            y0, x0 = plan[i][0], plan[i][1]
            x_rec[:, :, y0:y0 + th, x0:x0 + tw] += out["dec"][k:k + 1] * wts           # fp64 product rounded into fp32 (:248)
            contributors[:, :, y0:y0 + th, x0:x0 + tw] += wts
            bpp = out["bpp"][k]
            bit_sum += bpp * tw * th                                                  # :250
            tiles[i] = dict(y=y0, x=x0, h=th, w=tw, bpp=bpp, sizes=out["sizes_host"][k].tolist(), mode=out["mode"])
        if output_dir is not None and mine[-1] == len(plan) - 1:
            _write_streams(output_dir, out, len(mine) - 1, h_indices)
    if world == 1:
        x_rec /= contributors                                                              # :253
    else:
        x_rec = torch.where(contributors > 0, x_rec / contributors, x_rec)               # other ranks' tiles stay zero
    x_rec = F.pad(x_rec.clamp(0, 1), unpad)
    return x_rec, bit_sum / W0 / H0, tiles                                            # :256


def _write_streams(path, out, k: int, h_indices: HuffmanCoding) -> None:
    offs, _, _ = h_indices.table.layout(*out["grid"])
    blob = out["bytes"][k].cpu().numpy()
    sizes = out["sizes_host"][k].tolist()
    for s, name in enumerate(ops.STREAM_NAMES):
        if ops.stream_present(out["mode"], s):
            with open(os.path.join(path, name + ".bin"), "wb") as f:
                f.write(blob[offs[s]: offs[s] + sizes[s]].tobytes())


def tiled_hot_path(z_groups: Sequence[torch.Tensor], mask_groups: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]], mode: int,
                   table: ops.HuffTable, codebook: torch.Tensor, tile_pixels: Sequence[Tuple[int, int]], image_pixels: int,
                   prepared: "ops.Codebook" = None):
    """The hot path of BASELINE config 5 without the CNNs.  One entry per tile GROUP (equal-shape tiles):
    z_groups[g] = [n,4,h,w] latents of the group's n tiles, mask_groups[g] their router masks, tile_pixels[g]
    = (tile_h, tile_w).  Each group is one launch of VQ, pack and unpack.  Returns (bpp_image, [per-group dict])
    with bpp_image = sum of tile bits / image_pixels (inference_high_resolution.py:250,256)."""
    bit_sum = 0.0
    outs = []
    for z, (mc, mm, mf), (th, tw) in zip(z_groups, mask_groups, tile_pixels):
        h, w = z.shape[-2:]
        if prepared is not None:
            idx, zq, sq, packed, sizes = ops.encode(z, prepared, mc, mm, mf, mode, table)
        else:
            idx, zq, sq = ops.vq_assign(z, codebook)
            packed, sizes = ops.pack(idx, mc, mm, mf, mode, table, h, w)
        dmc, dmm, dmf, ind, quant, status = ops.unpack(packed, sizes, mode, table, codebook, h, w)
        sz = sizes.cpu()
        bpp = [int(sz[b].sum()) * 8 / (th * tw) for b in range(z.shape[0])]
        bit_sum += sum(b * th * tw for b in bpp)
        outs.append(dict(idx=idx, zq=zq, sqerr=sq, bytes=packed, sizes=sz, bpp=bpp, ind=ind, quant=quant, masks=(dmc, dmm, dmf), status=status))
    return bit_sum / image_pixels, outs


def run(model, dataloader, output_dir, h_indices: HuffmanCoding, h_mask: BinaryCoding, high_resolution: bool = False,
        write_image=None, n_images=None):
    """The loop of inference.py:156-171 / inference_high_resolution.py:220-262: per image compress (tiled when
    high_resolution), append `image: i \\t bpp: b` to bpp.txt, finally `Bpp Average: ...`."""
    os.makedirs(output_dir, exist_ok=True)
    bpp_sum, n = 0.0, 0
    with open(os.path.join(output_dir, "bpp.txt"), "a") as f:
        for i, x in enumerate(dataloader):
            x = x.cuda()
            with torch.no_grad():
                if high_resolution:
                    x_rec, bpp, _ = compress_tiled(model, x, h_indices, h_mask, output_dir)
                else:
                    x_rec, bpp, _ = model.compress(x, output_dir, h_indices, h_mask, False)
                    x_rec = x_rec.clamp(0, 1)
            if write_image is not None:
                write_image(x_rec, i, bpp)
            bpp_sum += bpp
            n += 1
            f.write(f"image: {i} \t bpp: {bpp}\n")
        total = n_images if n_images is not None else n
        f.write(f"Bpp Average: {bpp_sum / max(total, 1)}")
    print(f"Bpp Average: {bpp_sum / max(total, 1)}")
    return bpp_sum / max(total, 1)


def main(argv=None, high_resolution: bool = False, config_path: str = "./configs/config_inference.yaml", encoder=None, decoder=None):
    """inference.py:127-171 / inference_high_resolution.py:197-262: parse the reference's flags, load the images, build the
    model from the reference's YAML, `HuffmanCoding(model.quantize.embedding_counter)`, compress every image (tiled at 768
    pixels when high_resolution), write `reconstructed/{k:03d}_{bpp:05f}.png` and `bpp.txt`.  Returns the average bpp."""
    from pathlib import Path
    from torch.utils.data import DataLoader
    opt, _ = get_parser(high_resolution).parse_known_args(argv)
    dataset = ImageDataset(opt.images_dir, opt.image_size, tuple(opt.images_range))
    dataloader = DataLoader(dataset, opt.batch_size, num_workers=opt.num_workers)
    config = load_config(config_path, display=True)
    ckpt = config["model"]["params"].get("ckpt_path")
    model = load_model(config, ckpt if ckpt and os.path.exists(ckpt) else None, encoder=encoder, decoder=decoder).to("cuda")
    h_string = HuffmanCoding(model.quantize.embedding_counter)
    h_mask = BinaryCoding()
    print("number of params (M): %.2f" % (sum(p.numel() for p in model.parameters() if p.requires_grad) / 1.e6))
    Path(opt.output_dir).mkdir(parents=True, exist_ok=True)
    rec_output_dir = Path(opt.output_dir) / "reconstructed"
    rec_output_dir.mkdir(parents=True, exist_ok=True)
    start = opt.images_range[0]
    return run(model, dataloader, opt.output_dir, h_string, h_mask, high_resolution=high_resolution,
               write_image=lambda x_rec, i, bpp: write_images(x_rec, rec_output_dir, i, opt.batch_size, start, bpp=bpp), n_images=len(dataset))


if __name__ == "__main__":
    import sys
    main(high_resolution="--high-resolution" in sys.argv)
