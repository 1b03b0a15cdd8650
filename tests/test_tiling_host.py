"""Host logic of the tiling driver / CLI mirror (cgic_b200.inference) against vectors generated from the
unmodified reference (tests/golden/make_tiling_golden.py).  No GPU needed: geometry, weights, parser."""
import hashlib
import json
import os

import numpy as np
import torch

from conftest import GOLDEN


def _kats():
    with open(os.path.join(GOLDEN, "tiling_kats.json")) as f:
        return json.load(f)


def test_tiling_geometry_matches_reference():
    import cgic_b200 as cg
    inf = cg.inference
    for s in _kats()["shapes"]:
        pad, unpad = inf.compute_padding(s["H"], s["W"], min_div=16)
        assert list(pad) == s["pad"] and list(unpad) == s["unpad"], s
        Hp, Wp = s["H"] + pad[2] + pad[3], s["W"] + pad[0] + pad[1]
        h_list, w_list, th, tw = inf.nonoverlapping_grid_indices(torch.zeros(1, 3, Hp, Wp))
        assert (h_list, w_list, th, tw) == (s["h_list"], s["w_list"], s["tile_h"], s["tile_w"]), s
        plan = inf.tile_plan(Hp, Wp)
        assert len(plan) == len(h_list) * len(w_list) and sum(t[2] * t[3] for t in plan) == Hp * Wp
        groups = inf.group_tiles(plan)
        assert sorted(i for g in groups.values() for i in g) == list(range(len(plan)))


def test_div2k_shape_is_the_six_tiles_of_the_survey():
    import cgic_b200 as cg
    plan = cg.inference.tile_plan(1344, 2032)          # DIV2K 2040x1356 after the dataset's x16 centre crop
    assert [(t[2], t[3]) for t in plan] == [(768, 768), (768, 768), (768, 496), (576, 768), (576, 768), (576, 496)]


def test_gaussian_weights_match_reference():
    import cgic_b200 as cg
    for w in _kats()["weights"]:
        got = cg.inference.gaussian_weights(w["tile_w"], w["tile_h"], 1, "cpu")
        assert list(got.shape) == w["shape"] and str(got.dtype) == w["dtype"]
        assert hashlib.sha256(np.ascontiguousarray(got.numpy()).tobytes()).hexdigest() == w["sha256"], w


def test_parsers_and_config_keys():
    import cgic_b200 as cg
    k = _kats()
    for hr, want in ((False, k["parser"]), (True, k["parser_hr"])):
        got = {a: (list(v) if isinstance(v, tuple) else v) for a, v in vars(cg.inference.get_parser(high_resolution=hr).parse_args([])).items()}
        assert got == want
    args = cg.inference.get_parser().parse_args(["-i", "d", "-b", "2", "-n", "3", "-s", "256", "-o", "o", "-w", "-r", "4", "9"])
    assert (args.images_dir, args.batch_size, args.num_workers, args.image_size, args.output_dir, args.write_partiton_map,
            list(args.images_range)) == ("d", 2, 3, 256, "o", True, [4, 9])
    cfg = cg.inference._Cfg({"model": {"params": {"ddconfig": {"router_config": {"params": {"coarse_grain_ratio": 0.1}}}}}})
    assert cfg.model.params.ddconfig.router_config.params.coarse_grain_ratio == 0.1


def test_image_dataset_and_writer(tmp_path):
    """inference.py:34-79, 95-110: image discovery, centre crop to multiples of 16 (== torchvision's center_crop + ToTensor),
    `images_range`, and the `{k:03d}_{bpp:05f}.png` naming."""
    import numpy as np
    import torch
    import torchvision.transforms as T
    import torchvision.transforms.functional as TF
    from PIL import Image
    import cgic_b200 as cg
    rng = np.random.default_rng(0)
    for name, (h, w) in {"a.png": (70, 53), "b.jpg": (64, 64), "sub/c.PNG": (33, 47), ".hidden.png": (32, 32)}.items():
        path = tmp_path / name
        path.parent.mkdir(exist_ok=True)
        Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8)).save(path)
    (tmp_path / "notes.txt").write_text("x")
    ds = cg.inference.ImageDataset(tmp_path)
    assert [p.name for p in ds.image_paths] == ["a.png", "b.jpg", "c.PNG"]
    for i, p in enumerate(ds.image_paths):
        img = Image.open(p)
        w, h = img.size
        assert torch.equal(ds[i], T.ToTensor()(TF.center_crop(img, [16 * (h // 16), 16 * (w // 16)])))
    assert len(cg.inference.ImageDataset(tmp_path, images_range=(1, 3))) == 2
    out = tmp_path / "out"
    out.mkdir()
    cg.inference.write_images(torch.rand(2, 3, 16, 16), out, 3, 2, 10, bpp=0.22961)
    cg.inference.write_images(torch.rand(1, 3, 16, 16), out, 0, 1, 0)
    assert sorted(p.name for p in out.iterdir()) == ["000.png", "016_0.229610.png", "017_0.229610.png"]


def test_config5_tile_sharding_covers_every_tile_once():
    """bench.py deals the tiles of a 2032 x 1344 image round-robin over the ranks (tile i -> rank i % N); compress_tiled
    shards every equal-shape group contiguously.  Either way every tile is handled exactly once for N = 1, 2, 4, 8."""
    import cgic_b200 as cg
    from cgic_b200 import dist as cdist
    plan = cg.inference.tile_plan(2032, 1344)
    assert len(plan) == 6 and sorted({(p[2], p[3]) for p in plan}) == [(496, 576), (496, 768), (768, 576), (768, 768)]
    assert sum(p[2] * p[3] for p in plan) == 2032 * 1344
    groups = cg.inference.group_tiles(plan)
    for world in (1, 2, 4, 8):
        rr = [i for r in range(world) for i in range(len(plan)) if i % world == r]
        assert sorted(rr) == list(range(len(plan)))
        seen = []
        for members in groups.values():
            for r in range(world):
                lo, hi = cdist.shard_range(len(members), r, world)
                seen += members[lo:hi]
        assert sorted(seen) == list(range(len(plan)))
