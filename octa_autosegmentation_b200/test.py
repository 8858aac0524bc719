"""Drop-in for the reference's test.py restricted to GAN contrast adaptation (`General.inference: G`): the
second stage of docker/dockershell.sh:13-19 (`generation N`: generate graphs -> test.py -> visualize).

    python -m octa_autosegmentation_b200.test --config_file docker/trained_models/GAN/config.yml \
        [--num_samples N] [--batch_size B] [--Test.data.real_A.files '/var/generation/vessel_graphs/**/*.csv'] \
        [--Test.save_dir /var/generation/images] [--Test.model_path .../150_G_model.pth] [--General.seed S]

Same config keys as the reference (test.py:33-50, data/image_dataset.py:41-81): `Test.data.real_A.files` (graph CSVs,
natural-sorted), `Test.data.background.files` (background PNGs, one drawn per sample), `Test.model_path`,
`Test.save_dir` (default `Output.save_dir/test`), `General.inference` (file prefix), `Test.data_augmentation` (only the
stanza of the shipped config is supported: LoadGraphAndFilterByRandomRadiusd at one resolution, ScaleIntensityd,
Rotate90d(k=1) + Flipd(0) on the background, AddRandomBackgroundNoised).  Output: `<save_dir>/<inference>_<csv name>.png`
= uint8(pred * 255) (utils/visualizer.py:330-338).  Every stage after the CSV parse runs on the GPU in batches:
2-D raster, input transform, speckle stream, generator.

Randomness: the reference draws the background index with `random.randint` inside dataloader worker processes and the
speckle with the workers' `np.random`, so its outputs are not reproducible run to run.  Here sample i (index in the sorted
CSV list) uses `random.Random(seed + i).randint(0, n_bg - 1)` and the stream `np.random.seed(seed + i)` with seed =
`General.seed`.  Under `torchrun --nproc-per-node N` rank r renders samples i = r (mod N) on GPU `LOCAL_RANK`.
"""
from __future__ import annotations

import argparse
import csv
import os
import random
import sys
from glob import glob

import numpy as np

from . import config as cfgmod
from . import gan, graph_io, tree2img
from .visualize_vessel_graphs import natural_key


def _resolution(config: dict):
    for aug in config["Test"].get("data_augmentation", []):
        if aug.get("name") == "LoadGraphAndFilterByRandomRadiusd":
            res = aug.get("image_resolutions", [[304, 304]])[0]
            return int(res[0]), int(res[1]), float((aug.get("min_radius") or [0])[0])
    return 304, 304, 0.0


def run(config: dict, num_samples: int = 9999999, batch_size: int = 32, device=None) -> list:
    import torch
    from PIL import Image

    # one process per GPU (torchrun): rank r renders samples i = r (mod world) and writes its own files; a sample's output
    # depends on its global index only, so any world size writes the same set of files
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    device = torch.device(device or ("cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else "cuda"))
    test = config["Test"]
    seed = int(config["General"].get("seed") or 0)
    csvs = sorted(glob(test["data"]["real_A"]["files"], recursive=True), key=natural_key)
    assert len(csvs) > 0, f"Error: Your provided file path {test['data']['real_A']['files']} for real_A does not match any files!"
    csvs = csvs[:num_samples]
    mine = list(range(rank, len(csvs), world))
    bgs = []
    if "background" not in test["data"]:
        # data_transforms.py:508 falls back to torch.rand_like(img) -- an unseeded draw from torch's global generator
        raise NotImplementedError("Test.data.background is required: the reference's fallback (torch.rand_like noise) is not reproducible")
    if "background" in test["data"]:
        bgs = sorted(glob(test["data"]["background"]["files"], recursive=True), key=natural_key)
        assert len(bgs) > 0, f"Error: Your provided file path {test['data']['background']['files']} for background does not match any files!"
    W, H, min_radius = _resolution(config)
    save_dir = test.get("save_dir") or config["Output"]["save_dir"] + "/test"
    os.makedirs(save_dir, exist_ok=True)
    prefix = (config["General"].get("inference") or "pred") + "_"
    G = gan.ResnetGenerator9.from_checkpoint(test["model_path"], image_size=(H, W), max_images=batch_size, device=device)
    written = []
    with torch.cuda.device(device):
        for b0 in range(0, len(mine), batch_size):
            gidx = mine[b0:b0 + batch_size]
            paths = [csvs[i] for i in gidx]
            edges = []
            for p in paths:
                with open(p, "rb") as f:
                    edges.append(graph_io.parse_csv_bytes(f.read()))
            offs = np.zeros(len(edges) + 1, dtype=np.int64)
            offs[1:] = np.cumsum([e.shape[0] for e in edges])
            e7 = torch.from_numpy(np.concatenate(edges, axis=0) if offs[-1] else np.zeros((1, 7))).to(device)
            raster = tree2img.raster_batch_device(e7, offs, (W, H), 2, min_radius=min_radius)
            bg_t, seeds = None, None
            if bgs:
                idx = [random.Random(seed + i).randint(0, len(bgs) - 1) for i in gidx]
                bg = np.stack([load_background(bgs[j], W, H) for j in idx])
                bg_t = torch.from_numpy(bg).to(device)
                seeds = [(seed + i) & 0xFFFFFFFF for i in gidx]
            out = gan.contrast_adapt(G, raster, bg_t, seeds).cpu().numpy()
            gan.save_images(save_dir, paths, out, prefix=prefix)
            written += [os.path.join(save_dir, prefix + ".".join(os.path.basename(p).split(".")[:-1]) + ".png") for p in paths]
    G.close()
    return written


def load_background(path: str, W: int, H: int) -> np.ndarray:
    """What MONAI's LoadImaged hands on for a PNG (docker/trained_models/GAN/config.yml:49-55): PILReader's default
    reverse_indexing=True swaps the two spatial axes, i.e. the array arrives TRANSPOSED; the config's Rotate90d(k=1) + Flipd(0)
    (a transpose, applied on the device by octa_gan_input_dev) then restores the image's own orientation."""
    from PIL import Image
    im = Image.open(path).convert("L")
    if im.size != (W, H):
        im = im.resize((W, H))
    return np.ascontiguousarray(np.asarray(im, dtype=np.uint8).T)


def main(argv=None):
    parser = argparse.ArgumentParser(description="")
    parser.add_argument("--config_file", type=str, required=True)
    parser.add_argument("--epoch", type=str, default="best")
    parser.add_argument("--num_samples", type=int, default=9999999)
    parser.add_argument("--num_workers", type=int, default=None, help="accepted for compatibility; loading is batched on the GPU")
    parser.add_argument("--batch_size", type=int, default=32, help="images per GPU batch")
    args, unknown = parser.parse_known_args(argv)
    assert args.num_samples > 0
    path = os.path.abspath(args.config_file)
    assert os.path.isfile(path), f"Your provided config path {args.config_file} does not exist!"
    config = cfgmod.read_config(path)
    cfgmod.apply_cli_overrides_from_unknown_args(config, unknown)
    if (config["General"].get("inference") or "G") != "G" or config["General"].get("task") not in (None, "gan-ves-seg"):
        raise NotImplementedError("only GAN contrast adaptation (General.inference: G) runs on this path")
    files = run(config, args.num_samples, args.batch_size)
    print("wrote %d images to %s" % (len(files), os.path.dirname(files[0]) if files else "-"))


if __name__ == "__main__":
    main(sys.argv[1:])
