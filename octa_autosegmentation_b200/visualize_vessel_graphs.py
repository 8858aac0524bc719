"""Drop-in for the reference's visualize_vessel_graphs.py: re-renders stored graph CSVs at any resolution on the GPU, in
batches (one voxelize / raster launch per batch of files, file writers in a thread pool).  Same flags and output names
(visualize_vessel_graphs.py:33-46,69-104):

    python -m octa_autosegmentation_b200.visualize_vessel_graphs --source_dir D --out_dir O \
        [--resolution 1216,1216,16] [--save_2d|--no_save_2d] [--save_3d] [--save_3d_as .nii.gz|.npy] [--mip_axis 2]
        [--binarize] [--num_samples N] [--max_dropout_prob P] [--ignore_z] [--threads T] [--batch B]

  <name>.png | <name>_label.png (1-bit, PIL Floyd-Steinberg `convert("1")` after img<0.1 -> 0)
  <name>_3d[.nii.gz|.npy] | <name>_3d_label[...]      <name>_3d[_label]_blackdict.pkl when --max_dropout_prob > 0
Mirrored quirks of the reference: `.npy` volumes are written as bool (:94); the 2-D image is rendered WITHOUT dropout (:95);
with --save_3d the 2-D file names carry the 3-D suffix, because `name` is extended in place (:81-85,:99-101); the blackdict that
is pickled next to the 2-D image is the 3-D one (:102-104; without --save_3d that line raises NameError in the reference's
worker, which its pool swallows: no pickle is written).  NIfTI files are written by graph_io.save_nifti (no nibabel needed)."""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import csv
import os
import pickle
import re
import sys
from glob import glob
from random import random

import numpy as np

from . import graph_io, tree2img


def natural_key(s: str):
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", s)]


def _write_outputs(args, name, vol, black_dict, img):
    """File writes of render_graph (visualize_vessel_graphs.py:79-104) for one sample; vol / img may be None."""
    from PIL import Image

    if vol is not None:
        if args.binarize:
            name += "_3d_label"
            vol[vol < 0.1] = 0
            vol[vol >= 0.1] = 1
        else:
            name += "_3d"
        if args.save_3d_as == ".nii.gz":
            graph_io.save_nifti(os.path.join(args.out_dir, name + ".nii.gz"), vol)      # nib.Nifti1Image(vol, np.eye(4)), :85-87
        else:
            np.save(os.path.join(args.out_dir, name + ".npy"), vol.astype(np.bool_))
        if args.max_dropout_prob > 0:
            with open(os.path.join(args.out_dir, name + "_blackdict.pkl"), "wb") as f:
                pickle.dump(black_dict, f)
    if img is not None:
        if args.binarize:
            img[img < 0.1] = 0
            graph_io.save_png(os.path.join(args.out_dir, name + "_label.png"), np.array(Image.fromarray(img.astype(np.uint8)).convert("1")))
        else:
            graph_io.save_png(os.path.join(args.out_dir, name + ".png"), img.astype(np.uint8))
        if args.max_dropout_prob > 0 and vol is not None:
            with open(os.path.join(args.out_dir, name + "_blackdict.pkl"), "wb") as f:
                pickle.dump(black_dict, f)


_READERS = None


def _readers():
    global _READERS
    if _READERS is None:
        ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)
        _READERS = cf.ThreadPoolExecutor(max_workers=max(1, min(8, ncores)))
    return _READERS


def render_batch(files, args, resolution, img_res, writers):
    """One batch of csv files: parse, (3-D) host-side dropout exactly like voxelize_forest, ONE voxelize launch and ONE raster
    launch for the whole batch, file writes handed to `writers`."""
    import torch

    names, e3, e2, bds = [], [], [], []
    # the C parser releases the GIL (7 ms per 13 k-row file): the files of the batch are read side by side
    dropout3d = args.save_3d and args.max_dropout_prob > 0
    parsed = dict(zip(files, _readers().map(graph_io.read_csv, files))) if (args.save_2d or not dropout3d) else {}
    for fp in files:
        names.append(fp.split("/")[-1].removesuffix(".csv"))
        if args.save_3d and args.max_dropout_prob > 0:
            # subtree dropout consumes Python's RNG row by row (tree2img.py:218-241): same host loop as voxelize_forest
            with open(fp, newline="") as f:
                rows = list(csv.DictReader(f))
            kept, bd = tree2img.forest_to_edges7(rows, None, 0, 1, args.max_dropout_prob, None)
            e3.append(kept)
            bds.append(bd)
            if args.save_2d:
                # rasterize_forest(f, img_res, mip_axis) draws p, and one number per edge even though p = 0 (tree2img.py:62,78):
                # the next file's dropout continues from there
                for _ in range(1 + len(rows)):
                    random()
                e2.append(parsed[fp])
            else:
                e2.append(None)
        else:
            e = parsed[fp]                       # C parser of the `[x y z]` cells (tree2img.py:73-76 semantics)
            e3.append(e)
            e2.append(e)
            bds.append({})
    dev = torch.device("cuda", torch.cuda.current_device())
    vols = imgs = None
    if args.save_3d:
        offs = np.cumsum([0] + [len(e) for e in e3])
        cat = np.concatenate(e3) if offs[-1] else np.zeros((1, 7))
        vols = tree2img.voxelize_batch_device(torch.from_numpy(cat).to(dev), offs, [int(d) for d in resolution], ignore_z=args.ignore_z).cpu().numpy()
    if args.save_2d:
        offs = np.cumsum([0] + [len(e) for e in e2])
        cat = np.concatenate(e2) if offs[-1] else np.zeros((1, 7))
        imgs = tree2img.raster_batch_device(torch.from_numpy(cat).to(dev), offs, [int(d) for d in img_res], args.mip_axis).cpu().numpy()
    futs = []
    for i, name in enumerate(names):
        vol = vols[i] if vols is not None else None
        img = imgs[i].astype(np.uint16) if imgs is not None else None
        futs.append(writers.submit(_write_outputs, args, name, vol, bds[i], img))
    return futs


def main(argv=None):
    p = argparse.ArgumentParser(description="")
    p.add_argument("--source_dir", type=str, required=True)
    p.add_argument("--out_dir", type=str, required=True)
    p.add_argument("--resolution", type=str, default="1216,1216,16")
    p.add_argument("--save_2d", action="store_true")
    p.add_argument("--no_save_2d", action="store_false", dest="save_2d")
    p.add_argument("--save_3d", action="store_true")
    p.add_argument("--save_3d_as", choices=[".nii.gz", ".npy"], default=".nii.gz")
    p.add_argument("--mip_axis", type=int, default=2)
    p.add_argument("--binarize", action="store_true")
    p.add_argument("--num_samples", type=int, default=9999999)
    p.add_argument("--max_dropout_prob", type=float, default=0)
    p.add_argument("--ignore_z", action="store_true", default=False)
    p.add_argument("--threads", type=int, default=-1, help="file-writer threads (default: all cores but one, max 8)")
    p.add_argument("--batch", type=int, default=0, help="csv files per GPU launch (default: 32 for 2-D only, 4 with --save_3d)")
    p.set_defaults(save_2d=True)
    args = p.parse_args(argv)
    resolution = np.array([int(d) for d in args.resolution.split(",")])
    assert not args.save_3d or len(resolution) == 3, "If you want to generate the 3d volume, you need to specify the resolution of all three dimensions."
    assert os.path.isdir(args.source_dir), f"The provided source directory {args.source_dir} does not exist."
    assert args.mip_axis in [0, 1, 2], "The axis must be '0' (x), '1' (y) or '2' (z)."
    assert args.save_3d or args.save_2d, "You must either activate saving the 2D image or the 3D volume."
    os.makedirs(args.out_dir, exist_ok=True)
    img_res = None
    if args.save_2d:
        img_res = [*resolution]
        if len(resolution) == 3:
            del img_res[args.mip_axis]
    files = sorted(glob(os.path.join(args.source_dir, "**", "*.csv"), recursive=True), key=natural_key)[:args.num_samples]
    assert len(files) > 0, f"Your provided source directory {args.source_dir} does not contain any csv files."
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    mine = [fp for i, fp in enumerate(files) if i % world == rank]
    batch = args.batch if args.batch > 0 else (4 if args.save_3d else 32)
    threads = args.threads if args.threads > 0 else max(1, min(8, (os.cpu_count() or 2) - 1))
    futs = []
    with cf.ThreadPoolExecutor(max_workers=threads) as writers:
        for k in range(0, len(mine), batch):
            futs += render_batch(mine[k:k + batch], args, resolution, img_res, writers)
        for f in futs:
            f.result()
    return 0


if __name__ == "__main__":
    sys.exit(main())
