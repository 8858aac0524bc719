"""Drop-in for the reference's visualize_vessel_graphs.py: re-renders stored graph CSVs at any resolution on
the GPU.  Same flags and output names (visualize_vessel_graphs.py:33-46,80-104):

    python -m octa_autosegmentation_b200.visualize_vessel_graphs --source_dir D --out_dir O \
        [--resolution 1216,1216,16] [--save_2d|--no_save_2d] [--save_3d] [--save_3d_as .nii.gz|.npy] [--mip_axis 2]
        [--binarize] [--num_samples N] [--max_dropout_prob P] [--ignore_z] [--threads T]

  <name>.png | <name>_label.png (1-bit, PIL Floyd-Steinberg `convert("1")` after img<0.1 -> 0)
  <name>_3d[.nii.gz|.npy] | <name>_3d_label[...]      <name>[...]_blackdict.pkl when --max_dropout_prob > 0
(`.npy` volumes are written as bool exactly like the reference, :94; NIfTI needs nibabel.)"""
from __future__ import annotations

import argparse
import csv
import os
import pickle
import re
import sys
from glob import glob

import numpy as np

from .tree2img import rasterize_forest, voxelize_forest


def natural_key(s: str):
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", s)]


def render_graph(file_path: str, args, resolution, img_res):
    from PIL import Image

    name = file_path.split("/")[-1].removesuffix(".csv")
    with open(file_path, newline="") as f:
        rows = list(csv.DictReader(f))
    if args.save_3d:
        vol, black_dict = voxelize_forest(rows, resolution, max_dropout_prob=args.max_dropout_prob, ignore_z=args.ignore_z)
        vname = name + ("_3d_label" if args.binarize else "_3d")
        if args.binarize:
            vol[vol < 0.1] = 0
            vol[vol >= 0.1] = 1
        if args.save_3d_as == ".nii.gz":
            try:
                import nibabel as nib
            except ImportError as e:
                raise RuntimeError("--save_3d_as .nii.gz needs nibabel, which is not installed; use --save_3d_as .npy") from e
            nib.save(nib.Nifti1Image(vol, np.eye(4)), os.path.join(args.out_dir, vname + ".nii.gz"))
        else:
            np.save(os.path.join(args.out_dir, vname + ".npy"), vol.astype(np.bool_))
        if args.max_dropout_prob > 0:
            with open(os.path.join(args.out_dir, vname + "_blackdict.pkl"), "wb") as f:
                pickle.dump(black_dict, f)
    if args.save_2d:
        img, black_dict = rasterize_forest(rows, img_res, args.mip_axis, max_dropout_prob=args.max_dropout_prob)
        if args.binarize:
            img[img < 0.1] = 0
            Image.fromarray(img.astype(np.uint8)).convert("1").save(os.path.join(args.out_dir, name + "_label.png"))
        else:
            Image.fromarray(img.astype(np.uint8)).save(os.path.join(args.out_dir, name + ".png"))
        if args.max_dropout_prob > 0:
            with open(os.path.join(args.out_dir, name + "_blackdict.pkl"), "wb") as f:
                pickle.dump(black_dict, f)


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
    p.add_argument("--threads", type=int, default=-1, help="accepted for compatibility")
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
    if world > 1:
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    for i, fp in enumerate(files):
        if i % world == rank:
            render_graph(fp, args, resolution, img_res)
    return 0


if __name__ == "__main__":
    sys.exit(main())
