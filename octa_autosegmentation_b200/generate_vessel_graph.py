"""Drop-in for the reference's generate_vessel_graph.py (same flags, same files on disk), with the growth,
rasterization and voxelization running on the GPU in batches instead of one CPU process per sample.

    python -m octa_autosegmentation_b200.generate_vessel_graph --config_file cfg.yml --num_samples 64 \
        [--threads T] [--debug] [--seed S] [--batch B] [--Greenhouse.param_scale 3 ...dotted overrides]

Per sample, as generate_vessel_graph.py:24-89:  <output.directory>/<YYYYmmdd_HHMMSS>_<uuid4>/
    config.yml                      resolved config (yaml.dump)
    <dirname>.csv                   node1,node2,radius rows (arterial trees first), CRLF          (save_trees)
    art_ven_img_gray.npy            uint8 volume, np.maximum(arterial, venous)                    (save_3D_volumes: npy)
    art_ven_img_gray.png            uint8 gray image, np.maximum(arterial raster, venous raster)  (save_2D_image)
New optional flags: --seed (sample i uses seed+i for BOTH generators; default: drawn from os.urandom, i.e. unseeded
like the reference), --batch (samples per GPU launch).  Under torchrun every rank takes the samples i with
i mod world_size == rank and writes its own files (no collective on the data path).
"""
from __future__ import annotations

import argparse
import os
import sys
import warnings
from datetime import datetime
from uuid import uuid4

import numpy as np
import yaml

from . import graph_io, growth, tree2img
from .config import apply_cli_overrides_from_unknown_args, read_config
from .pipeline import shard_seeds


def prepare_output_dir(out_cfg: dict) -> str:
    # utilities.py:17-22
    d = os.path.join(os.path.abspath(out_cfg["directory"]), datetime.now().strftime("%Y%m%d_%H%M%S") + "_" + str(uuid4()))
    os.makedirs(d, exist_ok=True)
    return d


def write_sample(config: dict, art: np.ndarray, ven: np.ndarray) -> str:
    """generate_vessel_graph.py:28-30,43-86 for one grown sample."""
    from PIL import Image

    out_cfg = config["output"]
    out_dir = prepare_output_dir(out_cfg)
    with open(os.path.join(out_dir, "config.yml"), "w") as f:
        yaml.dump(config, f)
    shape = np.array([config["Greenhouse"]["SimulationSpace"][k] for k in ("no_voxel_x", "no_voxel_y", "no_voxel_z")])
    volume_dimension = [int(d) for d in shape * out_cfg["image_scale_factor"]]
    if out_cfg["save_trees"]:
        name = out_dir.split("/")[-1]
        graph_io.write_csv(os.path.join(out_dir, name + ".csv"), np.concatenate([art, ven]))
    if out_cfg.get("save_3D_volumes"):
        vol = np.maximum(tree2img.voxelize_edges(art, volume_dimension), tree2img.voxelize_edges(ven, volume_dimension)).astype(np.uint8)
        if out_cfg["save_3D_volumes"] == "npy":
            np.save(f"{out_dir}/art_ven_img_gray.npy", vol)
        else:
            try:
                import nibabel as nib
            except ImportError as e:
                raise RuntimeError("save_3D_volumes: nifti needs nibabel, which is not installed; use 'npy'") from e
            nib.save(nib.Nifti1Image(vol, np.eye(4)), f"{out_dir}/art_ven_img_gray.nii.gz")
    if out_cfg["save_2D_image"]:
        image_res = [*volume_dimension]
        del image_res[out_cfg["proj_axis"]]
        a = tree2img.raster_edges(art, image_res, out_cfg["proj_axis"])
        v = tree2img.raster_edges(ven, image_res, out_cfg["proj_axis"])
        Image.fromarray(np.maximum(a, v).astype(np.uint8)).save(f"{out_dir}/art_ven_img_gray.png")
    if out_cfg.get("save_stats"):
        warnings.warn("output.save_stats (matplotlib statistic plots) is not produced by the GPU path")
    return out_dir


def main(argv=None):
    parser = argparse.ArgumentParser(description="")
    parser.add_argument("--config_file", type=str, required=True)
    parser.add_argument("--num_samples", type=int, default=1)
    parser.add_argument("--debug", action="store_true")
    parser.add_argument("--threads", help="Accepted for compatibility (host-side writer threads).", type=int, default=-1)
    parser.add_argument("--seed", type=int, default=None, help="base seed; sample i uses seed+i (default: unseeded)")
    parser.add_argument("--batch", type=int, default=64, help="samples per GPU launch")
    args, unknown = parser.parse_known_args(argv)
    if args.debug:
        warnings.filterwarnings("error")
    assert os.path.isfile(args.config_file), f"Error: Your provided config path {args.config_file} does not exist!"
    config = read_config(args.config_file)
    apply_cli_overrides_from_unknown_args(config, unknown)
    assert config["output"].get("save_3D_volumes") in [None, "npy", "nifti"], \
        f"Your provided option {config['output'].get('save_3D_volumes')} for 'save_3D_volumes' does not exist. Choose one of 'null', 'npy' or 'nifti'."
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    base = args.seed if args.seed is not None else int.from_bytes(os.urandom(4), "little") % (2 ** 32 - args.num_samples - 1)
    seeds = shard_seeds(base, args.num_samples, rank, world)
    done, failed = 0, []
    ctx = None
    pending = None          # the previous chunk's files are written by a worker thread while the next chunk grows

    def write_chunk(graphs):
        for art, ven in graphs:
            write_sample(config, art, ven)
        return len(graphs)

    import concurrent.futures as cf
    with cf.ThreadPoolExecutor(max_workers=1) as writer:
        for k in range(0, len(seeds), args.batch):
            chunk = seeds[k:k + args.batch]
            try:
                if ctx is None:
                    ctx = growth.GrowContext(config, min(args.batch, len(seeds)))
                graphs, _, _ = ctx.run(chunk)
            except Exception as e:       # the reference swallows worker exceptions (futures never read); we report them
                failed.append((chunk, repr(e)))
                continue
            if pending is not None:
                done += pending.result()
                print(f"[rank {rank}] generated {done}/{len(seeds)} vessel graphs", flush=True)
            pending = writer.submit(write_chunk, graphs)
        if pending is not None:
            done += pending.result()
            print(f"[rank {rank}] generated {done}/{len(seeds)} vessel graphs", flush=True)
    if ctx is not None:
        ctx.close()
    if failed:
        for chunk, msg in failed:
            print(f"[rank {rank}] FAILED seeds {chunk[0]}..{chunk[-1]}: {msg}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
