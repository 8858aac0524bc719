"""Drop-in for the reference's generate_vessel_graph.py (same flags, same files on disk), with growth, rasterization and
voxelization running on the GPU through the batched, software-pipelined engine (`pipeline.Pipeline.run_pipelined`: several
growth loops in flight, post-processing and CSV text of finished batches beside them) instead of one CPU process per sample.

    python -m octa_autosegmentation_b200.generate_vessel_graph --config_file cfg.yml --num_samples 64 \
        [--threads T] [--debug] [--seed S] [--batch B] [--in_flight L] [--gather] [--Greenhouse.param_scale 3 ...dotted overrides]

Per sample, as generate_vessel_graph.py:24-89:  <output.directory>/<YYYYmmdd_HHMMSS>_<uuid4>/
    config.yml                      resolved config (yaml.dump)
    <dirname>.csv                   node1,node2,radius rows (arterial trees first), CRLF          (save_trees)
    art_ven_img_gray.npy | .nii.gz  uint8 volume, np.maximum(arterial, venous)                    (save_3D_volumes: npy | nifti)
    art_ven_img_gray.png            uint8 gray image, np.maximum(arterial raster, venous raster)  (save_2D_image)
New optional flags: --seed (sample i uses seed+i for BOTH generators; default: drawn from os.urandom, i.e. unseeded like the
reference), --batch (samples per growth loop), --in_flight (growth loops in flight), --threads (file-writer threads).  Under
torchrun every rank takes the samples i with i mod world_size == rank (no collective on the data path) and writes its own
files; with --gather the finished edge tables travel to rank 0 in ONE NCCL gather at the end and rank 0 writes every CSV.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import sys
import warnings
from datetime import datetime
from uuid import uuid4

import numpy as np
import yaml

from . import graph_io, growth
from .config import apply_cli_overrides_from_unknown_args, read_config
from .pipeline import Pipeline, shard_seeds


def prepare_output_dir(out_cfg: dict) -> str:
    # utilities.py:17-22
    d = os.path.join(os.path.abspath(out_cfg["directory"]), datetime.now().strftime("%Y%m%d_%H%M%S") + "_" + str(uuid4()))
    os.makedirs(d, exist_ok=True)
    return d


def volume_dimension(config: dict):
    """generate_vessel_graph.py:43 -- int(d) for d in greenhouse.simspace.shape * image_scale_factor."""
    return [int(d) for d in growth.simspace_shape(config) * config["output"]["image_scale_factor"]]


def write_sample(config: dict, csv_bytes, image, volume, stats=None) -> str:
    """Files of one grown sample (generate_vessel_graph.py:28-30,59-89); csv_bytes / image / volume / stats may be None.
    stats = (oxygen sinks, CO2 sources, per_step [iterations, 4], seconds per iteration, radius_list): output.save_stats."""
    out_cfg = config["output"]
    out_dir = prepare_output_dir(out_cfg)
    with open(os.path.join(out_dir, "config.yml"), "w") as f:
        yaml.dump(config, f)
    if csv_bytes is not None:
        with open(os.path.join(out_dir, out_dir.split("/")[-1] + ".csv"), "wb") as f:
            f.write(csv_bytes)
    if volume is not None:
        vol = volume.astype(np.uint8)
        if out_cfg["save_3D_volumes"] == "npy":
            np.save(f"{out_dir}/art_ven_img_gray.npy", vol)
        else:
            graph_io.save_nifti(f"{out_dir}/art_ven_img_gray.nii.gz", vol)          # nib.Nifti1Image(vol, np.eye(4)), :76-77
    if image is not None:
        graph_io.save_png(f"{out_dir}/art_ven_img_gray.png", image.astype(np.uint8))
    if stats is not None:
        from . import stats_plots
        oxys, co2s, per_step, seconds, radius_list = stats
        stats_plots.save_stats(out_dir, oxys, co2s, per_step, seconds)               # generate_vessel_graph.py:40-41
        stats_plots.plot_vessel_radii(out_dir, radius_list)                           # :88-89
    return out_dir


def stats_radius_list(out_cfg: dict, edges7: np.ndarray) -> np.ndarray:
    """The radius_list generate_vessel_graph.py:88-89 hands to plot_vessel_radii: rasterize_forest's 1.3 x radius of every edge
    (tree2img.py:82-83; arterial rows, then venous) when save_2D_image is on -- the list is reset before the rasters
    (:78-79) --, else voxelize_forest's plain radii (tree2img.py:196-238) when only volumes are written, else empty."""
    if out_cfg["save_2D_image"]:
        return edges7[:, 6] * 1.3
    if out_cfg.get("save_3D_volumes"):
        return edges7[:, 6].copy()
    return np.zeros(0)


def generate(config: dict, seeds, batch: int = 32, in_flight: int = 8, writer_threads: int = 4, device=None, gather: bool = False,
             on_progress=None):
    """Grow `seeds` and write the reference's files.  Returns (output dirs in seed order, {seed: edges7} when gather else {})."""
    out_cfg = config["output"]
    dims = volume_dimension(config)
    image_res = [*dims]
    del image_res[out_cfg["proj_axis"]]
    save3d = bool(out_cfg.get("save_3D_volumes"))
    save_stats = bool(out_cfg.get("save_stats"))
    pipe = Pipeline(config, device=device, volume_dims=dims, label_res=None, image_res=image_res, mip_axis=out_cfg["proj_axis"],
                    voxelize=save3d, growth_stats=save_stats)
    if save3d:
        in_flight = min(in_flight, 2)               # every buffer set holds a pinned copy of the batch's volumes
    batches = [list(seeds[k:k + batch]) for k in range(0, len(seeds), batch)]
    dirs, tables, futs = [], {}, []
    write_csv = bool(out_cfg["save_trees"]) and not gather
    with cf.ThreadPoolExecutor(max_workers=max(1, writer_threads)) as writers:
        for bi, out in enumerate(pipe.run_pipelined(batches, d2h=True, csv=write_csv, in_flight=in_flight, d2h_volume=save3d,
                                                    extra_slots=1 if save3d else None)):
            n = len(batches[bi])
            for i in range(n):
                img = np.array(out["image_host"][i]) if out_cfg["save_2D_image"] else None        # copies: the pinned buffers are recycled
                vol = np.array(out["volume_host"][i]) if save3d else None
                st = None
                if save_stats:
                    gs = out["growth_stats"]
                    # the loop grows the whole batch together: its device time, spread evenly over iterations and samples
                    sec = np.full(gs["iterations"], gs["loop_seconds"] / max(1, gs["iterations"] * n))
                    st = (gs["sinks"][i][0], gs["sinks"][i][1], gs["per_step"][i], sec,
                          stats_radius_list(out_cfg, np.concatenate(out["graphs"][i])))
                futs.append(writers.submit(write_sample, config, bytes(out["csv"][i]) if write_csv else None, img, vol, st))      # (copy: may be a view of a pinned buffer)
                if gather:
                    tables[batches[bi][i]] = np.concatenate(out["graphs"][i]).copy()
            if save3d:                                 # 157 MB per sample: let the files land before the next batch's copies
                for f in futs:
                    f.result()
            if on_progress:
                on_progress(sum(len(b) for b in batches[:bi + 1]))
        dirs = [f.result() for f in futs]
    return dirs, tables


def main(argv=None):
    parser = argparse.ArgumentParser(description="")
    parser.add_argument("--config_file", type=str, required=True)
    parser.add_argument("--num_samples", type=int, default=1)
    parser.add_argument("--debug", action="store_true")
    parser.add_argument("--threads", help="Number of file-writer threads. By default all available cores but one (max 8).", type=int, default=-1)
    parser.add_argument("--seed", type=int, default=None, help="base seed; sample i uses seed+i (default: unseeded)")
    parser.add_argument("--batch", type=int, default=32, help="samples per growth loop")
    parser.add_argument("--in_flight", type=int, default=8, help="growth loops in flight per GPU")
    parser.add_argument("--gather", action="store_true", help="torchrun only: one NCCL gather of the edge tables; rank 0 writes every CSV")
    args, unknown = parser.parse_known_args(argv)
    if args.debug:
        warnings.filterwarnings("error")
    assert os.path.isfile(args.config_file), f"Error: Your provided config path {args.config_file} does not exist!"
    config = read_config(args.config_file)
    apply_cli_overrides_from_unknown_args(config, unknown)
    assert config["output"].get("save_3D_volumes") in [None, "npy", "nifti"], \
        f"Your provided option {config['output'].get('save_3D_volumes')} for 'save_3D_volumes' does not exist. Choose one of 'null', 'npy' or 'nifti'."
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    gather = args.gather and world > 1
    if gather:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    base = args.seed if args.seed is not None else int.from_bytes(os.urandom(4), "little") % (2 ** 32 - args.num_samples - 1)
    seeds = shard_seeds(base, args.num_samples, rank, world)
    threads = args.threads if args.threads > 0 else max(1, min(8, (os.cpu_count() or 2) - 1))
    rc = 0
    try:
        dirs, tables = generate(config, seeds, batch=max(1, min(args.batch, max(len(seeds), 1))), in_flight=args.in_flight,
                                writer_threads=threads, device=torch.device("cuda", local), gather=gather,
                                on_progress=lambda k: print(f"[rank {rank}] generated {k}/{len(seeds)} vessel graphs", flush=True))
        if gather:
            from .distributed import gather_edge_tables
            allt = gather_edge_tables(tables, dst=0)
            if rank == 0 and config["output"]["save_trees"]:
                # rank 0 owns the CSV files of the whole job (sample order); its own samples already have their folders
                mine = dict(zip(seeds, dirs))
                for sid in sorted(allt):
                    d = mine.get(sid) or prepare_output_dir(config["output"])
                    graph_io.write_csv(os.path.join(d, d.split("/")[-1] + ".csv"), allt[sid])
    except Exception as e:       # the reference swallows worker exceptions (futures never read); we report them
        print(f"[rank {rank}] FAILED: {e!r}", file=sys.stderr)
        rc = 1
    if gather:
        import torch.distributed as dist
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
