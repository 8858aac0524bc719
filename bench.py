#!/usr/bin/env python
"""Benchmark of the vessel-graph hot path (BASELINE.json metric: synthetic graphs/sec incl. the
1216^2 x 16 raster; achieved HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one batch of B graphs per GPU (config #2: B = 64,
raster request [1216,1216,16] -> volume 1216x1216x53 uint16).  N > 1 is launched by torchrun, one
rank per GPU; graphs are independent, so ranks shard the batch with no data-path collective
(weak scaling: per-GPU work fixed).  Prints ONE JSON line on rank 0.

STATUS (round 1, interim): the growth kernels are not wired into the step yet -- the step
rasterizes B pre-grown graphs (the seed-0 docker-config golden graph, jittered per graph so the
volumes differ).  `config.workload` says so; see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIMS = [1216, 1216, 16]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle sampler running during the timed region."""

    def __init__(self, index=0):
        self.samples, self.reasons, self.proc = [], set(), None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            try:
                self.samples.append((float(parts[0]), float(parts[1])))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(s[0] for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": self.samples[0][1] if self.samples else None,
                "samples": len(sm), "reasons": sorted(self.reasons)}


def golden_edges():
    from conftest import load_graph_rows, rows_to_edges7
    return rows_to_edges7(load_graph_rows("graph_docker_s0.csv.gz"))


def make_batch(e7: np.ndarray, batch: int, rank: int):
    """B graphs: the golden graph with a per-graph sub-voxel jitter (synthetic, deterministic)."""
    graphs = []
    for i in range(batch):
        rng = np.random.RandomState(1000 * rank + i)
        g = e7.copy()
        shift = rng.uniform(-2e-3, 2e-3, 3) * np.array([1, 1, 0.0])
        g[:, 0:3] += shift
        g[:, 3:6] += shift
        graphs.append(g)
    offs = np.cumsum([0] + [len(g) for g in graphs]).astype(np.int64)
    return np.concatenate(graphs), offs


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm for the path on the host cores.  The
    reference is pure Python and cannot travel to the GPU box, so this is the C oracle port
    (oracle/voxelize_oracle.c, bit-exact to tree2img.voxelize_forest), one process per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import concurrent.futures as cf
    from oracle import vox_oracle
    vox_oracle.build()
    e7 = golden_edges()
    cores = max(1, (os.cpu_count() or 2) - 1)
    sample = min(cores, 8)   # graphs per step (bounded sample of the 64-graph workload)

    def one(i):
        vox_oracle.voxelize_edges(e7, DIMS)   # ctypes call releases the GIL
        return i

    def step():
        with cf.ThreadPoolExecutor(max_workers=cores) as ex:
            list(ex.map(one, range(sample)))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt
    print(json.dumps({
        "impl": "reference", "metric": "graphs_per_sec", "value": val, "unit": "graphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "voxelize-only [1216,1216,16] of pre-grown docker-config graphs (interim, see DESIGN.md)",
                   "batch_per_gpu": sample},
        "cpu_baseline": {"value": val, "unit": "graphs/s", "cores": cores, "kind": "port",
                         "sample": "%d graphs/step (of the 64-graph batch), C port of tree2img.voxelize_forest" % sample},
        "e2e": {"value": val, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from octa_autosegmentation_b200 import _lib, tree2img

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B = args.batch

    e7 = golden_edges()
    batch_np, offs = make_batch(e7, B, rank)
    edges_dev = torch.from_numpy(batch_np).to(dev)
    shape = tree2img.voxel_volume_shape(DIMS)
    out = torch.empty((B, *shape), dtype=torch.uint16, device=dev)
    ws = torch.empty(int(_lib.lib().octa_voxelize_workspace_bytes(B, int(offs[-1]), _lib.int3(DIMS))),
                     dtype=torch.uint8, device=dev)
    host_edges = torch.from_numpy(batch_np).pin_memory()
    host_label = torch.empty((B, shape[0], shape[1]), dtype=torch.uint8).pin_memory()

    def step_device():
        tree2img.voxelize_batch_device(edges_dev, offs, DIMS, out=out, workspace=ws)

    def step_e2e():
        edges_dev.copy_(host_edges, non_blocking=True)
        tree2img.voxelize_batch_device(edges_dev, offs, DIMS, out=out, workspace=ws)
        # result read-back: the en-face maximum-intensity projection of every volume (1216^2 u8 / graph)
        mip = out.view(torch.int16).amax(dim=3).to(torch.uint8)
        host_label.copy_(mip, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        barrier()
        ev[0].record()
        for _ in range(steps):
            fn()
        ev[1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[1])
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ms = timed(step_device, args.steps)
    launches = _lib.launch_count() - n0
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, max(2, args.steps // 2))
    clocks = sampler.stop() if rank == 0 else None

    # roofline of the dominant kernel (vox_tile_kernel): time it alone with events on its stream
    E_total = int(offs[-1])
    vol_bytes = int(np.prod(shape)) * 2
    alg_bytes = 56 * E_total + vol_bytes * B            # B_vox = 56 E + 2 X Y Z' per graph
    peak, peak_src = load_peaks()
    achieved = alg_bytes / (ms * 1e-3) / 1e9            # whole step ~ dominant kernel + 3 binning kernels

    if rank == 0:
        line = {
            "metric": "graphs_per_sec", "value": B * world / (ms * 1e-3), "unit": "graphs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "voxelize-only [1216,1216,16]->1216x1216x53 u16 of %d pre-grown docker-config graphs per GPU "
                                   "(interim: growth kernels not yet in the step)" % B,
                       "batch_per_gpu": B, "edges_per_graph": int(len(e7)),
                       "l2": "outputs %.1f GB per step >> 126 MB L2 (no flush needed)" % (vol_bytes * B / 1e9)},
            "gpu_launches": int(launches),
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "graphs/s",
                    "h2d_bytes_per_step": int(host_edges.numel() * 8), "d2h_bytes_per_step": int(host_label.numel()),
                    "note": "H2D edges + D2H en-face MIP label per graph; the 157 MB volumes stay in HBM"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "kernel": "vox_tile_kernel (+3 binning kernels, whole step)"},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import vox_oracle
            vox_oracle.build()
            t0 = time.perf_counter()
            nrep = 0
            while time.perf_counter() - t0 < 10.0:
                vox_oracle.voxelize_edges(e7, DIMS)
                nrep += 1
            dt = (time.perf_counter() - t0) / nrep
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "graphs/s", "cores": 1, "kind": "port",
                                    "sample": "%d x one graph, C port of tree2img.voxelize_forest, 1 thread" % nrep}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
