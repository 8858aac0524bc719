#!/usr/bin/env python
"""Benchmark of the vessel-graph hot path (BASELINE.json metric: synthetic graphs/sec incl. the
1216^2 raster; achieved HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one batch of B fresh samples per GPU (BASELINE config #2:
B = 64, docker / default 3x3 mm^2 config): seeded growth (250 iterations) -> exact radii + edge rows
-> voxel volume for the request [1216,1216,16] (1216x1216x53 uint16) -> 1216^2 label raster and
304^2 gray image.  Every step uses NEW seeds (nothing can be cached).  N > 1 is launched by torchrun,
one rank per GPU; samples are independent, so ranks own disjoint seeds and there is no data-path
collective (weak scaling: per-GPU batch fixed).  Rank 0 prints ONE JSON line.

  value : whole-job graphs/s with results left in HBM (CUDA events around the timed steps; the step
          contains host work -- D2H of the tree topology, libm radius replay, H2D of edge rows --
          which the events bracket too).
  e2e   : the same through host buffers: + byte-exact CSV text of every graph, + D2H of the 1216^2 label
          and the 304^2 image into pinned memory (the 157 MB volumes stay in HBM unless --d2h-volume).
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

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before CUDA initialises: one hardware queue per stream (see _lib.py)

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = [1216, 1216, 16]
WORKLOAD = ("BASELINE config #2: batch of %d seeded samples per GPU, default 3x3 mm^2 macular config "
            "(= docker/vessel_graph_gen_docker_config.yml), growth -> edge rows -> voxelize [1216,1216,16] "
            "(1216x1216x53 u16) -> 1216^2 label + 304^2 image")


def load_traffic(batch):
    """DRAM bytes per launch of vox_col_kernel -- NOT measured in this run: read from the committed `ncu --set full` capture
    of the same kernel, batch and volume (profiles/r02_vox_col_traffic.json; the line says so in roofline.traffic_source)."""
    p = os.path.join(ROOT, "profiles", "r02_vox_col_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            t = json.load(f)
        if int(t.get("batch", -1)) == int(batch):
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, index=0):
        self.samples, self.reasons, self.proc, self.index = [], set(), None, index

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
        threading.Thread(target=self._read, daemon=True).start()

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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.samples[0][1] if self.samples else None,
                "samples": len(sm), "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's own algorithm on the host cores.  The reference is pure Python and cannot
# travel to the GPU box (no /root/reference there), so this is the oracle PORT (oracle/growth_oracle.cpp +
# oracle/voxelize_oracle.c, byte-identical to the reference on the committed goldens), one process per
# core like generate_vessel_graph.py:112-126.
# ----------------------------------------------------------------------------------------------
def _cpu_one(seed):
    from oracle import growth_oracle as go, vox_oracle
    from octa_autosegmentation_b200.config import default_config
    t0 = time.perf_counter()
    art, ven, _ = go.run(default_config(), seed)
    t1 = time.perf_counter()
    vox_oracle.voxelize_edges(np.concatenate([art, ven]), DIMS)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_arm(n_graphs, workers, base_seed):
    import concurrent.futures as cf
    from oracle import growth_oracle as go, vox_oracle
    go.build()
    vox_oracle.build()
    t0 = time.perf_counter()
    with cf.ProcessPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(_cpu_one, [base_seed + i for i in range(n_graphs)]))
    dt = time.perf_counter() - t0
    return n_graphs / dt, dt, float(np.mean([p[0] for p in parts])), float(np.mean([p[1] for p in parts]))


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 2
    workers = max(1, cores - 1)
    sample = workers * 2
    for w in range(args.warmup):
        cpu_arm(workers, workers, 10_000 + w * 100)
    t0 = time.perf_counter()
    for s in range(args.steps):
        val, dt, tg, tv = cpu_arm(sample, workers, 20_000 + s * 1000)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt
    print(json.dumps({
        "impl": "reference", "metric": "graphs_per_sec", "value": val, "unit": "graphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD % 64 + " [CPU arm: growth + CSV rows + voxelize; the matplotlib 2-D stage is "
                   "excluded: matplotlib is not installed]", "sample_graphs_per_step": sample},
        "cpu_baseline": {"value": val, "unit": "graphs/s", "cores": workers, "kind": "port",
                         "sample": "%d graphs/step of the 64-graph batch; C++/C port of the reference "
                                   "(growth %.2f s + voxelize %.2f s per graph per core)" % (sample, tg, tv)},
        "e2e": {"value": val, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def pin_rank_to_cores():
    """One process per GPU: rank r of W keeps to its own slice of the host cores (launch-issuing grower threads, the libm radius
    replay and the CSV pool of 8 ranks otherwise migrate over all cores of the box).  OCTA_NO_PIN=1 disables."""
    world, local = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))), int(os.environ.get("LOCAL_RANK", "0"))
    if world <= 1 or os.environ.get("OCTA_NO_PIN") == "1" or not hasattr(os, "sched_setaffinity"):
        return
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if per >= 2:
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
    except OSError:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16, help="timed steps of --batch samples; the pipeline fills and drains inside the timed "
                                                            "region (a growth loop has ~0.35 s of latency), so few steps measure mostly that")
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--d2h-volume", action="store_true")
    ap.add_argument("--no-gan", action="store_true", help="skip the config #5 (GAN contrast adaptation) leg")
    ap.add_argument("--in-flight", type=int, default=7, help="growth loops (sub-batches) in flight per GPU in the pipelined API")
    ap.add_argument("--sub-batch", type=int, default=64, help="samples per growth loop; a step's --batch samples are fed to the "
                                                              "pipelined API as batch / sub-batch consecutive batches")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = the headline line (default), 3 / 4 / 5 = bench_configs.py")
    ap.add_argument("--samples", type=int, default=0, help="--config 3: samples of the regeneration job (default 500)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    pin_rank_to_cores()
    if args.config != 2:
        import bench_configs
        return {3: bench_configs.run_config3, 4: bench_configs.run_config4, 5: bench_configs.run_config5}[args.config](args)

    import torch
    import torch.distributed as dist
    from octa_autosegmentation_b200 import _lib, tree2img
    from octa_autosegmentation_b200.config import default_config
    from octa_autosegmentation_b200.pipeline import Pipeline

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
    SB = args.sub_batch if 0 < args.sub_batch < B and B % args.sub_batch == 0 else B
    NSUB = B // SB
    W = max(args.warmup, 3)
    pipe = Pipeline(default_config(), device=dev, volume_dims=DIMS, host_threads=max(1, len(os.sched_getaffinity(0)) if world > 1 else (os.cpu_count() or 2)))
    counter = [0]

    def seeds():
        # fresh seeds every step, disjoint across ranks
        s0 = 1_000_000 + counter[0] * B * world + rank * B
        counter[0] += 1
        return [s0 + i for i in range(B)]

    phase = {"grow_ms": 0.0, "vox_ms": 0.0, "n": 0, "sumA": 0, "sumM": 0, "sumP": 0, "sumS": 0, "V": 0, "E": 0}
    def account(out, d2h):
        phase["grow_ms"] += out["grow_device_ms"]
        phase["n"] += 1
        for st in out["stats"]:
            phase["sumA"] += st["sum_A"]; phase["sumM"] += st["sum_M"]; phase["sumP"] += st["sum_P"]; phase["sumS"] += st["sum_S"]
            phase["V"] += st["n_art_nodes"] + st["n_ven_nodes"]
        phase["E"] += int(out["offsets"][-1])
        return out

    def step(d2h):
        return account(pipe.run(seeds(), d2h=d2h, csv=d2h), d2h)

    def run_steps(k, d2h):
        # the public batched API: a step's B samples go in as B / SB batches; several growth loops are in flight while
        # voxelize / raster / CSV of finished batches run on another stream
        last = None
        batches = []
        for _ in range(k):
            s = seeds()
            batches += [s[j:j + SB] for j in range(0, B, SB)]
        h2d = d2h_b = 0
        for i, out in enumerate(pipe.run_pipelined(batches, d2h=d2h, csv=d2h, in_flight=args.in_flight, d2h_volume=d2h and args.d2h_volume,
                                                   extra_slots=1 if (d2h and args.d2h_volume) else None)):      # (10 GB of pinned volumes per set)
            last = account(out, d2h)
            if d2h:
                h2d += int(out["h2d_bytes"]); d2h_b += int(out["d2h_bytes"])
        if d2h and last is not None:
            last = dict(last)
            last["h2d_bytes"], last["d2h_bytes"] = h2d // k, d2h_b // k      # per step of B samples
        return last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        last = fn(steps)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, last

    NSETS = Pipeline.buffer_sets(args.in_flight)
    run_steps(max(W, (NSETS + NSUB - 1) // NSUB), False)       # (also allocates every buffer set of the pipelined path)
    for k in phase:
        phase[k] = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ms, _ = timed(lambda k: run_steps(k, False), args.steps)
    launches = _lib.launch_count() - n0
    ph = dict(phase)
    # voxelizer alone, on its stream, over the last batch's edges (the HBM-bound kernel of the path)
    out = step(False)
    offs = out["offsets"]
    edges_dev = pipe._buf["edges_dev0"]
    vol = out["volume"]
    torch.cuda.synchronize()
    ve = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    nrep = 5
    ve[0].record()
    for _ in range(nrep):
        tree2img.voxelize_batch_device(edges_dev, offs, DIMS, out=vol, workspace=pipe._buf.get("vox_ws0"))
    ve[1].record()
    torch.cuda.synchronize()
    vox_ms = ve[0].elapsed_time(ve[1]) / nrep
    del out, vol, edges_dev
    pipe.release_device_buffers()                              # 10 GB of volumes per device-resident buffer set
    NSETS_H = Pipeline.buffer_sets(args.in_flight, True, 1 if args.d2h_volume else None)
    run_steps(max(1, (NSETS_H + NSUB - 1) // NSUB), True)      # (pinned result buffers of every set)
    ms_e2e, last = timed(lambda k: run_steps(k, True), max(2, args.steps))
    clocks = sampler.stop() if rank == 0 else None
    gan_line = None
    if rank == 0 and world == 1 and not args.no_gan:      # (N = 1 only, like cpu_baseline: the other ranks would idle behind it)
        # BASELINE config #5 tail: the 304^2 images of the last batch -> background noise -> resnetGenerator9
        # (random-init weights, synthetic backgrounds) -> uint8 images; tensor-core rate of its 3x3 convolutions
        from octa_autosegmentation_b200 import gan
        sd = gan.random_init_state_dict(0)
        G = gan.ResnetGenerator9(sd, image_size=(304, 304), max_images=32, device=dev)
        img = step(False)["image"]
        nimg = int(img.shape[0])
        bg = torch.randint(0, 255, tuple(img.shape), device=dev, dtype=torch.uint8)
        sp_seeds = list(range(nimg))
        gan.contrast_adapt(G, img, bg, sp_seeds)
        torch.cuda.synchronize()
        ge = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        l0 = _lib.launch_count()
        ge[0].record()
        for _ in range(3):
            g_out = gan.contrast_adapt(G, img, bg, sp_seeds)
        ge[1].record()
        torch.cuda.synchronize()
        gan_ms = ge[0].elapsed_time(ge[1]) / 3
        tfl = nimg * gan.conv_flops_per_image(304, 304) / (gan_ms * 1e-3) / 1e12
        tpeak = None
        try:
            tpeak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            pass
        gan_line = {"workload": "BASELINE config #5 tail: %d rasters 304^2 (this step's images) -> ScaleIntensity + background speckle -> "
                                "resnetGenerator9 (random-init, bf16 tcgen05 convolutions) -> uint8 images, device resident" % nimg,
                    "images_per_sec": nimg / (gan_ms * 1e-3), "ms_per_batch": gan_ms, "gpu_launches": int((_lib.launch_count() - l0) // 3),
                    "roofline": {"bound": "tensor", "achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": (tfl / tpeak) if tpeak else None,
                                 "traffic": None, "kernel": "whole forward; algorithmic flops = the 22 3x3 convolutions (177.2 GF/image); "
                                                            "k_gan_conv3 alone: see profiles/"},
                    "mean_gray": float(g_out.float().mean().item())}
        if not args.no_cpu_baseline and world == 1:
            import time as _t
            from oracle import gan_oracle as go
            xs = torch.rand(2, 1, 304, 304)
            with torch.no_grad():
                go.generator_forward(sd, xs[:1])
                t0 = _t.time()
                go.generator_forward(sd, xs)
                dt = _t.time() - t0
            gan_line["cpu_baseline"] = {"value": 2 / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                        "sample": "2 images through the fp32 PyTorch restatement of the generator (the reference runs it on CPU, "
                                                  "docker/trained_models/GAN/config.yml General.device: cpu)"}
        G.close()

    if rank == 0:
        peak, peak_src = load_peaks()
        shape = tree2img.voxel_volume_shape(DIMS)
        vol_bytes = int(np.prod(shape)) * 2
        E_last = int(offs[-1])
        vox_alg = 56 * E_last + vol_bytes * B                         # B_vox = 56 E + 2 X Y Z' per graph
        n = max(ph["n"], 1)                                          # growth loops (sub-batches of SB samples) accounted
        nsteps = n / NSUB
        b_grow = (28 * ph["sumA"] + 24 * ph["sumM"] + 24 * 2000 * 250 * SB * n + 32 * ph["sumP"] + 24 * ph["sumS"] + 40 * ph["V"]) / nsteps
        grow_ms = ph["grow_ms"] / n
        vox_ach = vox_alg / (vox_ms * 1e-3) / 1e9
        # growth loops of several batches overlap: bytes of one step over the step time (the loop dominates the timeline)
        grow_ach = b_grow / (ms * 1e-3) / 1e9
        line = {
            "metric": "graphs_per_sec", "value": B * world / (ms * 1e-3), "unit": "graphs/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD % B, "batch_per_gpu": B, "edges_per_graph_mean": ph["E"] / n / SB,
                       "seeds": "fresh every step (1000000 + step*B*world + rank*B + i)",
                       "l2": "per step %.1f GB of volumes + ~1.6 GB of growth state >> 126 MB L2 (no flush needed)" % (vol_bytes * B / 1e9),
                       "pipelining": "Pipeline.run_pipelined: a step's %d samples enter as %d batches of %d; %d growth loops in flight per "
                                     "GPU, each on its own pair of high-priority streams (CUDA_DEVICE_MAX_CONNECTIONS=%s); voxelize / "
                                     "raster / CSV of finished batches run on another stream and a worker thread; the timed region "
                                     "contains every step completely (all batches in, all results out)"
                                     % (B, NSUB, SB, args.in_flight, os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS")),
                       "in_flight_batches": args.in_flight, "sub_batch": SB,
                       "phase_ms": {"growth_loop_device": grow_ms, "voxelize_4_kernels": vox_ms, "step": ms}},
            "gpu_launches": int(launches),
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": int(last["h2d_bytes"]),
                    "d2h_bytes_per_step": int(last["d2h_bytes"]),
                    "note": "host API: + CSV text of every graph (byte-exact), + D2H of label (1216^2 u8) and image (304^2 u8) "
                            "into pinned memory; growth topology D2H and edge-row H2D are inside both numbers"},
            "roofline": {"bound": "hbm", "achieved": vox_ach, "peak": peak, "unit": "GB/s", "frac": vox_ach / peak,
                         "traffic": load_traffic(B), "traffic_source": "committed ncu --set full capture of the same launch shape "
                         "(profiles/r02_vox_col_traffic.json), not measured in this run",
                         "peak_source": peak_src, "kernel": "vox_col_kernel (+prep/scan/fill, <1 %)",
                         "algorithmic_bytes_per_launch": vox_alg, "ms_per_launch": vox_ms,
                         "note": "the HBM-bound kernel of the path; see roofline_growth for the phase that dominates step time"},
            "roofline_growth": {"bound": "hbm", "achieved": grow_ach, "peak": peak, "unit": "GB/s", "frac": grow_ach / peak,
                                "traffic": None, "kernel": "growth loop, 15 launches x 250 iterations on two streams (latency / pair-scan bound; "
                                                           "state is L2-resident)", "algorithmic_bytes_per_step": b_grow,
                                "ms_per_step": ms, "device_ms_per_batch_loop": grow_ms},
            "clocks": clocks,
        }
        if gan_line is not None:
            line["config5_gan"] = gan_line
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 2
            workers = max(1, cores - 1)
            val, dt, tg, tv = cpu_arm(workers, workers, 30_000)
            line["cpu_baseline"] = {"value": val, "unit": "graphs/s", "cores": workers, "kind": "port",
                                    "sample": "%d graphs (one per worker) of the 64-graph batch; C++/C port of the reference, "
                                              "growth %.2f s + voxelize %.2f s per graph per core; 2-D matplotlib stage "
                                              "excluded (not installed)" % (workers, tg, tv)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
