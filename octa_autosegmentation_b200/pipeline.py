"""Batched hot path: seeds -> vessel graphs (CSV rows) -> 3-D voxel volume -> 2-D label / gray image.

This is the device-side composition of the C-ABI entry points that `generate_vessel_graph.py`
(growth, CSV, 304^2 gray image, optional 3-D volume) and `visualize_vessel_graphs.py --resolution
1216,1216,16 [--binarize]` (1216^2 image / label, optional 3-D volume) drive per sample in the
reference.  torch is used for device memory, pinned host buffers and streams only."""
from __future__ import annotations

import collections
import concurrent.futures as cf
import ctypes
import os
import threading
import time
from typing import Sequence

import numpy as np

from . import graph_io, growth, tree2img


class Pipeline:
    def __init__(self, config: dict, device=None, volume_dims: Sequence[int] = (1216, 1216, 16),
                 label_res: Sequence[int] = (1216, 1216), image_res: Sequence[int] = (304, 304), mip_axis: int = 2,
                 voxelize: bool = True, host_threads: int = 0, growth_stats: bool = False):
        import torch

        self.torch = torch
        self.config = config
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.volume_dims = [int(d) for d in volume_dims]
        self.label_res = [int(d) for d in label_res] if label_res is not None else None
        self.image_res = [int(d) for d in image_res]
        self.mip_axis = int(mip_axis)
        self.voxelize = bool(voxelize)
        # output.save_stats (greenhouse.py:401-441): every result also carries out["growth_stats"] = {"per_step": int32
        # [n, iterations, 4], "sinks": [(oxygen sinks [k, 3], CO2 sources [m, 3]) per sample], "iterations", "loop_seconds"}
        self.growth_stats = bool(growth_stats)
        ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)     # (ranks are pinned to core slices)
        self.host_threads = host_threads or max(1, ncores - 1)
        self._buf = {}
        self._grows = {}           # growth contexts by index (run_pipelined keeps several batches in flight)
        self._grow_locks = {}
        self._post_stream = None
        self._h2d_done = {}
        self._csv_pool = None
        self.edge_cap = 24000      # rows reserved per sample in the pinned edge buffer (docker config: ~13 k)
        # Who writes the CSV text (same bytes either way, tests/test_csv_device_gpu.py): the host writer costs 3.7 ms of a core per
        # docker-config graph -- 2.4 cores per GPU at 650 graphs/s --, the device writer ~1 ms of GPU time and 90 MB of D2H per batch
        # of 64.  Measured on one B200: 16 cores 649 (host) vs 638 (device) graphs/s, 4 cores (the share of a rank on an 8-GPU box)
        # 593 vs 621: the device writes when this process has fewer than 8 cores.  OCTA_CSV=host|device overrides.
        mode = os.environ.get("OCTA_CSV", "auto")
        self.device_csv = (mode == "device") or (mode != "host" and ncores < 8)

    def _tensor(self, key, shape, dtype, pinned=False):
        t = self._buf.get(key)
        n = int(np.prod(shape))
        if t is None or t.numel() < n:
            if pinned:
                t = self.torch.empty(n, dtype=dtype).pin_memory()
            else:
                t = self.torch.empty(n, dtype=dtype, device=self.device)
            self._buf[key] = t
        return t[:n].view(*shape)

    def release_device_buffers(self):
        """Drop every device tensor of the buffer sets (device-resident results handed out earlier become invalid) and give
        the memory back to the driver: a batch of 64 volumes [1216,1216,53] u16 is 10 GB per set."""
        for k in [k for k, t in self._buf.items() if t.is_cuda]:
            del self._buf[k]
        self._h2d_done.clear()
        self.torch.cuda.empty_cache()

    @property
    def _grow(self):
        return self._grows.get(0)

    def _grow_stage(self, seeds: Sequence[int], slot: int = 0, ctx: int = 0) -> dict:
        """Growth of one batch (blocking; a growth context owns two high-priority streams) into pinned edge rows."""
        torch = self.torch
        lock = self._grow_locks.setdefault(ctx, threading.Lock())
        t_sub = time.perf_counter()
        with lock, torch.cuda.device(self.device):
            t_g0 = time.perf_counter()
            g = self._grows.get(ctx)
            if g is None or g.max_graphs < len(seeds):
                if g is not None:
                    g.close()
                g = self._grows[ctx] = growth.GrowContext(self.config, len(seeds))
            n = len(seeds)
            cap = n * self.edge_cap
            host_edges = self._tensor("edges_host%d" % slot, (cap, 7), torch.float64, pinned=True)
            # The loop writes its rows into a buffer private to this growth context; they move into the slot's pinned rows only
            # after the slot's previous upload has executed on the (default-priority, possibly lagging) post stream -- device-
            # resident results are handed out before that stream has run.  Waiting here, after ~0.4 s of growth, never blocks in
            # practice; the same wait before the loop delayed every start by the post stream's lag (482 -> 450 graphs/s).
            ctx_edges = self._tensor("edges_ctx%d" % ctx, (cap, 7), torch.float64, pinned=True)
            tr = np.zeros((n, 4096, 4), dtype=np.int32) if self.growth_stats else None
            offs, n_art, stats, grow_ms = g.run_packed(seeds, ctx_edges.numpy(), trace=tr)
            gstats = None
            if self.growth_stats:       # the context keeps the final sink lists until its next run: fetched under its lock
                n_it = int(stats[0]["n_iters"]) if n else 0
                gstats = {"per_step": tr[:, :n_it].copy(), "sinks": [g.sinks(i) for i in range(n)], "iterations": n_it,
                          "loop_seconds": grow_ms * 1e-3}
            ev = self._h2d_done.get(slot)
            if ev is not None:
                ev.synchronize()
            E = int(offs[-1])
            # plain memmove (ctypes releases the GIL): torch's copy_ would open a 16-thread OpenMP region per call, from every
            # grower thread at once, next to the CSV pool
            ctypes.memmove(host_edges.data_ptr(), ctx_edges.data_ptr(), E * 56)
            return {"n": n, "cap": cap, "host_edges": host_edges, "offs": offs, "n_art": n_art, "stats": stats, "grow_ms": grow_ms,
                    "growth_stats": gstats,
                    "trace": {"ctx": ctx, "slot": slot, "t_submit": t_sub, "t_grow0": t_g0, "t_grow1": time.perf_counter()}}

    def _post_stage(self, g: dict, slot: int, d2h: bool, csv: bool, stream=None, d2h_volume: bool = False,
                    shared_device: bool = False) -> dict:
        """Edge rows -> device, voxelize, 2-D rasters, optional D2H + CSV text, all ENQUEUED on `stream` (default: current).
        Nothing here waits for the device: out["ready"] is recorded behind the last operation and `_finish` (or the caller)
        waits on it; the CSV text is formatted by a thread pool meanwhile."""
        torch = self.torch
        t_p0 = time.perf_counter()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream() if stream is None else stream
            with torch.cuda.stream(stream):
                n, cap, host_edges, offs, n_art = g["n"], g["cap"], g["host_edges"], g["offs"], g["n_art"]
                # Host-buffer mode hands out pinned host memory only, so everything on the device is scratch of this stream
                # (reused in stream order by the next batch): ONE shared set of device buffers, and buffer sets are cheap
                # (73 MB of pinned memory each).  Device-resident mode hands the device tensors out: one set per slot.
                sfx = "S" if (d2h and shared_device) else str(slot)
                hsfx = str(slot)
                E = int(offs[-1])
                he = host_edges.numpy()
                edges_dev = self._tensor("edges_dev" + sfx, (cap, 7), torch.float64)
                edges_dev[:max(E, 1)].copy_(host_edges[:max(E, 1)], non_blocking=True)
                h2d = torch.cuda.Event()
                h2d.record(stream)
                self._h2d_done[slot] = h2d
                graphs = [(he[offs[i]:offs[i] + n_art[i]], he[offs[i] + n_art[i]:offs[i + 1]]) for i in range(n)]   # views
                out = {"graphs": graphs, "stats": g["stats"], "offsets": offs, "n_art": n_art, "grow_device_ms": g["grow_ms"],
                       "edges_host": he[:E], "h2d_bytes": int(E * 56), "d2h_bytes": 0}
                if g.get("growth_stats") is not None:
                    out["growth_stats"] = g["growth_stats"]
                from . import _lib
                L = _lib.lib()
                if self.voxelize:
                    shape = tree2img.voxel_volume_shape(self.volume_dims)
                    vol = self._tensor("vol" + sfx, (n, *shape), torch.uint16)
                    # sized for the edge capacity, not for this batch: a growing workspace would mean a cudaMalloc (device-wide
                    # synchronisation) in the middle of the growth loops that are in flight
                    need = int(L.octa_voxelize_workspace_bytes(n, max(E, cap), _lib.int3(self.volume_dims)))
                    ws = self._tensor("vox_ws" + sfx, (need,), torch.uint8)
                    vol = tree2img.voxelize_batch_device(edges_dev[:max(E, 1)], offs, self.volume_dims, out=vol, workspace=ws)
                    out["volume"] = vol
                L.octa_raster2d_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
                L.octa_raster2d_workspace_bytes.restype = ctypes.c_size_t
                # label: ONE collection over all rows of the csv, as visualize_vessel_graphs.py:95 renders it
                lab = None
                if self.label_res is not None:
                    lab = self._tensor("label" + sfx, (n, self.label_res[1], self.label_res[0]), torch.uint8)
                    ws_l = self._tensor("r2d_ws_label" + sfx, (int(L.octa_raster2d_workspace_bytes(n, max(E, cap), self.label_res[1], self.label_res[0])),), torch.uint8)
                    tree2img.raster_batch_device(edges_dev, offs, self.label_res, self.mip_axis, out=lab, workspace=ws_l)
                # gray image: arterial and venous forest on separate canvases, np.maximum (generate_vessel_graph.py:80-85)
                img = self._tensor("image" + sfx, (n, self.image_res[1], self.image_res[0]), torch.uint8)
                ws_i = self._tensor("r2d_ws_image" + sfx, (int(L.octa_raster2d_workspace_bytes(n, max(E, cap), self.image_res[1], self.image_res[0])),), torch.uint8)
                tree2img.raster_batch_device(edges_dev, offs, self.image_res, self.mip_axis, out=img, workspace=ws_i, layer_split=n_art)
                out["label"], out["image"] = lab, img
                if d2h:
                    lab_h = None
                    if lab is not None:
                        lab_h = self._tensor("label_host" + hsfx, tuple(lab.shape), torch.uint8, pinned=True)
                        lab_h.copy_(lab, non_blocking=True)
                    img_h = self._tensor("image_host" + hsfx, tuple(img.shape), torch.uint8, pinned=True)
                    img_h.copy_(img, non_blocking=True)
                    out["_host"] = (lab_h, img_h)
                    out["d2h_bytes"] = int((lab_h.numel() if lab_h is not None else 0) + img_h.numel())
                    if d2h_volume and self.voxelize:
                        vol_h = self._tensor("vol_host" + hsfx, tuple(out["volume"].shape), torch.uint16, pinned=True)
                        vol_h.copy_(out["volume"], non_blocking=True)
                        out["_vol_host"] = vol_h
                        out["d2h_bytes"] += int(vol_h.numel() * 2)
                    if csv and self.device_csv:
                        # the files of the batch are written on the device (csrc/octa_csv_dev.cu) and come back as text; the host
                        # only slices them (graphs the device formatters decline are formatted by the host writer in _finish)
                        rows = max(E, cap)
                        ws_b, cap_b = graph_io.csv_device_sizes(n, rows)
                        text_dev = self._tensor("csv_text" + sfx, (cap_b,), torch.uint8)
                        toff_dev = self._tensor("csv_off" + sfx, (n + 1,), torch.int64)
                        fb_dev = self._tensor("csv_fb" + sfx, (n,), torch.int32)
                        ws_c = self._tensor("csv_ws" + sfx, (ws_b,), torch.uint8)
                        graph_io.csv_batch_device(edges_dev, offs, text_dev, toff_dev, fb_dev, ws_c, stream)
                        # (a row is 94 bytes; files that end beyond the copied part fall back to the host writer)
                        host_cap = min(cap_b, 104 * E + 64 * n)
                        self._tensor("csv_text_host" + hsfx, (min(cap_b, 104 * rows + 64 * n),), torch.uint8, pinned=True)   # sized once, for the capacity
                        text_h = self._tensor("csv_text_host" + hsfx, (host_cap,), torch.uint8, pinned=True)
                        toff_h = self._tensor("csv_off_host" + hsfx, (n + 1,), torch.int64, pinned=True)
                        fb_h = self._tensor("csv_fb_host" + hsfx, (n,), torch.int32, pinned=True)
                        text_h.copy_(text_dev[:host_cap], non_blocking=True)
                        toff_h.copy_(toff_dev, non_blocking=True)
                        fb_h.copy_(fb_dev, non_blocking=True)
                        out["_csv_dev"] = (text_h, toff_h, fb_h, he, offs, n)
                        out["d2h_bytes"] += int(host_cap)
                    elif csv:
                        if self._csv_pool is None:                                          # one pool for the pipeline's lifetime
                            self._csv_pool = cf.ThreadPoolExecutor(max_workers=self.host_threads)
                        # ctypes releases the GIL; the text is collected in _finish
                        out["_csv"] = [self._csv_pool.submit(graph_io.csv_bytes, he[offs[i]:offs[i + 1]]) for i in range(n)]
                if d2h and shared_device:             # the device tensors of this batch are recycled by the next one
                    for k in ("volume", "label", "image"):
                        out.pop(k, None)
                ready = torch.cuda.Event()
                ready.record(stream)
                out["ready"] = ready          # device results (and the pinned host copies) are complete once this event has fired
                out["trace"] = dict(g.get("trace", {}), t_post0=t_p0, t_post1=time.perf_counter())
                return out

    def _finish(self, out: dict, wait: bool = True) -> dict:
        """Hand a result out: wait for its device work (host-buffer mode) and collect the CSV text."""
        t_f0 = time.perf_counter()
        if "_host" in out:
            out["ready"].synchronize()
            lab_h, img_h = out.pop("_host")
            out["label_host"], out["image_host"] = (lab_h.numpy() if lab_h is not None else None), img_h.numpy()
            if "_vol_host" in out:
                out["volume_host"] = out.pop("_vol_host").numpy()
        elif wait:
            out["ready"].synchronize()
        t_r = time.perf_counter()
        if "_csv" in out:
            out["csv"] = [f.result() for f in out.pop("_csv")]
        if "_csv_dev" in out:
            text_h, toff_h, fb_h, he, offs, n = out.pop("_csv_dev")
            out["ready"].synchronize()
            t, o, fb = text_h.numpy(), toff_h.numpy(), fb_h.numpy()
            # views of the pinned text (bytes-like: ==, len, hashlib, file.write; valid as long as label_host / image_host are --
            # bytes(view) makes a copy that outlives the buffer set)
            out["csv"] = [graph_io.csv_bytes(he[offs[i]:offs[i + 1]]) if (fb[i] or o[i + 1] > len(t)) else t[o[i]:o[i + 1]].data
                          for i in range(n)]
            out["csv_host_fallbacks"] = int(sum(1 for i in range(n) if fb[i] or o[i + 1] > len(t)))
        if "trace" in out:
            out["trace"].update(t_fin0=t_f0, t_ready=t_r, t_fin1=time.perf_counter())
        return out

    def run(self, seeds: Sequence[int], d2h: bool = True, csv: bool = True, d2h_volume: bool = False) -> dict:
        """One step over len(seeds) samples.  Results: edges (host), offsets, volume / label / image (device
        tensors, complete), and with d2h: label_host / image_host (pinned uint8) and csv (list of bytes)."""
        return self._finish(self._post_stage(self._grow_stage(seeds, 0), 0, d2h, csv, d2h_volume=d2h_volume))

    @staticmethod
    def buffer_sets(in_flight: int, d2h: bool = False, extra_slots=None) -> int:
        """Buffer sets run_pipelined cycles through: one per loop in flight, one being post-processed, and in host-buffer mode
        twelve more (73 MB of pinned memory each) whose post-processing may lag behind: it runs at default priority in whatever
        the high-priority growth loops leave free, and a result is only handed out once its copies have landed -- measured on
        one B200, 8 loops in flight: 322 graphs/s with 2 spare sets, 405 with 6 (OCTA_EXTRA_SLOTS overrides).  Each set is
        allocated on first use: warm up with at least this many batches."""
        extra = extra_slots if extra_slots is not None else os.environ.get("OCTA_EXTRA_SLOTS")
        return max(1, int(in_flight)) + 1 + (max(0, int(extra)) if extra is not None else (12 if d2h else 1))

    def run_pipelined(self, seed_batches, d2h: bool = True, csv: bool = True, in_flight: int = 2, d2h_volume: bool = False,
                      extra_slots=None):
        """Generator over batches, results in order, software-pipelined.

        The growth loop is a chain of short latency-bound launches that leaves most of the GPU idle, voxelize / raster
        are throughput kernels, CSV text is host work: `in_flight` growth loops run side by side (one growth context,
        two high-priority streams and one host thread each) while a worker thread ENQUEUES the post-processing of finished
        batches on a second stream -- it never waits for the device.  A result is handed out once its `ready` event has fired
        (host-buffer mode) or right away with that event attached (d2h=False: wait on out["ready"] before reading the device
        tensors on another stream).  Every result equals what run() returns for the same seeds (a sample depends on its seed
        only); a yielded result stays valid until buffer_sets(in_flight, d2h) further batches have been started."""
        torch = self.torch
        in_flight = max(1, int(in_flight))
        nslots = self.buffer_sets(in_flight, d2h, extra_slots)
        if self._post_stream is None:
            with torch.cuda.device(self.device):
                self._post_stream = torch.cuda.Stream()
        pending = collections.deque()
        with cf.ThreadPoolExecutor(max_workers=in_flight) as growers, cf.ThreadPoolExecutor(max_workers=1) as poster:
            def post(gf, slot):
                return self._post_stage(gf.result(), slot, d2h, csv, self._post_stream, d2h_volume, shared_device=True)

            for k, seeds in enumerate(seed_batches):
                while len(pending) >= nslots:                 # buffer set k % nslots is free once result k - nslots is out
                    yield self._finish(pending.popleft().result(), wait=False)
                gf = growers.submit(self._grow_stage, seeds, k % nslots, k % in_flight)
                pending.append(poster.submit(post, gf, k % nslots))
            while pending:
                yield self._finish(pending.popleft().result(), wait=False)
            if not d2h:
                self._post_stream.synchronize()


def shard_seeds(base_seed: int, num_samples: int, rank: int, world: int):
    """Sample i (seed base+i) belongs to rank i mod world; the result of a sample depends on its seed only."""
    return [base_seed + i for i in range(num_samples) if i % world == rank]
