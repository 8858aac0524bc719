"""GPU probe: times octa_grow_batch_host on a batch of docker-config samples and checks a few of them
against the CPU oracle (exact cKDTree ball order AND list-index order)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from octa_autosegmentation_b200 import growth  # noqa: E402
from octa_autosegmentation_b200.config import default_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--check", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--stress", action="store_true")
a = ap.parse_args()
cfg = default_config()
if a.stress:
    for m in cfg["Greenhouse"]["modes"]:
        m["N"] = 8000
seeds = list(range(a.batch))
for rep in range(a.reps):
    t0 = time.time()
    graphs, stats, extra = growth.grow_batch(cfg, seeds, cap_edges=60000 if a.stress else 40000)
    dt = time.time() - t0
    ne = [len(g[0]) + len(g[1]) for g in graphs]
    print("rep %d: batch %d  device %.1f ms  wall %.2f s  -> %.1f graphs/s (device), edges mean %.0f min %d max %d" %
          (rep, a.batch, extra["device_ms"], dt, a.batch / (extra["device_ms"] * 1e-3), np.mean(ne), min(ne), max(ne)), flush=True)
    cyc = np.array([s["commit_cycles"] for s in stats], dtype=np.float64)
    print("   k_commit cycles per graph (mean | max over graphs) prologue/replay/refresh/active: %s | %s  [ms at 1.9 GHz: %s]" %
          (np.round(cyc.mean(0) / 1e6, 1), np.round(cyc.max(0) / 1e6, 1), np.round(cyc.max(0) / 1.9e6, 1)), flush=True)
    det = np.array([s["replay_detail"] for s in stats], dtype=np.float64).mean(0)
    print("   replay detail (mean/graph): Mcycles walk %.1f recheck %.1f | entries %d events %d walk steps %d re-evaluations %d records %d" %
          (det[1] / 1e6, det[2] / 1e6, det[3], det[4], det[5], det[6], det[7]), flush=True)
if a.check:
    from oracle import growth_oracle as go
    exact = idx = 0
    for s in seeds[:a.check]:
        e = np.concatenate(graphs[s])
        o0 = np.concatenate(go.run(cfg, s, ball_order=0)[:2])
        o1 = np.concatenate(go.run(cfg, s, ball_order=1)[:2])
        same0 = e.shape == o0.shape and np.array_equal(e[:, 6], o0[:, 6]) and np.abs(e[:, :6] - o0[:, :6]).max() < 1e-11
        same1 = e.shape == o1.shape and np.array_equal(e[:, 6], o1[:, 6]) and np.abs(e[:, :6] - o1[:, :6]).max() < 1e-11
        exact += same0
        idx += same1
        print("seed %d: %d edges; == oracle(kd order) %s; == oracle(index order) %s" % (s, len(e), same0, same1), flush=True)
    print("parity: %d/%d vs exact-order oracle, %d/%d vs index-order oracle" % (exact, a.check, idx, a.check))
