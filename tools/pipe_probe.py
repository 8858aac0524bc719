"""GPU probe: device-resident throughput of Pipeline.run_pipelined for the current environment switches
(OCTA_GROW_GRAPH, OCTA_BALL_ORDER, ...).  usage: pipe_probe.py [steps] [in_flight] [sub_batch] [d2h]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
from octa_autosegmentation_b200 import _lib
from octa_autosegmentation_b200.config import default_config
from octa_autosegmentation_b200.pipeline import Pipeline

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
SB = int(sys.argv[3]) if len(sys.argv) > 3 else 32
D2H = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
pipe = Pipeline(default_config(), volume_dims=(1216, 1216, 16), host_threads=max(1, (os.cpu_count() or 2) - 1))
seed = [1_000_000]
def batches(k):
    out = []
    for _ in range(k * 64 // SB):
        out.append(list(range(seed[0], seed[0] + SB))); seed[0] += SB
    return out
need = []
traces = []
def run(k):
    gm = []
    for out in pipe.run_pipelined(batches(k), d2h=D2H, csv=D2H, in_flight=L):
        gm.append(out["grow_device_ms"])
        traces.append(out["trace"])
        need.extend(s["replay_detail"][0] for s in out["stats"])
    torch.cuda.synchronize()
    return float(np.mean(gm))
run(max(3, (Pipeline.buffer_sets(L, D2H) * SB + 63) // 64 + 1))
need.clear(); traces.clear()
n0 = _lib.launch_count()
torch.cuda.synchronize(); t = time.time(); gm = run(K); dt = time.time() - t
print("PROBE graph=%s ball=%s in_flight=%d sub_batch=%d d2h=%d: %.1f graphs/s (%.1f ms per 64), loop %.0f ms, %d launches/step, kd-needed iterations/graph %.1f"
      % (os.environ.get("OCTA_GROW_GRAPH", "0"), os.environ.get("OCTA_BALL_ORDER", "ondemand"), L, SB, D2H, K * 64 / dt, dt / K * 1e3, gm,
         (_lib.launch_count() - n0) // K, float(np.mean(need)) if need else -1), flush=True)

# where the wall time of a batch goes (ms, means over the timed batches)
import collections
tr = traces
def m(f): return 1e3 * float(np.mean([f(t) for t in tr]))
by_ctx = collections.defaultdict(list)
for rec in tr: by_ctx[rec["ctx"]].append(rec)
gaps = [b["t_grow0"] - a["t_grow1"] for v in by_ctx.values() for a, b in zip(v, v[1:])]
print("TRACE grow stage %.0f (lock wait %.0f) | idle between loops of a context %.0f | post enqueue %.1f (waited for grow %.0f) | ready lag after enqueue %.0f | finish: wait ready %.1f, csv %.1f | t0 %.3f t1 %.3f"
      % (m(lambda t: t["t_grow1"] - t["t_grow0"]), m(lambda t: t["t_grow0"] - t["t_submit"]), 1e3 * float(np.mean(gaps)) if gaps else -1,
         m(lambda t: t["t_post1"] - t["t_post0"]), m(lambda t: t["t_post0"] - t["t_grow1"]), m(lambda t: t["t_ready"] - t["t_post1"]),
         m(lambda t: t["t_ready"] - t["t_fin0"]), m(lambda t: t["t_fin1"] - t["t_ready"]), 0.0, tr[-1]["t_fin1"] - tr[0]["t_submit"]), flush=True)
