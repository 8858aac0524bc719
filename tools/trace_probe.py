"""GPU probe: kernel timeline (CUPTI through torch.profiler) of the pipelined path in steady state; writes
gpurun_out/trace_kernels.csv.gz (name, stream, start_us, dur_us) for offline analysis.
usage: trace_probe.py [steps] [in_flight] [sub_batch]"""
import gzip, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from octa_autosegmentation_b200.config import default_config
from octa_autosegmentation_b200.pipeline import Pipeline

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
L = int(sys.argv[2]) if len(sys.argv) > 2 else 7
SB = int(sys.argv[3]) if len(sys.argv) > 3 else 64
pipe = Pipeline(default_config(), volume_dims=[1216, 1216, 16])
seed = [1_000_000]
def batches(k):
    out = []
    for _ in range(k):
        out.append(list(range(seed[0], seed[0] + SB))); seed[0] += SB
    return out
def run(k):
    for out in pipe.run_pipelined(batches(k), d2h=True, csv=True, in_flight=L):
        pass
    torch.cuda.synchronize()
run(Pipeline.buffer_sets(L, True) + 1)
t = time.time()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(K)
dt = time.time() - t
print("traced %d batches of %d in %.2f s (%.1f graphs/s under the profiler)" % (K, SB, dt, K * SB / dt), flush=True)
rows = []
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    rows.append((ev.name.split("(")[0][-48:], ev.device_index, getattr(ev, "stream", -1) if hasattr(ev, "stream") else -1,
                 ev.time_range.start, ev.time_range.end - ev.time_range.start))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with gzip.open(os.path.join(ROOT, "gpurun_out", "trace_kernels.csv.gz"), "wt") as f:
    for r in rows:
        f.write("%s,%s,%s,%.3f,%.3f\n" % r)
print("events", len(rows))
prof.export_chrome_trace(os.path.join(ROOT, "gpurun_out", "trace_chrome.json"))
