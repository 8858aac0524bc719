"""GPU probe: wall-clock breakdown of one pipeline step (host + device phases)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from octa_autosegmentation_b200 import graph_io, growth, tree2img, _lib
from octa_autosegmentation_b200.config import default_config
from octa_autosegmentation_b200.pipeline import Pipeline

B = 64
pipe = Pipeline(default_config())
for w in range(2):
    pipe.run(list(range(1000 + w * B, 1000 + (w + 1) * B)), d2h=True, csv=True)
torch.cuda.synchronize()
# manual replay of Pipeline.run with timers
seeds = list(range(5000, 5000 + B))
t = [time.perf_counter()]
def lap(name):
    torch.cuda.synchronize(); t.append(time.perf_counter()); print("%-28s %7.1f ms" % (name, (t[-1] - t[-2]) * 1e3), flush=True)
cap = B * pipe.edge_cap
host_edges = pipe._tensor("edges_host", (cap, 7), torch.float64, pinned=True)
offs, n_art, stats, grow_ms = pipe._grows[0].run_packed(seeds, host_edges.numpy()); lap("grow.run_packed (device loop %.1f)" % grow_ms)
E = int(offs[-1]); he = host_edges.numpy()
edges_dev = torch.empty((E, 7), dtype=torch.float64, device="cuda"); edges_dev.copy_(host_edges[:E], non_blocking=True); lap("H2D edges")
vol = tree2img.voxelize_batch_device(edges_dev, offs, [1216, 1216, 16], out=pipe._buf["vol"][:B * 1216 * 1216 * 53].view(B, 1216, 1216, 53)); lap("voxelize")
lab = tree2img.raster_batch_device(edges_dev[:E], offs, [1216, 1216]); lap("raster 1216")
img = tree2img.raster_batch_device(edges_dev[:E], offs, [304, 304]); lap("raster 304")
lab_h = torch.empty(lab.shape, dtype=torch.uint8).pin_memory(); lap("pin alloc")
lab_h.copy_(lab); lap("D2H label")
import concurrent.futures as cf
with cf.ThreadPoolExecutor(max_workers=15) as ex:
    out = list(ex.map(lambda i: graph_io.csv_bytes(he[offs[i]:offs[i + 1]]), range(B)))
lap("csv 64 graphs, 15 threads")
one = graph_io.csv_bytes(he[offs[0]:offs[1]]); lap("csv 1 graph")
