"""GPU probe for profilers: voxelize + 2-D rasters of 64 stored graphs (the 8 shipped csv fixtures, 8 copies each); no growth.
usage: post_only.py [reps] [what: vox|r1216|r304|all]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from conftest import GOLDEN, load_graph_rows, rows_to_edges7
from octa_autosegmentation_b200 import tree2img, _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
what = sys.argv[2] if len(sys.argv) > 2 else "all"
names = sorted(f for f in os.listdir(GOLDEN) if f.startswith("shipped_") and f.endswith(".csv.gz"))
gs = [rows_to_edges7(load_graph_rows(n)) for n in names]
graphs = [gs[i % len(gs)] for i in range(64)]
e7 = np.concatenate(graphs); offs = np.cumsum([0] + [len(g) for g in graphs])
dev = torch.from_numpy(e7).cuda()
def timeit(name, fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-28s %8.3f ms" % (name, e0.elapsed_time(e1) / reps), flush=True)
if what in ("vox", "all"):
    vol = torch.empty((64, *tree2img.voxel_volume_shape([1216, 1216, 16])), dtype=torch.uint16, device="cuda")
    timeit("voxelize [1216,1216,16]", lambda: tree2img.voxelize_batch_device(dev, offs, [1216, 1216, 16], out=vol))
if what in ("r1216", "all"):
    lab = torch.empty((64, 1216, 1216), dtype=torch.uint8, device="cuda")
    timeit("raster 1216^2", lambda: tree2img.raster_batch_device(dev, offs, [1216, 1216], out=lab))
if what in ("r304", "all"):
    img = torch.empty((64, 304, 304), dtype=torch.uint8, device="cuda")
    timeit("raster 304^2", lambda: tree2img.raster_batch_device(dev, offs, [304, 304], out=img))
