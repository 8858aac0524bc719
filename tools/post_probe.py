"""GPU probe: CUDA-event times of the post-processing kernels (voxelize, 2-D rasters) on one batch of grown graphs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from octa_autosegmentation_b200 import growth, tree2img, _lib
from octa_autosegmentation_b200.config import default_config

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
graphs, stats, extra = growth.grow_batch(default_config(), list(range(100, 100 + B)))
e7 = np.concatenate([np.concatenate(g) for g in graphs])
offs = np.cumsum([0] + [len(g[0]) + len(g[1]) for g in graphs])
split = np.array([len(g[0]) for g in graphs], dtype=np.int64)
dev = torch.from_numpy(e7).cuda()
print("batch %d, %d edges, growth %.0f ms" % (B, len(e7), extra["device_ms"]))
def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print("%-34s %8.3f ms  (%d launches)" % (name, e0.elapsed_time(e1) / reps, (_lib.launch_count() - n0) // reps), flush=True)
vol = torch.empty((B, *tree2img.voxel_volume_shape([1216, 1216, 16])), dtype=torch.uint16, device="cuda")
timeit("voxelize [1216,1216,16]", lambda: tree2img.voxelize_batch_device(dev, offs, [1216, 1216, 16], out=vol))
lab = torch.empty((B, 1216, 1216), dtype=torch.uint8, device="cuda")
timeit("raster 1216^2", lambda: tree2img.raster_batch_device(dev, offs, [1216, 1216], out=lab))
img = torch.empty((B, 304, 304), dtype=torch.uint8, device="cuda")
timeit("raster 304^2", lambda: tree2img.raster_batch_device(dev, offs, [304, 304], out=img))
timeit("raster 304^2, art/ven layers", lambda: tree2img.raster_batch_device(dev, offs, [304, 304], out=img, layer_split=split))
