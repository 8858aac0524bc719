"""compute-sanitizer target: one small growth batch + voxelize + raster (keeps the run under a minute under the tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from octa_autosegmentation_b200 import growth, tree2img
from octa_autosegmentation_b200.config import default_config
cfg = default_config()
for m, i in zip(cfg["Greenhouse"]["modes"], (8, 8)):
    m["I"], m["N"] = i, 300
graphs, stats, _ = growth.grow_batch(cfg, [0, 1, 2])
e7 = np.concatenate(graphs[0])
vol = tree2img.voxelize_edges(e7, [96, 96, 8])
img = tree2img.raster_edges(e7, [128, 128])
print("ok", [len(a) + len(v) for a, v in graphs], int((vol > 0).sum()), int(img.max()))
