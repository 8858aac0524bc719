"""GPU probe (diagnostics): how often k_kill asks for the exact cKDTree ball order (graph-iterations per graph)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from octa_autosegmentation_b200 import growth
from octa_autosegmentation_b200.config import default_config
cfg = default_config()
for rep in range(2):
    graphs, stats, extra = growth.grow_batch(cfg, list(range(32)), cap_edges=40000)
    print("exact order needed in %s iterations of %d; device ms %.1f" % ([int(s["replay_detail"][0]) for s in stats], stats[0]["n_iters"], extra["device_ms"]))
