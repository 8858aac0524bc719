"""Config-space fuzzer of the growth oracle against the UNMODIFIED reference (build container only; test infrastructure).

    python oracle/fuzz_vs_reference.py [rng_seed] [cases] [max_iterations_per_mode]

Draws configs around the shipped docker config and the 12x12 mm^2 nerve variant (step length, radii, FAZ, param_scale, slab
thickness, non-square spaces, 1-3 modes incl. repeated mode names, every per-mode parameter, tree counts, source-wall sets,
32-bit seeds), grows each with the reference (oracle/ref_harness.py) and with oracle/growth_oracle.cpp, and compares the CSV bytes.
Diverging cases are dumped to /tmp/fuzz_bad_<case>.json.  Results on record: DESIGN.md section 2."""
import copy
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh, growth_oracle as go
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 5)
bad = 0
MAX_I = int(sys.argv[3]) if len(sys.argv) > 3 else 16
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    nerve = rng.random() < 0.3
    cfg = rh.nerve_config() if nerve else rh.load_config()
    g = cfg["Greenhouse"]
    def jig(v, lo=0.7, hi=1.4): return float(v * rng.uniform(lo, hi))
    g["d"] = jig(g["d"]); g["r"] = jig(g["r"])
    g["FAZ_radius_bound"] = [jig(g["FAZ_radius_bound"][0]), jig(g["FAZ_radius_bound"][1])]
    g["FAZ_center"] = [float(rng.uniform(0.3, 0.7)), float(rng.uniform(0.3, 0.7))]
    g["rotation_radius"] = jig(g["rotation_radius"])
    if not nerve:
        g["param_scale"] = float(rng.choice([2, 3, 3, 4, 6]))
        g["SimulationSpace"]["no_voxel_z"] = float(rng.choice([0.0131, 0.0131, 0.03, 0.1]))
        if rng.random() < 0.3:
            g["SimulationSpace"]["no_voxel_x"] = float(rng.choice([1, 0.8, 0.5]))
        elif rng.random() < 0.3:
            g["SimulationSpace"]["no_voxel_y"] = float(rng.choice([1, 0.75, 0.6]))
    nm = int(rng.choice([1, 2, 2, 3]))
    modes = g["modes"]
    while len(modes) > nm: modes.pop()
    while len(modes) < nm:
        m = copy.deepcopy(modes[-1]); m["name"] = rng.choice(["SVC", "DVC", "X%d" % len(modes)]).item(); modes.append(m)
    for m in modes:
        m["I"] = int(rng.integers(4, MAX_I)); m["N"] = int(rng.integers(150, 700))
        for k in ("eps_n", "eps_s", "eps_k", "delta_art", "delta_ven"): m[k] = jig(m[k], 0.8, 1.3)
        m["gamma_art"] = float(rng.uniform(20, 110)); m["gamma_ven"] = float(rng.uniform(20, 110))
        m["phi"] = float(rng.uniform(5, 40)); m["omega"] = float(rng.choice([0, 0.3, 0.7, 1.0])); m["kappa"] = float(rng.uniform(2.0, 4.0))
        m["delta_sigma"] = jig(m["delta_sigma"], 0.3, 2.0)
    cfg["Forest"]["N_trees"] = int(rng.integers(1, 14))
    if not nerve:
        walls = {k: bool(rng.random() > 0.4) for k in ("x0", "x1", "y0", "y1")}
        if not any(walls.values()): walls["y1"] = True
        walls.update(z0=False, z1=False)
        cfg["Forest"]["source_walls"] = walls
    seed = int(rng.integers(0, 2**31))
    t = time.time()
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            art, ven, gh = rh.run_growth(cfg, seed)
    except Exception as e:
        print(case, "reference raised", type(e).__name__, str(e)[:100], flush=True); continue
    b = rh.csv_bytes(art, ven)
    try:
        a2, v2, st = go.run(cfg, seed)
    except Exception as e:
        print(case, "ORACLE raised", type(e).__name__, str(e)[:100]); bad += 1
        json.dump({"cfg": cfg, "seed": seed}, open('/tmp/fuzz_bad_%d.json' % case, 'w')); continue
    ok = b == go.csv_bytes(np.concatenate([a2, v2]))
    if not ok:
        bad += 1
        json.dump({"cfg": cfg, "seed": seed}, open('/tmp/fuzz_bad_%d.json' % case, 'w'))
    print(case, "nerve" if nerve else "stumps", nm, len(art), len(ven), ok, round(time.time() - t, 1), flush=True)
print("BAD", bad)
