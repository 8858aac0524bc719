"""Seeded harness around the UNMODIFIED reference modules (TEST INFRASTRUCTURE ONLY).

Runs only in the build container, where /root/reference exists; it is how the oracle
restatements in this directory and the fixtures under tests/golden/ are pinned to the real
reference.  Nothing in the product package, the `-m gpu` tests, smoke() or bench.py imports
this file (the GPU box has no /root/reference).

What it restates (glue only, no arithmetic): generate_vessel_graph.py:24-66 `main` --
Greenhouse -> Forest(arterial) -> Forest(venous) -> develop_forest -> edge list -> csv.writer.
`generate_vessel_graph.py` itself is not importable here (nibabel / utils.visualizer chain).

Seeding: the reference never seeds.  The harness seeds BOTH generators right before
`Greenhouse(...)`:  random.seed(s); np.random.seed(s)   (SURVEY.md Appendix A1).
"""
from __future__ import annotations

import copy
import csv
import io
import os
import random
import sys
import time

import numpy as np
import yaml

REFERENCE_ROOT = os.environ.get("OCTA_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vessel_graph_generation"))


def import_reference():
    """Put the shims (anytree, matplotlib stubs) and the reference root on sys.path."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import vessel_graph_generation.greenhouse as gh  # noqa: F401
    import vessel_graph_generation.forest as fo  # noqa: F401
    import vessel_graph_generation.tree2img as t2i  # noqa: F401
    return gh, fo, t2i


def load_config(path=None) -> dict:
    path = path or os.path.join(REFERENCE_ROOT, "docker", "vessel_graph_gen_docker_config.yml")
    with open(path) as f:
        return yaml.safe_load(f)


def small_config(I=(12, 12), N=400) -> dict:
    """The docker config with fewer iterations / candidates (seconds instead of minutes)."""
    cfg = load_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], I):
        m["I"] = i
        m["N"] = N
    return cfg


def run_growth(config: dict, seed: int, trace=None):
    """generate_vessel_graph.py:24-56 with seeding.  Returns (art_edges, ven_edges, greenhouse).

    `trace`, if given, is called as trace(greenhouse, tag) from inside develop_forest at the
    points the oracle restatement snapshots (used to bisect divergences)."""
    gh, fo, _ = import_reference()
    config = copy.deepcopy(config)
    random.seed(seed)
    np.random.seed(seed)
    greenhouse = gh.Greenhouse(config["Greenhouse"])
    art = fo.Forest(config["Forest"], greenhouse.d, greenhouse.r, greenhouse.simspace,
                    nerve_center=greenhouse.nerve_center, nerve_radius=greenhouse.nerve_radius)
    ven = fo.Forest(config["Forest"], greenhouse.d, greenhouse.r, greenhouse.simspace, arterial=False,
                    nerve_center=greenhouse.nerve_center, nerve_radius=greenhouse.nerve_radius)
    greenhouse.set_forests(art, ven)
    if trace is not None:
        _install_trace(greenhouse, trace)
    greenhouse.develop_forest()

    def edges(forest):
        return [{"node1": n.position, "node2": n.get_proximal_node().position, "radius": n.radius}
                for tree in forest.get_trees()
                for n in tree.get_tree_iterator(exclude_root=True, only_active=False)]

    return edges(art), edges(ven), greenhouse


def _install_trace(greenhouse, trace):
    """Wrap simulation_space_expansion (called once per iteration, greenhouse.py:125)."""
    orig = greenhouse.simulation_space_expansion

    def wrapped():
        trace(greenhouse, "iter_end")
        return orig()

    greenhouse.simulation_space_expansion = wrapped


def csv_bytes(art_edges, ven_edges) -> bytes:
    """generate_vessel_graph.py:59-66 (csv.writer default dialect -> CRLF)."""
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow(["node1", "node2", "radius"])
    for row in art_edges + ven_edges:
        w.writerow([row["node1"], row["node2"], row["radius"]])
    return buf.getvalue().encode()


def edges_to_array(edges) -> np.ndarray:
    """E x 7 float64: node1 xyz, node2 xyz, radius."""
    out = np.empty((len(edges), 7))
    for i, e in enumerate(edges):
        out[i, 0:3] = e["node1"]
        out[i, 3:6] = e["node2"]
        out[i, 6] = e["radius"]
    return out


def read_csv_rows(path):
    """The way every reference consumer reads a graph (visualize_vessel_graphs.py:72-75)."""
    with open(path, newline="") as f:
        return list(csv.DictReader(f))


def voxelize(forest, dims, **kw):
    _, _, t2i = import_reference()
    return t2i.voxelize_forest(forest, dims, **kw)


def rasterize(forest, image_resolution, MIP_axis=2, **kw):
    """The UNMODIFIED tree2img.rasterize_forest (tree2img.py:12-114).  matplotlib is absent from this image: the calls it
    makes land in oracle/shims/matplotlib, which renders the LineCollection with oracle/agg_oracle.c (the Agg restatement that
    reproduces the reference's 500 shipped labels bit for bit)."""
    _, _, t2i = import_reference()
    return t2i.rasterize_forest(forest, image_resolution, MIP_axis, **kw)


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    cfg = small_config() if a.small else load_config()
    t0 = time.time()
    art, ven, g = run_growth(cfg, a.seed)
    dt = time.time() - t0
    data = csv_bytes(art, ven)
    print("seed %d: %d art + %d ven edges, %.1f s, %d bytes" % (a.seed, len(art), len(ven), dt, len(data)))
    if a.out:
        with open(a.out, "wb") as f:
            f.write(data)


def nerve_config(I=(12, 12), N=600) -> dict:
    """The 12x12 mm^2 'nerve' variant of example_custom_vessel_simulation.ipynb (cell at :137-160), shortened."""
    cfg = load_config()
    g = cfg["Greenhouse"]
    g["param_scale"] = 12
    cfg["Forest"]["type"] = "nerve"
    cfg["output"]["image_scale_factor"] = 1216
    g["SimulationSpace"]["no_voxel_z"] = 0.0033
    g["d"] = 0.15
    for m, i in zip(g["modes"], I):
        m["I"], m["N"], m["delta_sigma"] = i, N, 0.002222
    cfg["Forest"]["N_trees"] = 16
    return cfg
