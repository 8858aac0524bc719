"""Greenhouse of the reference (greenhouse.py:15-441) over the GPU engine: same constructor attributes, `set_forests`,
`develop_forest`, `save_stats`; the growth loop of :58-147 is ONE run of `growth.GrowContext` (csrc/octa_grow_*.cu)."""
from __future__ import annotations

import copy
import random

import numpy as np

from .. import growth, stats_plots
from .forest import Forest
from .simulation_space import SimulationSpace


class _SinkList:
    """What `oxy_mesh` / `co2_mesh` are asked for after the growth: `.get_all_elements()` (greenhouse.py:403,412)."""

    def __init__(self, xyz: np.ndarray):
        self._xyz = xyz

    def get_all_elements(self):
        return list(self._xyz)


class Greenhouse:
    def __init__(self, config: dict, seed: int = None):
        """config: the `Greenhouse` block of the YAML (greenhouse.py:17-32).  seed: the graph equals the reference's after
        `random.seed(seed); np.random.seed(seed)` right before `Greenhouse(...)`; None draws one from Python's `random`."""
        self.config = config
        self.modes = config["modes"]
        self.seed = random.getrandbits(32) if seed is None else int(seed)
        self.sigma_t = 1
        self.param_scale = config["param_scale"]
        self.d = config["d"] / self.param_scale
        self.r = config["r"] / self.param_scale
        # the first draw of the seeded numpy stream (greenhouse.py:25): the engine draws the same number on its side
        self.FAZ_radius = np.random.RandomState(self.seed).normal(config["FAZ_radius_bound"][0] / self.param_scale,
                                                                  config["FAZ_radius_bound"][1] / self.param_scale)
        self.rotation_radius = config["rotation_radius"] / self.param_scale
        self.FAZ_center = config["FAZ_center"]
        self.nerve_center = np.array(config["nerve_center"]) / self.param_scale
        self.nerve_radius = np.array(config["nerve_radius"]) / self.param_scale
        self.simspace = SimulationSpace(config["SimulationSpace"], self.FAZ_center, self.FAZ_radius,
                                        nerve_center=self.nerve_center, nerve_radius=self.nerve_radius)
        self.arterial_forest = self.venous_forest = None
        self.art_nodes_per_step, self.oxys_per_step, self.ven_nodes_per_step, self.co2_per_step = [0], [0], [0], [0]
        self.time_per_step = []
        self.oxy_mesh, self.co2_mesh = _SinkList(np.zeros((0, 3))), _SinkList(np.zeros((0, 3)))
        self.stats = None

    def set_forests(self, arterialForest: Forest, venousForest: Forest = None):
        self.arterial_forest = arterialForest
        self.venous_forest = venousForest

    def develop_forest(self):
        """The main loop (greenhouse.py:58-147) on the GPU; fills both forests, the per-step lists and the sink lists."""
        art, ven = self.arterial_forest, self.venous_forest
        if art is None:
            raise RuntimeError("set_forests() first")
        if ven is None:
            raise NotImplementedError("the engine grows the arterial and the venous forest together; pass both to set_forests()")
        if art.config != ven.config:
            raise ValueError("both forests must be built from the same Forest config block")
        cfg = {"Greenhouse": copy.deepcopy(self.config), "Forest": copy.deepcopy(art.config)}
        ctx = growth.GrowContext(cfg, 1)
        try:
            graphs, stats, extra = ctx.run([self.seed], trace=True)
            oxys, co2s = ctx.sinks(0)
        finally:
            ctx.close()
        art._fill(graphs[0][0])
        ven._fill(graphs[0][1])
        tr = np.asarray(extra["trace"][0]).reshape(-1, 4)
        self.art_nodes_per_step = [0] + tr[:, 0].tolist()
        self.oxys_per_step = [0] + tr[:, 1].tolist()
        self.ven_nodes_per_step = [0] + tr[:, 2].tolist()
        self.co2_per_step = [0] + tr[:, 3].tolist()
        n_it = max(1, len(tr))
        self.time_per_step = [extra["device_ms"] * 1e-3 / n_it] * len(tr)          # the loop's device time, evenly spread
        self.oxy_mesh, self.co2_mesh = _SinkList(oxys), _SinkList(co2s)
        self.stats = stats[0]

    def save_stats(self, out_dir: str):
        """greenhouse.py:401-441 (PIL plots of the same data: stats_plots.py)."""
        ps = np.stack([self.art_nodes_per_step, self.oxys_per_step, self.ven_nodes_per_step, self.co2_per_step], 1)[1:]
        stats_plots.save_stats(out_dir, np.array(self.oxy_mesh.get_all_elements()).reshape(-1, 3),
                               np.array(self.co2_mesh.get_all_elements()).reshape(-1, 3) if self.venous_forest is not None else None,
                               ps, self.time_per_step)
