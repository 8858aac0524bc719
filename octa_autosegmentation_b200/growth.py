"""Host-side mirror of the reference's growth entry (generate_vessel_graph.py:24-56) over the C ABI:
YAML-style config dict + per-sample seeds in, arterial / venous edge tables out.

Seeding contract: sample i behaves exactly like the unmodified reference run with
`random.seed(seeds[i]); np.random.seed(seeds[i])` issued right before `Greenhouse(...)`."""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np

from . import _lib

WALLS = {"x0": 0, "x1": 1, "y0": 2, "y1": 3, "z0": 4, "z1": 5}


class OctaGrowMode(ctypes.Structure):
    _fields_ = [("I", ctypes.c_int32), ("N", ctypes.c_int32)] + \
        [(k, ctypes.c_double) for k in ("eps_n", "eps_s", "eps_k", "delta_art", "delta_ven", "gamma_art", "gamma_ven",
                                        "phi", "omega", "kappa", "delta_sigma")] + \
        [("reinit", ctypes.c_int32), ("first_mode", ctypes.c_int32)]


class OctaGrowConfig(ctypes.Structure):
    _fields_ = [("d", ctypes.c_double), ("r", ctypes.c_double), ("faz_radius_bound", ctypes.c_double * 2),
                ("rotation_radius", ctypes.c_double), ("faz_center", ctypes.c_double * 2),
                ("nerve_center", ctypes.c_double * 2), ("nerve_radius", ctypes.c_double),
                ("param_scale", ctypes.c_double), ("size", ctypes.c_double * 3), ("n_modes", ctypes.c_int32),
                ("modes", OctaGrowMode * 8), ("forest_type", ctypes.c_int32), ("n_trees", ctypes.c_int32),
                ("n_walls", ctypes.c_int32), ("walls", ctypes.c_int32 * 6), ("cap_nodes", ctypes.c_int32),
                ("cap_sinks", ctypes.c_int32), ("geometry", ctypes.c_void_p), ("geom_dims", ctypes.c_int32 * 3)]


class OctaGrowStats(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int64) for k in ("n_art_nodes", "n_ven_nodes", "n_oxy_left", "n_co2_left", "py_draws",
                                              "sum_A", "sum_M", "sum_P", "sum_S")] + \
        [("commit_cycles", ctypes.c_int64 * 4), ("replay_detail", ctypes.c_int64 * 8), ("err", ctypes.c_int32), ("n_iters", ctypes.c_int32)]


def simspace_shape(config: dict) -> np.ndarray:
    """greenhouse.simspace.shape (simulation_space.py:29-34): the normalised shape of the sampling geometry when
    SimulationSpace.oxygen_sample_geometry_path is set, else (no_voxel_x, no_voxel_y, no_voxel_z)."""
    ss = config["Greenhouse"]["SimulationSpace"]
    if ss.get("oxygen_sample_geometry_path") is not None:
        shape = np.array(np.load(ss["oxygen_sample_geometry_path"], mmap_mode="r").shape, dtype=np.float64)
        return shape / shape.max()
    return np.array([float(ss["no_voxel_x"]), float(ss["no_voxel_y"]), float(ss["no_voxel_z"])])


def make_config(config: dict, cap_nodes: int = 0, cap_sinks: int = 0) -> OctaGrowConfig:
    g, f = config["Greenhouse"], config["Forest"]
    ss = g["SimulationSpace"]
    modes = g["modes"]
    if not 1 <= len(modes) <= 8:
        raise ValueError("between 1 and 8 growth modes are supported")
    c = OctaGrowConfig()
    c.d, c.r = float(g["d"]), float(g["r"])
    c.faz_radius_bound[0], c.faz_radius_bound[1] = [float(x) for x in g["FAZ_radius_bound"]]
    c.rotation_radius = float(g["rotation_radius"])
    c.faz_center[0], c.faz_center[1] = [float(x) for x in g["FAZ_center"]]
    c.nerve_center[0], c.nerve_center[1] = [float(x) for x in g["nerve_center"]]
    c.nerve_radius = float(g["nerve_radius"])
    c.param_scale = float(g["param_scale"])
    if ss.get("oxygen_sample_geometry_path") is not None:
        # simulation_space.py:29-34: the mask replaces no_voxel_x/y/z (the library copies it at octa_grow_create)
        geo = np.ascontiguousarray(np.load(ss["oxygen_sample_geometry_path"]).astype(bool).astype(np.uint8))
        if geo.ndim != 3:
            raise ValueError("oxygen_sample_geometry_path: the array must be 3-D (simulation_space.py:32 unpacks three sizes)")
        c._geometry_keepalive = geo
        c.geometry = geo.ctypes.data
        for k in range(3):
            c.geom_dims[k] = geo.shape[k]
            c.size[k] = geo.shape[k] / max(geo.shape)
    else:
        c.size[0], c.size[1], c.size[2] = float(ss["no_voxel_x"]), float(ss["no_voxel_y"]), float(ss["no_voxel_z"])
    c.n_modes = len(modes)
    for i, m in enumerate(modes):
        om = c.modes[i]
        om.I, om.N = int(m["I"]), int(m["N"])
        for k in ("eps_n", "eps_s", "eps_k", "delta_art", "delta_ven", "gamma_art", "gamma_ven", "phi", "omega",
                  "kappa", "delta_sigma"):
            setattr(om, k, float(m[k]))
        om.reinit = int(m["name"] != modes[0]["name"])
        om.first_mode = int(m == modes[0])
    if f["type"] not in ("stumps", "nerve"):
        # forest.py:36
        raise NotImplementedError("The Forest initialization type '%s' is not implemented. Try 'stump' or 'nerve' instead." % f["type"])
    c.forest_type = 0 if f["type"] == "stumps" else 1
    c.n_trees = int(f["N_trees"])
    walls = [WALLS[k] for k, v in f.get("source_walls", {}).items() if v]
    c.n_walls = len(walls)
    for i, w in enumerate(walls):
        c.walls[i] = w
    c.cap_nodes, c.cap_sinks = int(cap_nodes), int(cap_sinks)
    return c


def _bind():
    L = _lib.lib()
    run_args = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    L.octa_grow_batch_host.argtypes = [ctypes.POINTER(OctaGrowConfig)] + run_args
    L.octa_grow_create.argtypes = [ctypes.POINTER(OctaGrowConfig), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.octa_grow_run.argtypes = [ctypes.c_void_p] + run_args
    L.octa_grow_run_packed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_double)]
    L.octa_grow_sinks.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                  ctypes.POINTER(ctypes.c_int64)]
    L.octa_grow_destroy.argtypes = [ctypes.c_void_p]
    L.octa_grow_destroy.restype = None
    return L


class GrowContext:
    """Persistent device state for batches of up to `max_graphs` samples of one config (octa_grow_create/run/destroy)."""

    def __init__(self, config: dict, max_graphs: int, cap_edges: int = 40000, cap_nodes: int = 0, cap_sinks: int = 0):
        self.L = _bind()
        self.max_graphs = int(max_graphs)
        self.cap_edges = int(cap_edges)
        self._cfg = make_config(config, cap_nodes, cap_sinks)
        self._h = ctypes.c_void_p()
        _lib.check(self.L.octa_grow_create(ctypes.byref(self._cfg), self.max_graphs, ctypes.byref(self._h)))
        self._out = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.L.octa_grow_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_packed(self, seeds: Sequence[int], out: np.ndarray, trace: np.ndarray = None):
        """Rows of all graphs packed back to back into `out` (float64 [cap, 7], e.g. a view of pinned memory).
        Returns (offsets int64 [n+1], n_art int64 [n], stats, device_ms).  `trace` (optional, int32 [n, 4096, 4], C order)
        receives the per-iteration counts (arterial nodes, O2 sinks, venous nodes, CO2 sources)."""
        n = len(seeds)
        if n > self.max_graphs:
            raise ValueError("batch larger than the context")
        sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        offs = np.zeros(n + 1, dtype=np.int64)
        na = np.zeros(n, dtype=np.int64)
        nv = np.zeros(n, dtype=np.int64)
        st = (OctaGrowStats * n)()
        ms = ctypes.c_double(0)
        rc = self.L.octa_grow_run_packed(self._h, sd.ctypes.data, n, out.ctypes.data, out.shape[0], offs.ctypes.data,
                                         na.ctypes.data, nv.ctypes.data, ctypes.cast(st, ctypes.c_void_p),
                                         trace.ctypes.data if trace is not None else None, ctypes.byref(ms))
        stats = [{k: (list(getattr(st[i], k)) if k in ('commit_cycles', 'replay_detail') else getattr(st[i], k)) for k, _ in OctaGrowStats._fields_} for i in range(n)]
        if rc != 0:
            err = _lib.OctaError(rc, self.L.octa_last_error().decode(errors="replace"))
            err.stats = stats
            raise err
        return offs, na, stats, ms.value

    def sinks(self, graph: int):
        """(oxygen sinks [n, 3], CO2 sources [m, 3]) graph `graph` of the LAST run ended with, in list order: what
        Greenhouse.save_stats scatters (greenhouse.py:403,412; octa_grow_sinks)."""
        out = []
        for which in (0, 1):
            n = ctypes.c_int64(0)
            rc = self.L.octa_grow_sinks(self._h, int(graph), which, None, 0, ctypes.byref(n))
            if rc not in (0, _lib.OCTA_E_NOMEM):
                _lib.check(rc)
            a = np.empty((n.value, 3))
            if n.value:
                _lib.check(self.L.octa_grow_sinks(self._h, int(graph), which, a.ctypes.data, n.value, ctypes.byref(n)))
            out.append(a)
        return tuple(out)

    def run(self, seeds: Sequence[int], trace: bool = False, copy: bool = True):
        n = len(seeds)
        if n > self.max_graphs:
            raise ValueError("batch larger than the context")
        sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        na = np.zeros(n, dtype=np.int64)
        nv = np.zeros(n, dtype=np.int64)
        st = (OctaGrowStats * n)()
        tr = np.zeros((n, 4096, 4), dtype=np.int32) if trace else None
        ms = ctypes.c_double(0)
        if self._out is None:
            self._out = np.empty((self.max_graphs, self.cap_edges, 7), dtype=np.float64)
        rc = self.L.octa_grow_run(self._h, sd.ctypes.data, n, self._out.ctypes.data, self.cap_edges, na.ctypes.data,
                                  nv.ctypes.data, ctypes.cast(st, ctypes.c_void_p), tr.ctypes.data if trace else None,
                                  ctypes.byref(ms))
        stats = [{k: (list(getattr(st[i], k)) if k in ('commit_cycles', 'replay_detail') else getattr(st[i], k)) for k, _ in OctaGrowStats._fields_} for i in range(n)]
        if rc != 0:
            err = _lib.OctaError(rc, self.L.octa_last_error().decode(errors="replace"))
            err.stats = stats
            raise err
        out = self._out
        if copy:
            graphs = [(out[i, :na[i]].copy(), out[i, na[i]:na[i] + nv[i]].copy()) for i in range(n)]
        else:   # views into the context's buffer, valid until the next run
            graphs = [(out[i, :na[i]], out[i, na[i]:na[i] + nv[i]]) for i in range(n)]
        n_it = stats[0]["n_iters"] if n else 0
        return graphs, stats, {"device_ms": ms.value, "trace": tr[:, :n_it] if trace else None}


def grow_batch(config: dict, seeds: Sequence[int], cap_edges: int = 40000, trace: bool = False, cap_nodes: int = 0,
               cap_sinks: int = 0):
    """Grow len(seeds) independent samples on the GPU (one-shot context).

    Returns (graphs, stats, extra): graphs[i] = (art_edges7, ven_edges7) float64 arrays in the reference's row order;
    stats[i] = dict of per-sample counters; extra = {"device_ms": ..., "trace": int32 [n, iters, 4] or None}."""
    ctx = GrowContext(config, len(seeds), cap_edges, cap_nodes, cap_sinks)
    try:
        return ctx.run(seeds, trace=trace)
    finally:
        ctx.close()
