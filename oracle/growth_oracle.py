"""ctypes front-end of oracle/growth_oracle.cpp (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import csv
import ctypes
import io
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libgrowth_oracle.so")
WALLS = {"x0": 0, "x1": 1, "y0": 2, "y1": 3, "z0": 4, "z1": 5}


class OGMode(ctypes.Structure):
    _fields_ = [("I", ctypes.c_int), ("N", ctypes.c_int)] + \
        [(k, ctypes.c_double) for k in ("eps_n", "eps_s", "eps_k", "delta_art", "delta_ven", "gamma_art", "gamma_ven",
                                        "phi", "omega", "kappa", "delta_sigma")] + \
        [("reinit", ctypes.c_int), ("first_mode", ctypes.c_int)]


class OGConfig(ctypes.Structure):
    _fields_ = [("d", ctypes.c_double), ("r", ctypes.c_double), ("faz_bound0", ctypes.c_double),
                ("faz_bound1", ctypes.c_double), ("rotation_radius", ctypes.c_double),
                ("faz_center", ctypes.c_double * 2), ("nerve_center", ctypes.c_double * 2),
                ("nerve_radius", ctypes.c_double), ("param_scale", ctypes.c_double), ("size", ctypes.c_double * 3),
                ("n_modes", ctypes.c_int), ("modes", OGMode * 8), ("forest_type", ctypes.c_int),
                ("n_trees", ctypes.c_int), ("n_walls", ctypes.c_int), ("walls", ctypes.c_int * 6),
                ("ball_order", ctypes.c_int), ("venous", ctypes.c_int),
                ("geometry", ctypes.c_void_p), ("geom_dims", ctypes.c_int * 3)]


class OGStats(ctypes.Structure):
    _fields_ = [(k, ctypes.c_long) for k in ("n_art_nodes", "n_ven_nodes", "n_oxy_left", "n_co2_left", "py_draws",
                                             "np_u32", "nn_queries", "ball_queries", "bifurcations", "sprouts",
                                             "elongations", "walk_steps", "sum_A", "sum_M", "sum_P", "sum_S",
                                             "multi_balls", "reordered_balls", "interacting_groups", "kd_builds",
                                             "inter_evals", "inter_r1_changed", "max_dict", "max_list",
                                             "flag_iters_cons", "flag_iters_perm", "diff_iters", "perm_groups")]


EIG_HOOK = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                            ctypes.POINTER(ctypes.c_double))
TRACE_HOOK = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_long,
                              ctypes.c_long, ctypes.c_long)


def build():
    subprocess.run(["make", "-C", _HERE, "--quiet"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "growth_oracle.cpp")):
            build()
        L = ctypes.CDLL(_LIB)
        L.growth_oracle_run.restype = ctypes.c_long
        L.growth_oracle_run.argtypes = [ctypes.POINTER(OGConfig), ctypes.c_uint64, ctypes.c_void_p, ctypes.c_long,
                                        ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long),
                                        ctypes.POINTER(OGStats), ctypes.c_void_p, ctypes.c_void_p]
        L.og_last_sinks.restype = ctypes.c_long
        L.og_last_sinks.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_long]
        L.og_hash_tuple3.restype = ctypes.c_int64
        L.og_hash_tuple3.argtypes = [ctypes.c_void_p]
        L.og_set_order.restype = ctypes.c_long
        L.og_set_order.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        L.og_kd_indices.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        L.og_py_random.argtypes = [ctypes.c_uint64, ctypes.c_long, ctypes.c_void_p]
        L.og_np_stream.argtypes = [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_long, ctypes.c_uint32, ctypes.c_void_p,
                                   ctypes.c_long, ctypes.c_void_p]
        L.og_py_choice.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_long, ctypes.c_void_p]
        _lib = L
    return _lib


def make_config(config: dict, ball_order: int = 0, venous: bool = True) -> OGConfig:
    g, f = config["Greenhouse"], config["Forest"]
    c = OGConfig()
    c.d, c.r = float(g["d"]), float(g["r"])
    c.faz_bound0, c.faz_bound1 = [float(x) for x in g["FAZ_radius_bound"]]
    c.rotation_radius = float(g["rotation_radius"])
    c.faz_center[0], c.faz_center[1] = [float(x) for x in g["FAZ_center"]]
    c.nerve_center[0], c.nerve_center[1] = [float(x) for x in g["nerve_center"]]
    c.nerve_radius = float(g["nerve_radius"])
    c.param_scale = float(g["param_scale"])
    ss = g["SimulationSpace"]
    if ss.get("oxygen_sample_geometry_path") is not None:
        geo = np.ascontiguousarray(np.load(ss["oxygen_sample_geometry_path"]).astype(bool).astype(np.uint8))
        if geo.ndim != 3:
            raise ValueError("the sampling geometry must be a 3-D array (simulation_space.py:32 unpacks three sizes)")
        c._geometry_keepalive = geo
        c.geometry = geo.ctypes.data
        for k in range(3):
            c.geom_dims[k] = geo.shape[k]
            c.size[k] = geo.shape[k] / max(geo.shape)
    else:
        c.size[0], c.size[1], c.size[2] = float(ss["no_voxel_x"]), float(ss["no_voxel_y"]), float(ss["no_voxel_z"])
    modes = g["modes"]
    c.n_modes = len(modes)
    for i, m in enumerate(modes):
        om = c.modes[i]
        om.I, om.N = int(m["I"]), int(m["N"])
        for k in ("eps_n", "eps_s", "eps_k", "delta_art", "delta_ven", "gamma_art", "gamma_ven", "phi", "omega",
                  "kappa", "delta_sigma"):
            setattr(om, k, float(m[k]))
        om.reinit = int(m["name"] != modes[0]["name"])      # greenhouse.py:84
        om.first_mode = int(m == modes[0])                   # greenhouse.py:95
    c.forest_type = {"stumps": 0, "nerve": 1}[f["type"]]
    c.n_trees = int(f["N_trees"])
    walls = [WALLS[k] for k, v in f.get("source_walls", {}).items() if v]   # forest.py:81-84 (dict order)
    c.n_walls = len(walls)
    for i, w in enumerate(walls):
        c.walls[i] = w
    c.ball_order = ball_order
    c.venous = int(venous)
    return c


@EIG_HOOK
def _numpy_eig(cov9, w3, v9):
    """greenhouse.py:229 -- the same LAPACK dgeev the reference calls (numpy is test-side here)."""
    a = np.array([cov9[i] for i in range(9)]).reshape(3, 3)
    w, v = np.linalg.eig(a)
    k = int(np.argmax(w))
    order = [k] + [i for i in range(3) if i != k]
    # hand back real parts; column `k` is what the reference uses (np.real at :232-233)
    for i in range(3):
        w3[i] = float(np.real(w[i]))
        for j in range(3):
            v9[3 * j + i] = float(np.real(v[j, i]))
    # make argmax unambiguous for the C side when numpy returned complex eigenvalues
    if np.iscomplexobj(w):
        for i in range(3):
            w3[i] = 1.0 if i == k else 0.0


def run(config: dict, seed: int, ball_order: int = 0, use_numpy_eig: bool = True, trace=None):
    c = make_config(config, ball_order)
    cap = 40000
    st = OGStats()
    na, nv = ctypes.c_long(), ctypes.c_long()
    tr = TRACE_HOOK(trace) if trace else None
    while True:
        out = np.empty((cap, 7))
        n = lib().growth_oracle_run(ctypes.byref(c), int(seed), out.ctypes.data, cap, ctypes.byref(na), ctypes.byref(nv),
                                    ctypes.byref(st), ctypes.cast(_numpy_eig, ctypes.c_void_p) if use_numpy_eig else None,
                                    ctypes.cast(tr, ctypes.c_void_p) if tr else None)
        if n < 0:
            raise RuntimeError("growth_oracle_run failed")
        if n <= cap:
            break
        cap = int(n)
    stats = {k: getattr(st, k) for k, _ in OGStats._fields_}
    return out[:na.value].copy(), out[na.value:n].copy(), stats


def last_sinks():
    """(oxygen sinks [n, 3], CO2 sources [m, 3]) of this thread's last run(), in list order (greenhouse.py:403,412)."""
    out = []
    for which in (0, 1):
        n = lib().og_last_sinks(which, None, 0)
        a = np.empty((n, 3))
        lib().og_last_sinks(which, a.ctypes.data, n)
        out.append(a)
    return tuple(out)


def csv_bytes(edges7: np.ndarray) -> bytes:
    """generate_vessel_graph.py:59-66 with numpy itself doing the cell formatting."""
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow(["node1", "node2", "radius"])
    for row in edges7:
        w.writerow([row[0:3], row[3:6], float(row[6])])
    return buf.getvalue().encode()
