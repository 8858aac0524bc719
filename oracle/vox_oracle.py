"""ctypes front-end of oracle/voxelize_oracle.c (TEST INFRASTRUCTURE ONLY) plus the host-side
edge-list preparation that mirrors tree2img.py:218-241 (string parsing, dropout, blackdict)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from random import random

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libvox_oracle.so")


def build():
    subprocess.run(["make", "-C", _HERE, "--quiet"], check=True)


def _lib():
    if not os.path.exists(_LIB):
        build()
    lib = ctypes.CDLL(_LIB)
    lib.vox_oracle.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                               ctypes.c_int, ctypes.c_void_p]
    lib.vox_oracle.restype = ctypes.c_int
    lib.vox_oracle_out_dims.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    return lib


def out_dims(dims):
    d = (ctypes.c_int * 3)(*[int(x) for x in dims])
    o = (ctypes.c_int * 3)()
    _lib().vox_oracle_out_dims(d, o)
    return tuple(o)


def voxelize_edges(edges7: np.ndarray, dims, min_radius=0.0, max_radius=1.0, ignore_z=False) -> np.ndarray:
    edges7 = np.ascontiguousarray(edges7, dtype=np.float64).reshape(-1, 7)
    D = out_dims(dims)
    out = np.empty(D, dtype=np.uint16)
    d = (ctypes.c_int * 3)(*[int(x) for x in dims])
    rc = _lib().vox_oracle(edges7.ctypes.data, edges7.shape[0], d, float(min_radius), float(max_radius),
                           int(bool(ignore_z)), out.ctypes.data)
    if rc != 0:
        raise MemoryError("vox_oracle failed")
    return out


def parse_node(s):
    """tree2img.py:235  -- legacy string form "[x y z]"."""
    return tuple([float(c) for c in s[1:-1].split(" ") if len(c) > 0])


def voxelize_forest(forest, volume_dimensions, radius_list=None, min_radius=0, max_radius=1, max_dropout_prob=0,
                    blackdict=None, ignore_z=False):
    """Same signature / return as tree2img.py:176-183; host loop restated from :218-241."""
    if radius_list is None:
        radius_list = []
    if blackdict is None:
        blackdict = dict()
        p = random() ** 10 * max_dropout_prob
    else:
        p = 0
    kept = []
    for edge in forest:
        radius = float(edge["radius"])
        if radius < min_radius or radius > max_radius:
            continue
        if isinstance(edge["node1"], (np.ndarray, list)):
            cur, prox = tuple(edge["node1"]), tuple(edge["node2"])
        else:
            cur, prox = parse_node(edge["node1"]), parse_node(edge["node2"])
        if prox in blackdict or random() < p:
            blackdict[cur] = True
            continue
        radius_list.append(radius)
        kept.append((*cur, *prox, radius))
    e7 = np.array(kept, dtype=np.float64).reshape(-1, 7)
    return voxelize_edges(e7, volume_dimensions, 0.0, 1.0, ignore_z), blackdict
