"""ctypes front-end of oracle/agg_oracle.c (TEST INFRASTRUCTURE ONLY): the CPU restatement of what
tree2img.rasterize_forest (tree2img.py:12-114) makes matplotlib's Agg backend compute, plus the host loop of
rasterize_forest (radius filter, legacy string rows, dropout: same code path as oracle/vox_oracle.voxelize_forest) and the
label step of visualize_vessel_graphs.py:95-101 (PIL convert("1") = Floyd-Steinberg)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libagg_oracle.so")


def build():
    subprocess.run(["make", "-C", _HERE, "--quiet"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "agg_oracle.c")):
            build()
        L = ctypes.CDLL(_LIB)
        L.agg_rasterize.restype = ctypes.c_long
        L.agg_rasterize.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
        L.agg_rasterize_segments.restype = ctypes.c_long
        L.agg_rasterize_segments.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_void_p]
        _lib = L
    return _lib


def raster_segments(segments: np.ndarray, linewidths: np.ndarray, H: int, W: int, variant: int = 0) -> np.ndarray:
    """One LineCollection: data-space segments [n, 4] (x0, y0, x1, y1) and line widths in points -> uint8 [H, W]."""
    segments = np.ascontiguousarray(segments, dtype=np.float64).reshape(-1, 4)
    linewidths = np.ascontiguousarray(linewidths, dtype=np.float64).reshape(-1)
    out = np.empty((int(H), int(W)), dtype=np.uint8)
    lib().agg_rasterize_segments(segments.ctypes.data, linewidths.ctypes.data, segments.shape[0], int(H), int(W), int(variant),
                                 out.ctypes.data)
    return out


def raster_edges(edges7: np.ndarray, image_resolution, MIP_axis: int = 2, min_radius: float = 0.0, max_radius: float = 1.0,
                 variant: int = 0) -> np.ndarray:
    """uint8 [H, W] gray image of the kept edges (tree2img.py:46-57,82-113)."""
    edges7 = np.ascontiguousarray(edges7, dtype=np.float64).reshape(-1, 7)
    W, H = int(image_resolution[0]), int(image_resolution[1])      # no_pixels_x, no_pixels_y (:48)
    axes = [a for a in (0, 1, 2) if a != MIP_axis]                 # :46
    out = np.empty((H, W), dtype=np.uint8)
    lib().agg_rasterize(edges7.ctypes.data, edges7.shape[0], H, W, axes[0], axes[1], float(min_radius), float(max_radius),
                        int(variant), out.ctypes.data)
    return out


def to_label(gray: np.ndarray) -> np.ndarray:
    """visualize_vessel_graphs.py:95-101 with --binarize: img[img<0.1]=0 (a no-op on integers), uint8, PIL convert("1")."""
    from PIL import Image
    return np.array(Image.fromarray(gray.astype(np.uint8)).convert("1"))


def rasterize_forest(forest, image_resolution, MIP_axis=2, radius_list=None, min_radius=0, max_radius=1, max_dropout_prob=0,
                     blackdict=None):
    """Restatement of tree2img.py:46-114 (host loop: radius filter :66-68, legacy string rows :70-76, subtree dropout
    :58-62,78-80, thickness :82-86) around the Agg restatement.  Same return value: (uint16 [H, W], blackdict)."""
    from random import random
    axes = [a for a in [0, 1, 2] if a != MIP_axis]
    if radius_list is None:
        radius_list = []
    no_pixels_x, no_pixels_y = image_resolution
    scale_factor = max(no_pixels_x, no_pixels_y)
    if blackdict is None:
        blackdict = dict()
        p = random() ** 10 * max_dropout_prob
    else:
        p = 0
    segs, lws = [], []
    for edge in forest:
        radius = float(edge["radius"])
        if radius < min_radius or radius > max_radius:
            continue
        if isinstance(edge["node1"], (np.ndarray, list)):
            current_node, proximal_node = tuple(edge["node1"]), tuple(edge["node2"])
        else:
            current_node = tuple([float(c) for c in edge["node1"][1:-1].split(" ") if len(c) > 0])
            proximal_node = tuple([float(c) for c in edge["node2"][1:-1].split(" ") if len(c) > 0])
        if proximal_node in blackdict or random() < p:
            blackdict[current_node] = True
            continue
        radius *= 1.3
        radius_list.append(radius)
        segs.append((current_node[axes[1]], current_node[axes[0]], proximal_node[axes[1]], proximal_node[axes[0]]))
        lws.append(radius * scale_factor)
    gray = raster_segments(np.array(segs, dtype=np.float64).reshape(-1, 4), np.array(lws, dtype=np.float64),
                           int(no_pixels_y), int(no_pixels_x))
    return gray.astype(np.uint16), blackdict
