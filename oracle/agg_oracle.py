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
        _lib = L
    return _lib


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
