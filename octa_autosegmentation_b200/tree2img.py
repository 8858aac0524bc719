"""Drop-in for the reference's vessel_graph_generation/tree2img.py raster entry points, running on
hand-written sm_100a kernels behind the C ABI (include/octa_b200.h).

`voxelize_forest` keeps the reference signature and return value (tree2img.py:176-183,:279-280):
list-of-dict `forest` with ndarray/list or legacy "[x y z]" string nodes in, (uint16 ndarray,
blackdict) out; it mutates the caller's `radius_list` / `blackdict` exactly like the reference and
draws from Python's global `random` in the same order (one draw for p, one per surviving edge),
so a seeded caller sees the same stream position afterwards.
"""
from __future__ import annotations

import ctypes
from random import random
from typing import Sequence

import numpy as np

from . import _lib


def _parse_node(s: str):
    # tree2img.py:235 (legacy string rows written by csv.writer from str(ndarray))
    return tuple([float(c) for c in s[1:-1].split(" ") if len(c) > 0])


def forest_to_edges7(forest, radius_list, min_radius, max_radius, max_dropout_prob, blackdict):
    """Host side of tree2img.py:218-241: radius filter, node parsing, subtree dropout.  Returns the
    kept edges as E x 7 float64 (node1 xyz, node2 xyz, UNSCALED radius) and the blackdict."""
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
        n1 = edge["node1"]
        if isinstance(n1, (np.ndarray, list)):
            cur, prox = tuple(n1), tuple(edge["node2"])
        elif isinstance(n1, str):
            cur, prox = _parse_node(n1), _parse_node(edge["node2"])
        else:
            raise TypeError("edge['node1'] must be ndarray, list or str")
        if prox in blackdict or random() < p:
            blackdict[cur] = True
            continue
        radius_list.append(radius)
        kept.append((cur[0], cur[1], cur[2], prox[0], prox[1], prox[2], radius))
    e7 = np.array(kept, dtype=np.float64).reshape(-1, 7)
    return e7, blackdict


def voxel_volume_shape(volume_dimensions: Sequence[int]):
    out = (ctypes.c_int * 3)()
    _lib.check(_lib.lib().octa_voxelize_out_dims(_lib.int3(volume_dimensions), out))
    return tuple(out)


def voxelize_edges(edges7: np.ndarray, volume_dimensions: Sequence[int], ignore_z: bool = False,
                   min_radius: float = 0.0, max_radius: float = 1.0) -> np.ndarray:
    """Host-buffer call through the C ABI (octa_voxelize_host): H2D, kernels, D2H."""
    edges7 = np.ascontiguousarray(edges7, dtype=np.float64).reshape(-1, 7)
    shape = voxel_volume_shape(volume_dimensions)
    out = np.empty(shape, dtype=np.uint16)
    opts = _lib.OctaVoxOpts(float(min_radius), float(max_radius), int(bool(ignore_z)), 0)
    _lib.check(_lib.lib().octa_voxelize_host(edges7.ctypes.data, edges7.shape[0], _lib.int3(volume_dimensions),
                                             ctypes.byref(opts), out.ctypes.data))
    return out


def voxelize_forest(forest, volume_dimensions: Sequence[float], radius_list: list = None, min_radius=0,
                    max_radius=1, max_dropout_prob=0, blackdict: dict = None, ignore_z=False):
    """Reference: tree2img.py:176-280."""
    dims = [int(d) for d in volume_dimensions]
    if any(float(d) != float(v) for d, v in zip(dims, volume_dimensions)):
        raise ValueError("volume_dimensions must be integral")
    e7, blackdict = forest_to_edges7(forest, radius_list, min_radius, max_radius, max_dropout_prob, blackdict)
    return voxelize_edges(e7, dims, ignore_z=ignore_z), blackdict


def voxelize_batch_device(edges7, edge_offsets, volume_dimensions: Sequence[int], ignore_z: bool = False,
                          min_radius: float = 0.0, max_radius: float = 1.0, out=None, workspace=None, stream=None):
    """Batched device-resident voxelization (octa_voxelize_batch_dev).

    edges7: CUDA float64 tensor [E_total, 7]; edge_offsets: host int64 sequence of n_graphs+1 entries.
    Returns a CUDA uint16 tensor [n_graphs, X, Y, Z'].  torch is used only for device memory and the
    stream handle."""
    import torch

    if not (edges7.is_cuda and edges7.dtype == torch.float64 and edges7.is_contiguous()):
        raise ValueError("edges7 must be a contiguous CUDA float64 tensor")
    offs = np.ascontiguousarray(np.asarray(edge_offsets, dtype=np.int64))
    n_graphs = offs.shape[0] - 1
    dims = _lib.int3(volume_dimensions)
    shape = voxel_volume_shape(volume_dimensions)
    if out is None:
        out = torch.empty((n_graphs, *shape), dtype=torch.uint16, device=edges7.device)
    ws_bytes = int(_lib.lib().octa_voxelize_workspace_bytes(n_graphs, int(offs[-1]), dims))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=edges7.device)
    if stream is None:
        stream = torch.cuda.current_stream(edges7.device)
    opts = _lib.OctaVoxOpts(float(min_radius), float(max_radius), int(bool(ignore_z)), 0)
    with torch.cuda.device(edges7.device):
        _lib.check(_lib.lib().octa_voxelize_batch_dev(edges7.data_ptr(), offs.ctypes.data, n_graphs, dims,
                                                      ctypes.byref(opts), out.data_ptr(), workspace.data_ptr(),
                                                      workspace.numel(), ctypes.c_void_p(stream.cuda_stream)))
    return out


def raster_edges(edges7: np.ndarray, image_resolution: Sequence[int], MIP_axis: int = 2, min_radius: float = 0.0,
                 max_radius: float = 1.0) -> np.ndarray:
    """Host-buffer call through the C ABI (octa_raster2d_host).  Returns uint8 [H, W]."""
    edges7 = np.ascontiguousarray(edges7, dtype=np.float64).reshape(-1, 7)
    W, H = int(image_resolution[0]), int(image_resolution[1])
    out = np.empty((H, W), dtype=np.uint8)
    opts = _lib.OctaVoxOpts(float(min_radius), float(max_radius), 0, 0)
    L = _lib.lib()
    L.octa_raster2d_host.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(_lib.OctaVoxOpts), ctypes.c_void_p]
    _lib.check(L.octa_raster2d_host(edges7.ctypes.data, edges7.shape[0], H, W, int(MIP_axis), ctypes.byref(opts),
                                    out.ctypes.data))
    return out


def rasterize_forest(forest, image_resolution: Sequence[float], MIP_axis: int = 2, radius_list: list = None,
                     min_radius: float = 0, max_radius: float = 1, max_dropout_prob=0, blackdict: dict = None,
                     colorize: str = None):
    """Reference: tree2img.py:12-114.  Same signature and return value ((H, W) uint16 gray 0..255, blackdict);
    same host-side semantics for the radius filter, the legacy string rows and the subtree dropout (one
    `random()` draw for p unless a blackdict is given, one per surviving edge).  `radius_list` receives
    1.3 * radius like the reference (:82-83).  `colorize` (RGB plasma rendering for figures) is not
    provided by the GPU path."""
    if colorize is not None:
        raise NotImplementedError("colorize is a matplotlib colormap feature of the reference and is not part of the GPU path")
    axes_ok = MIP_axis in (0, 1, 2)
    if not axes_ok:
        raise ValueError("MIP_axis must be 0, 1 or 2")
    rl = []
    e7, blackdict = forest_to_edges7(forest, rl, min_radius, max_radius, max_dropout_prob, blackdict)
    if radius_list is not None:
        radius_list.extend([r * 1.3 for r in rl])
    res = [int(image_resolution[0]), int(image_resolution[1])]
    return raster_edges(e7, res, MIP_axis).astype(np.uint16), blackdict


def raster_batch_device(edges7, edge_offsets, image_resolution: Sequence[int], MIP_axis: int = 2, min_radius: float = 0.0,
                        max_radius: float = 1.0, out=None, workspace=None, stream=None, layer_split=None):
    """Batched device-resident 2-D rasterization (octa_raster2d_batch_dev): CUDA uint8 tensor [n_graphs, H, W].
    `layer_split` (host int64 [n_graphs]): the first layer_split[g] edges of graph g and the rest are drawn on separate canvases
    and combined with max, as generate_vessel_graph.py:80-85 does with the arterial and the venous forest."""
    import torch

    if not (edges7.is_cuda and edges7.dtype == torch.float64 and edges7.is_contiguous()):
        raise ValueError("edges7 must be a contiguous CUDA float64 tensor")
    offs = np.ascontiguousarray(np.asarray(edge_offsets, dtype=np.int64))
    n_graphs = offs.shape[0] - 1
    W, H = int(image_resolution[0]), int(image_resolution[1])
    L = _lib.lib()
    L.octa_raster2d_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    L.octa_raster2d_workspace_bytes.restype = ctypes.c_size_t
    L.octa_raster2d_batch_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.POINTER(_lib.OctaVoxOpts), ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    if out is None:
        out = torch.empty((n_graphs, H, W), dtype=torch.uint8, device=edges7.device)
    ws_bytes = int(L.octa_raster2d_workspace_bytes(n_graphs, int(offs[-1]), H, W))
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=edges7.device)
    if stream is None:
        stream = torch.cuda.current_stream(edges7.device)
    opts = _lib.OctaVoxOpts(float(min_radius), float(max_radius), 0, 0)
    with torch.cuda.device(edges7.device):
        if layer_split is None:
            _lib.check(L.octa_raster2d_batch_dev(edges7.data_ptr(), offs.ctypes.data, n_graphs, H, W, int(MIP_axis),
                                                 ctypes.byref(opts), out.data_ptr(), workspace.data_ptr(),
                                                 workspace.numel(), ctypes.c_void_p(stream.cuda_stream)))
        else:
            sp = np.ascontiguousarray(np.asarray(layer_split, dtype=np.int64))
            if sp.shape[0] != n_graphs:
                raise ValueError("layer_split needs one entry per graph")
            L.octa_raster2d_batch_layers_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                         ctypes.c_int, ctypes.c_int, ctypes.POINTER(_lib.OctaVoxOpts), ctypes.c_void_p,
                                                         ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
            _lib.check(L.octa_raster2d_batch_layers_dev(edges7.data_ptr(), offs.ctypes.data, sp.ctypes.data, n_graphs, H, W, int(MIP_axis),
                                                        ctypes.byref(opts), out.data_ptr(), workspace.data_ptr(),
                                                        workspace.numel(), ctypes.c_void_p(stream.cuda_stream)))
    return out
