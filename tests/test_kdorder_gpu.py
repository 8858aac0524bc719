"""GPU: the block-parallel cKDTree-permutation build (csrc/octa_kdorder_par.cuh, one CTA, named-barrier warp groups +
prefix-sum Hoare partitions) must reproduce scipy's tree.indices exactly -- it decides the order in which satisfied
O2 sinks enter the `to_add` set (greenhouse.py:101-111)."""
import ctypes

import numpy as np
import pytest
from scipy.spatial import cKDTree

pytestmark = pytest.mark.gpu


def gpu_indices(pts):
    from octa_autosegmentation_b200 import _lib
    L = _lib.lib()
    L.octa_test_kd_indices_gpu.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p]
    x, y, z = [np.ascontiguousarray(pts[:, k]) for k in range(3)]
    idx = np.zeros(len(pts), dtype=np.int32)
    _lib.check(L.octa_test_kd_indices_gpu(x.ctypes.data, y.ctypes.data, z.ctypes.data, len(pts), idx.ctypes.data))
    return idx


def test_block_parallel_build_equals_scipy():
    rng = np.random.RandomState(5)
    for n in (1, 3, 16, 17, 33, 64, 100, 513, 1024, 1025, 2048, 4097, 9000, 12127, 20000, 46000):
        for rep in range(2):
            pts = rng.uniform(0, 1, (n, 3)) * np.array([1, 1, 0.0131])
            if rep == 1:
                pts[:, 2] = 0.003
            assert np.array_equal(gpu_indices(pts), cKDTree(pts).indices), (n, rep)


def test_sorted_input_depth_limit_path():
    n = 3000
    base = np.linspace(0, 1, n)
    for arr in (base, base[::-1], np.concatenate([base[::2], base[1::2][::-1]])):
        pts = np.stack([arr, np.zeros(n), np.zeros(n)], axis=1)
        assert np.array_equal(gpu_indices(pts), cKDTree(pts).indices)


def gpu_ranks_smem(pts):
    from octa_autosegmentation_b200 import _lib
    L = _lib.lib()
    L.octa_test_kd_ranks_gpu_smem.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p]
    x, y, z = [np.ascontiguousarray(pts[:, k]) for k in range(3)]
    rk = np.zeros(len(pts), dtype=np.int32)
    _lib.check(L.octa_test_kd_ranks_gpu_smem(x.ctypes.data, y.ctypes.data, z.ctypes.data, len(pts), rk.ctypes.data))
    return rk


def test_shared_memory_build_equals_scipy():
    """The shared-memory resident replay (fused median / single-scan partition) must give cKDTree's permutation.
    (n = 16000, first draw: one node exhausts introselect's depth limit -> exercises the heap-select fallback.)"""
    rng = np.random.RandomState(11)
    for n in (1, 3, 16, 17, 18, 33, 64, 100, 513, 1024, 1025, 2048, 4097, 9000, 12127, 16000, 17000):
        for rep in range(3):
            pts = rng.uniform(0, 1, (n, 3)) * np.array([1, 1, 0.0131])
            if rep == 1:
                pts[:, 2] = 0.003
            if rep == 2:
                pts = pts * np.array([0.01, 1, 1])        # z can become the split dimension deep in the tree
            idx = cKDTree(pts).indices
            want = np.empty(n, dtype=np.int32)
            want[idx] = np.arange(n, dtype=np.int32)
            got = gpu_ranks_smem(pts)
            assert got[0] != -1, "bail-out (slice wider than the 64-bit mask) is not expected at these sizes"
            assert np.array_equal(got, want), (n, rep)
