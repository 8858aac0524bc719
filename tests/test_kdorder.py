"""CPU (distinct coordinates, as produced by the jittered sink sampler; exact coordinate ties on the split axis are not
covered): the product's restatement of cKDTree's build permutation (csrc/octa_kdorder.h: scipy build.cxx +
libstdc++ std::nth_element, host build of the host/device code) against scipy itself."""
import ctypes

import numpy as np
from scipy.spatial import cKDTree

from octa_autosegmentation_b200 import _lib


def product_indices(pts):
    L = _lib.lib()
    L.octa_test_kd_indices.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int, ctypes.c_void_p]
    L.octa_test_kd_indices.restype = None
    x, y, z = [np.ascontiguousarray(pts[:, k]) for k in range(3)]
    idx = np.zeros(len(pts), dtype=np.int32)
    L.octa_test_kd_indices(x.ctypes.data, y.ctypes.data, z.ctypes.data, len(pts), idx.ctypes.data)
    return idx


def test_indices_equal_scipy_on_sink_like_clouds():
    rng = np.random.RandomState(3)
    for trial in range(120):
        n = int(rng.choice([1, 5, 16, 17, 33, 100, 777, 2048, 5000, 12000]))
        pts = rng.uniform(0, 1, (n, 3)) * np.array([1, 1, 0.0131])
        if trial % 9 == 0:
            pts[:, 2] = 0.004                              # degenerate axis
        assert np.array_equal(product_indices(pts), cKDTree(pts).indices), (trial, n)


def test_adversarial_order_hits_heap_select_path():
    # sorted / organ-pipe inputs drive introselect's depth limit; the permutation must still match
    for n in (1000, 4097):
        base = np.linspace(0, 1, n)
        for arr in (base, base[::-1], np.concatenate([base[::2], base[1::2][::-1]])):
            pts = np.stack([arr, np.zeros(n), np.zeros(n)], axis=1)
            assert np.array_equal(product_indices(pts), cKDTree(pts).indices)
