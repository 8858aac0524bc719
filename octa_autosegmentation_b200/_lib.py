"""ctypes binding of the C-ABI library (include/octa_b200.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry point
fails loudly, and if no CUDA device is present every compute call raises OctaError.
"""
from __future__ import annotations

import ctypes
import os

# The pipelined API keeps several growth loops in flight, two streams each.  With the default of 8 hardware work queues,
# streams alias onto the same queue and serialise (3 loops in flight measured SLOWER than 2); 32 queues remove that.  Must be
# set before the CUDA context exists, hence at import time; an explicit setting of the user wins.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libocta_b200.so")

OCTA_OK, OCTA_E_ARG, OCTA_E_CUDA, OCTA_E_NOMEM, OCTA_E_STATE = 0, -1, -2, -3, -4


class OctaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("octa_b200 error %d: %s" % (code, msg))
        self.code = code


class OctaVoxOpts(ctypes.Structure):
    _fields_ = [("min_radius", ctypes.c_double), ("max_radius", ctypes.c_double),
                ("ignore_z", ctypes.c_int), ("reserved", ctypes.c_int)]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: run `python -m octa_autosegmentation_b200.build` (nvcc, sm_100a). "
            "There is no CPU fallback for the vessel-graph hot path." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c_int3 = ctypes.POINTER(ctypes.c_int)
    L.octa_abi_version.restype = ctypes.c_int
    L.octa_last_error.restype = ctypes.c_char_p
    L.octa_launch_count.restype = ctypes.c_uint64
    L.octa_device_count.restype = ctypes.c_int
    L.octa_voxelize_out_dims.argtypes = [c_int3, c_int3]
    L.octa_voxelize_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, c_int3]
    L.octa_voxelize_workspace_bytes.restype = ctypes.c_size_t
    L.octa_voxelize_batch_dev.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_int3,
                                          ctypes.POINTER(OctaVoxOpts), ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.c_void_p]
    L.octa_voxelize_host.argtypes = [ctypes.c_void_p, ctypes.c_int64, c_int3, ctypes.POINTER(OctaVoxOpts),
                                     ctypes.c_void_p]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise OctaError(rc, lib().octa_last_error().decode(errors="replace"))


def int3(v):
    return (ctypes.c_int * 3)(*[int(x) for x in v])


def launch_count() -> int:
    return int(lib().octa_launch_count())
