"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def libpath():
    from octa_autosegmentation_b200 import build
    return build.build_library(verbose=False)


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(octa_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_exports_every_declared_symbol(libpath):
    L = ctypes.CDLL(libpath)
    syms = declared_symbols()
    assert "octa_voxelize_batch_dev" in syms and len(syms) >= 6
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s


def test_version_and_no_device_behaviour(libpath):
    from octa_autosegmentation_b200 import _lib, tree2img
    assert _lib.lib().octa_abi_version() == 3
    assert tree2img.voxel_volume_shape([1216, 1216, 16]) == (1216, 1216, 53)
    assert _lib.lib().octa_voxelize_workspace_bytes(1, 1000, _lib.int3([304, 304, 4])) > 0
    if _lib.lib().octa_device_count() == 0:
        # the product path must fail loudly, never fall back to a CPU implementation
        with pytest.raises(_lib.OctaError):
            tree2img.voxelize_edges(np.zeros((1, 7)), [32, 32, 32])


def test_product_does_not_import_oracle():
    """Guard for the layering rule: nothing under the package may reference oracle/."""
    pkg = os.path.join(ROOT, "octa_autosegmentation_b200")
    for p in glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True) + glob.glob(os.path.join(pkg, "csrc", "*")):
        if os.path.isfile(p):
            assert not re.search(r"^\s*(from|import)\s+oracle\b|oracle/", open(p, errors="ignore").read(), re.M), p
