"""CPU: the Agg restatement behind the 2-D label path (oracle/agg_oracle.c) against the reference's own fixtures, and the
row-parallel formulation the CUDA kernel uses (csrc/octa_aggcells.cuh, host build) against that restatement.

The reference's rasterize_forest (tree2img.py:12-114) does its arithmetic inside matplotlib's Agg backend, which is not
installed here.  The pins are the (graph csv -> 1216^2 1-bit label) pairs the reference ships under datasets/: a label is
the gray image through PIL's Floyd-Steinberg convert("1") (visualize_vessel_graphs.py:95-101), which is chaotic in the gray
values -- the oracle has to be exact for the labels to match, and it reproduces every pixel of all 500."""
import ctypes
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, load_graph_rows, rows_to_edges7
from oracle import agg_oracle, ref_harness

SHIPPED = sorted(f[len("shipped_"):-len(".csv.gz")] for f in os.listdir(GOLDEN) if f.startswith("shipped_") and f.endswith(".csv.gz"))


def shipped_label(name):
    z = np.load(os.path.join(GOLDEN, "shipped_%s_label.npz" % name))
    return np.unpackbits(z["packed"])[: int(np.prod(z["shape"]))].reshape(z["shape"]).astype(bool)


def test_oracle_reproduces_the_committed_shipped_labels_bit_for_bit():
    assert len(SHIPPED) >= 8
    for name in SHIPPED:
        e7 = rows_to_edges7(load_graph_rows("shipped_%s.csv.gz" % name))
        lab = agg_oracle.to_label(agg_oracle.raster_edges(e7, [1216, 1216]))
        ref = shipped_label(name)
        assert lab.shape == ref.shape and int((lab != ref).sum()) == 0, name


def _one_pair(name):
    import csv
    from PIL import Image
    root = ref_harness.REFERENCE_ROOT
    with open(os.path.join(root, "datasets", "vessel_graphs", name + ".csv"), newline="") as f:
        e7 = rows_to_edges7(list(csv.DictReader(f)))
    ref = np.array(Image.open(os.path.join(root, "datasets", "labels", name + ".png")))
    lab = agg_oracle.to_label(agg_oracle.raster_edges(e7, [1216, 1216]))
    return name, int((lab != ref).sum()), float(ref.mean())


@pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference (build container)")
def test_oracle_reproduces_all_500_shipped_labels():
    import concurrent.futures as cf
    names = sorted(f[:-4] for f in os.listdir(os.path.join(ref_harness.REFERENCE_ROOT, "datasets", "vessel_graphs")) if f.endswith(".csv"))
    assert len(names) == 500
    agg_oracle.lib()
    with cf.ProcessPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as ex:
        res = list(ex.map(_one_pair, names))
    bad = [(n, m) for n, m, _ in res if m]
    assert not bad, bad
    frac = np.mean([f for _, _, f in res])
    assert abs(frac - 0.352) < 0.005          # population vessel fraction of the shipped labels


@pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference (build container)")
def test_unmodified_reference_rasterize_forest_equals_the_restated_host_loop_and_goldens():
    """tree2img.rasterize_forest itself (matplotlib calls served by oracle/shims/matplotlib) vs oracle.agg_oracle.rasterize_forest
    (restated host loop): image, blackdict, radius_list, RNG position -- and the committed r2d_small_s0.npz."""
    rows = load_graph_rows("graph_small_s0.csv")
    gold = np.load(os.path.join(GOLDEN, "r2d_small_s0.npz"))
    for res, mip, kw, key in (([304, 304], 2, {}, "a_304x304_mip2"), ([200, 120], 0, {"min_radius": 0.001}, "b_200x120_mip0_minr"),
                              ([96, 160], 1, {"max_radius": 0.002}, "c_96x160_mip1_maxr")):
        a, _ = ref_harness.rasterize(rows, res, mip, **kw)
        b, _ = agg_oracle.rasterize_forest(rows, res, mip, **kw)
        assert a.dtype == np.uint16 and np.array_equal(a, b) and np.array_equal(a, gold[key]), key
    random.seed(153)
    rl1 = []
    a, bd1 = ref_harness.rasterize(rows, [304, 304], 2, radius_list=rl1, max_dropout_prob=0.3)
    n1 = random.random()
    random.seed(153)
    rl2 = []
    b, bd2 = agg_oracle.rasterize_forest(rows, [304, 304], 2, radius_list=rl2, max_dropout_prob=0.3)
    n2 = random.random()
    assert np.array_equal(a, b) and bd1 == bd2 and rl1 == rl2 and n1 == n2 and len(bd1) > 0
    assert np.array_equal(a, gold["d_304x304_dropout"]) and n1 == float(gold["d_next_random"][0])


def _rows_host(e7, res, mip=2, minr=0.0, maxr=1.0):
    from octa_autosegmentation_b200 import _lib
    L = _lib.lib()
    L.octa_test_raster2d_rows_host.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    e7 = np.ascontiguousarray(e7, dtype=np.float64)
    W, H = res
    out = np.empty((H, W), dtype=np.uint8)
    assert L.octa_test_raster2d_rows_host(e7.ctypes.data, len(e7), H, W, mip, minr, maxr, out.ctypes.data) == 0
    return out


def hard_cases(trial, n=200):
    """Strokes leaving the canvas on every side, exactly / nearly axis-aligned ones (PathSnapper), zero-length ones, hair-thin and
    very thick ones."""
    rng = np.random.default_rng(trial)
    p = rng.uniform(-0.1, 1.1, (n, 3))
    q = p + rng.normal(0, 0.05, (n, 3))
    q[:20, 0] = p[:20, 0]
    q[20:40, 1] = p[20:40, 1] + rng.uniform(-2e-7, 2e-7, 20)
    q[40:45] = p[40:45]
    r = rng.uniform(0.0002, 0.05, n)
    r[50:60] = 0.2
    return np.concatenate([p, q, r[:, None]], 1)


def test_row_parallel_formulation_equals_the_sequential_scanline_rasterizer():
    """csrc/octa_aggcells.cuh evaluates Agg's scanline DDA in closed form, one row (and one 32-pixel tile) at a time."""
    e7 = rows_to_edges7(load_graph_rows("graph_small_s0.csv"))
    for res, mip in (([304, 304], 2), ([97, 61], 0), ([1216, 1216], 2)):
        assert np.array_equal(_rows_host(e7, res, mip), agg_oracle.raster_edges(e7, res, MIP_axis=mip))
    for trial in range(12):
        e7 = hard_cases(trial)
        for res, mip in (([160, 120], 2), ([64, 200], 0), ([333, 333], 1)):
            a, b = _rows_host(e7, res, mip, 0.0003, 0.1), agg_oracle.raster_edges(e7, res, MIP_axis=mip, min_radius=0.0003, max_radius=0.1)
            assert np.array_equal(a, b), (trial, res, mip, int((a != b).sum()))
