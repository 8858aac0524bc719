"""CPU: pins oracle/voxelize_oracle.c to the real reference through the committed fixtures
(tests/golden/vox_*.npz|json were written by oracle/make_golden.py from tree2img.voxelize_forest)."""
import glob
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, load_graph_rows, rows_to_edges7
from oracle import vox_oracle


def test_out_dims_match_reference_formula():
    # tree2img.py:206-210; (1216,1216,16) -> 53 is the value measured on the reference
    assert vox_oracle.out_dims([1216, 1216, 16]) == (1216, 1216, 53)
    assert vox_oracle.out_dims([304, 304, 4]) == (304, 304, 14)
    assert vox_oracle.out_dims([1216, 1216, 64]) == (1216, 1216, 64)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "vox_small_s0_*.npz"))))
def test_oracle_bit_exact_vs_reference_small(path):
    z = np.load(path)
    kw = json.loads(str(z["kw"]))
    rows = load_graph_rows("graph_small_s0.csv")
    vol, _ = vox_oracle.voxelize_forest(rows, [int(d) for d in z["dims"]], **kw)
    assert vol.dtype == np.uint16 and vol.shape == z["vol"].shape
    assert np.array_equal(vol, z["vol"])


def test_oracle_vs_reference_docker_digest():
    with open(os.path.join(GOLDEN, "vox_docker_s0.json")) as f:
        gold = json.load(f)
    e7 = rows_to_edges7(load_graph_rows("graph_docker_s0.csv.gz"))
    vol = vox_oracle.voxelize_edges(e7, [304, 304, 4])
    g = gold["304x304x4"]
    assert list(vol.shape) == g["shape"]
    assert int((vol > 0).sum()) == g["nonzero"]
    assert hashlib.sha256(vol.tobytes()).hexdigest() == g["sha256"]


def test_dropout_and_blackdict_semantics():
    """tree2img.py:220-224,238-240: subtree dropout keyed on parent position; a supplied blackdict
    disables random dropout; RNG consumption = 1 + one draw per non-blacklisted surviving edge."""
    rows = load_graph_rows("graph_small_s0.csv")
    random.seed(153)
    vol_a, bd = vox_oracle.voxelize_forest(rows, [96, 96, 4], max_dropout_prob=0.05)
    assert len(bd) > 0
    after = random.random()
    random.seed(153)
    vol_b, bd_b = vox_oracle.voxelize_forest(rows, [96, 96, 4], max_dropout_prob=0.05)
    assert after == random.random() and np.array_equal(vol_a, vol_b) and bd == bd_b
    # blackdict supplied -> p = 0 and no draw for p (:223-224); only edges whose PARENT is listed are
    # dropped, so the randomly dropped edges themselves come back while their subtrees stay away
    random.seed(1)
    vol_c, bd_c = vox_oracle.voxelize_forest(rows, [96, 96, 4], max_dropout_prob=0.05, blackdict=dict(bd))
    assert (vol_a <= vol_c).all() and set(bd).issubset(bd_c)
    full, _ = vox_oracle.voxelize_forest(rows, [96, 96, 4])
    assert (vol_a <= full).all() and (vol_a < full).any()


@pytest.mark.skipif(not os.path.isdir("/root/reference/vessel_graph_generation"), reason="needs /root/reference (build container)")
def test_random_requests_vs_the_reference_itself():
    """Random volume shapes (flat, tall, permuted axes), row ranges of three reference graphs, radius filters and ignore_z:
    the oracle against `tree2img.voxelize_forest` of the unmodified reference run here."""
    from oracle import ref_harness as rh
    rng = np.random.default_rng(21)
    graphs = [load_graph_rows(n) for n in ("graph_small_s0.csv", "graph_geom3d_s0.csv", "graph_docker_s0.csv.gz")]
    filled = 0
    for case in range(12):
        rows = graphs[case % 3]
        k = int(rng.integers(20, 400))
        start = int(rng.integers(0, len(rows) - k))
        dims = [int(rng.integers(8, 200)), int(rng.integers(8, 200)), int(rng.integers(1, 40))]
        if case % 4 == 0:
            dims[2] = 1
        if case % 5 == 4:
            dims = [dims[2] + 3, dims[0], dims[1]]
        kw = {}
        if rng.random() < 0.3:
            kw["ignore_z"] = True
        if rng.random() < 0.3:
            kw["min_radius"] = float(rng.uniform(0.0005, 0.002))
        if rng.random() < 0.3:
            kw["max_radius"] = float(rng.uniform(0.002, 0.01))
        ref, _ = rh.voxelize(rows[start:start + k], dims, **kw)
        got, _ = vox_oracle.voxelize_forest(rows[start:start + k], dims, **kw)
        assert ref.shape == got.shape and np.array_equal(ref, got), (case, dims, kw)
        filled += int((ref > 0).any())
    assert filled >= 8
