"""GPU parity tests for K6 (octa_voxelize.cu) through the C ABI.  Bar: bit-exact uint16 volumes
against the CPU oracle (itself pinned bit-exact to tree2img.voxelize_forest) and against the
committed reference fixtures."""
import glob
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, load_graph_rows, rows_to_edges7

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def t2i():
    from octa_autosegmentation_b200 import _lib, tree2img
    assert _lib.lib().octa_device_count() > 0, "GPU tests need a CUDA device"
    return tree2img


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "vox_small_s0_*.npz"))))
def test_reference_fixtures_bit_exact(t2i, path):
    z = np.load(path)
    kw = json.loads(str(z["kw"]))
    rows = load_graph_rows("graph_small_s0.csv")
    vol, _ = t2i.voxelize_forest(rows, [int(d) for d in z["dims"]], **kw)
    assert vol.dtype == np.uint16 and vol.shape == z["vol"].shape
    assert np.array_equal(vol, z["vol"])


def test_docker_graph_vs_reference_digest_and_oracle(t2i):
    from oracle import vox_oracle
    gold = json.load(open(os.path.join(GOLDEN, "vox_docker_s0.json")))
    e7 = rows_to_edges7(load_graph_rows("graph_docker_s0.csv.gz"))
    vol = t2i.voxelize_edges(e7, [304, 304, 4])
    assert hashlib.sha256(vol.tobytes()).hexdigest() == gold["304x304x4"]["sha256"]
    # BASELINE config #2 request: [1216,1216,16] -> (1216,1216,53)
    vol = t2i.voxelize_edges(e7, [1216, 1216, 16])
    ref = vox_oracle.voxelize_edges(e7, [1216, 1216, 16])
    assert vol.shape == (1216, 1216, 53)
    assert np.array_equal(vol, ref)
    if "1216x1216x16" in gold:
        assert hashlib.sha256(vol.tobytes()).hexdigest() == gold["1216x1216x16"]["sha256"]
    # config #4 depth: Z' = 64
    vol = t2i.voxelize_edges(e7, [1216, 1216, 64])
    assert np.array_equal(vol, vox_oracle.voxelize_edges(e7, [1216, 1216, 64]))


def test_edge_cases_vs_oracle(t2i):
    from oracle import vox_oracle
    rng = np.random.RandomState(3)
    dims = [72, 40, 33]
    cases = {
        "empty": np.zeros((0, 7)),
        "degenerate": np.array([[0.5, 0.3, 0.2, 0.5, 0.3, 0.2, 0.02]]),          # zero-length segment
        "outside": np.array([[1.5, 1.5, 1.5, 1.7, 1.6, 1.5, 0.01], [-0.5, -0.2, 0, -0.4, -0.1, 0, 0.01]]),
        "crossing": np.array([[-0.2, 0.1, 0.1, 1.3, 0.5, 0.4, 0.03]]),          # spans the volume -> "big" list
        "axis": np.array([[0.1, 0.25, 0.25, 0.9, 0.25, 0.25, 0.004]]),
        "thin": np.array([[0.1, 0.1, 0.1, 0.3, 0.3, 0.2, 1e-6]]),
        "filtered": np.array([[0.1, 0.1, 0.1, 0.3, 0.3, 0.2, 2.0]]),            # radius > max_radius
        "random": np.concatenate([rng.uniform(-0.1, 1.1, (400, 6)), rng.uniform(0.0005, 0.03, (400, 1))], axis=1),
    }
    for name, e7 in cases.items():
        for ignore_z in (False, True):
            got = t2i.voxelize_edges(e7, dims, ignore_z=ignore_z)
            ref = vox_oracle.voxelize_edges(e7, dims, ignore_z=ignore_z)
            assert np.array_equal(got, ref), (name, ignore_z, int((got != ref).sum()))
    # tall volume: more than one z tile
    e7 = cases["random"]
    assert np.array_equal(t2i.voxelize_edges(e7, [40, 36, 150]), vox_oracle.voxelize_edges(e7, [40, 36, 150]))


def test_properties_at_full_size(t2i):
    """Size-independent properties at the BASELINE raster size: max-composition (art/ven volumes
    max-combined == one pass over all edges, generate_vessel_graph.py:70-72), permutation
    invariance, idempotence of duplicated edges."""
    e7 = rows_to_edges7(load_graph_rows("graph_docker_s0.csv.gz"))
    dims = [1216, 1216, 16]
    full = t2i.voxelize_edges(e7, dims)
    a, b = t2i.voxelize_edges(e7[:9025], dims), t2i.voxelize_edges(e7[9025:], dims)
    assert np.array_equal(np.maximum(a, b), full)
    perm = np.random.RandomState(0).permutation(len(e7))
    assert np.array_equal(t2i.voxelize_edges(e7[perm], dims), full)
    assert np.array_equal(t2i.voxelize_edges(np.concatenate([e7, e7[:500]]), dims), full)
    assert full.max() == 255 and 4_000_000 < int((full > 0).sum()) < 6_000_000


def test_batch_device_api_matches_single(t2i):
    import torch
    small = rows_to_edges7(load_graph_rows("graph_small_s0.csv"))
    small1 = rows_to_edges7(load_graph_rows("graph_small_s1.csv"))
    graphs = [small, small1, np.zeros((0, 7)), small[:100]]
    offs = np.cumsum([0] + [len(g) for g in graphs])
    e = torch.from_numpy(np.concatenate(graphs)).cuda()
    out = t2i.voxelize_batch_device(e, offs, [304, 304, 4])
    torch.cuda.synchronize()
    assert out.shape == (4, 304, 304, 14) and out.dtype == torch.uint16
    host = out.cpu().numpy()
    for i, g in enumerate(graphs):
        assert np.array_equal(host[i], t2i.voxelize_edges(g, [304, 304, 4])), i


def test_dropout_matches_oracle_host_logic(t2i):
    from oracle import vox_oracle
    rows = load_graph_rows("graph_small_s0.csv")
    random.seed(153)
    got, bd = t2i.voxelize_forest(rows, [128, 128, 4], max_dropout_prob=0.05)
    s_after = random.random()
    random.seed(153)
    ref, bd_ref = vox_oracle.voxelize_forest(rows, [128, 128, 4], max_dropout_prob=0.05)
    assert random.random() == s_after and bd == bd_ref and np.array_equal(got, ref)
    rl = []
    got2, _ = t2i.voxelize_forest(rows, [128, 128, 4], radius_list=rl, min_radius=0.001, blackdict=dict(bd))
    ref2, _ = vox_oracle.voxelize_forest(rows, [128, 128, 4], min_radius=0.001, blackdict=dict(bd))
    assert np.array_equal(got2, ref2) and len(rl) > 0 and min(rl) >= 0.001


@pytest.mark.parametrize("knob", [{"OCTA_VOX_SLOWCAP": "0"}, {"OCTA_VOX_SLOWCAP": "3"}, {"OCTA_VOX_TILE_Y": "16"}, {"OCTA_VOX_KERNEL": "rows"}])
def test_fallback_paths_give_the_same_volume(t2i, monkeypatch, knob):
    """The paths real graphs never (or only by choice) take: a full deferred-cell queue (cells are marked and recomputed from every
    edge of the tile), half-size tiles with four warps, and the row kernel of round 1 -- all bit-identical to the reference digest."""
    gold = json.load(open(os.path.join(GOLDEN, "vox_docker_s0.json")))
    e7 = rows_to_edges7(load_graph_rows("graph_docker_s0.csv.gz"))
    for k, v in knob.items():
        monkeypatch.setenv(k, v)
    vol = t2i.voxelize_edges(e7, [304, 304, 4])
    assert hashlib.sha256(vol.tobytes()).hexdigest() == gold["304x304x4"]["sha256"]
    if "1216x1216x16" in gold:
        vol = t2i.voxelize_edges(e7, [1216, 1216, 16])
        assert hashlib.sha256(vol.tobytes()).hexdigest() == gold["1216x1216x16"]["sha256"]
