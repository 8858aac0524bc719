"""CPU: pins oracle/growth_oracle.cpp to the real reference.

The graph_*.csv fixtures were written by oracle/ref_harness.py from the UNMODIFIED reference modules
(seeded random / np.random); the oracle must reproduce them byte for byte.  The emulated pieces
(two MT19937 streams, CPython set order, cKDTree index permutation) are additionally checked
against the real CPython / numpy / scipy in this interpreter."""
import gzip
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import growth_oracle as go


def docker_config():
    from octa_autosegmentation_b200.config import default_config
    return default_config()


def small_config():
    cfg = docker_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (12, 12)):
        m["I"], m["N"] = i, 400
    return cfg


def oracle_csv(cfg, seed, **kw):
    art, ven, st = go.run(cfg, seed, **kw)
    return go.csv_bytes(np.concatenate([art, ven])), st


def test_python_random_stream():
    for seed in (0, 1, 12345, 2**31 + 7, 2**40 + 3):
        out = np.empty(50)
        go.lib().og_py_random(seed, 50, out.ctypes.data)
        random.seed(seed)
        assert [random.random() for _ in range(50)] == list(out)
        ch = (np.zeros(64, dtype=np.int32))
        go.lib().og_py_choice(seed, 4, 64, ch.ctypes.data)
        random.seed(seed)
        assert [random.choice([0, 1, 2, 3]) for _ in range(64)] == list(ch)


def test_numpy_legacy_stream():
    import ctypes
    for seed in (0, 7, 4242):
        nrm = ctypes.c_double()
        ints = np.zeros(3000, dtype=np.uint32)
        dbl = np.zeros(999)
        go.lib().og_np_stream(seed, ctypes.byref(nrm), 3000, 5663, ints.ctypes.data, 999, dbl.ctypes.data)
        np.random.seed(seed)
        assert np.random.normal(0.25, 0.5) == nrm.value
        assert np.array_equal(np.random.randint(0, 5663, 3000), ints)
        assert np.array_equal(np.random.uniform(0, 1, (333, 3)).ravel(), dbl)


def test_cpython_tuple_hash_and_set_order():
    rng = np.random.RandomState(0)
    for trial in range(200):
        n = int(rng.randint(1, 400))
        pts = rng.uniform(-0.01, 1.01, (n, 3))
        dup = rng.randint(0, n, n // 5)
        pts = np.concatenate([pts, pts[dup]])          # duplicates, as when two new nodes hit one sink
        pts = pts[rng.permutation(len(pts))]
        tuples = [tuple(np.float64(v) for v in p) for p in pts]
        assert go.lib().og_hash_tuple3(np.ascontiguousarray(pts[0]).ctypes.data) == hash(tuples[0])
        s = set()
        for t in tuples:
            s.add(t)
        order = np.zeros(len(pts), dtype=np.int64)
        k = go.lib().og_set_order(np.ascontiguousarray(pts).ctypes.data, len(pts), order.ctypes.data)
        assert [tuples[i] for i in order[:k]] == list(s)


def test_ckdtree_index_permutation():
    from scipy.spatial import cKDTree
    rng = np.random.RandomState(1)
    for trial in range(60):
        n = int(rng.randint(1, 3000))
        pts = rng.uniform(0, 1, (n, 3)) * np.array([1, 1, 0.0131])
        if trial % 7 == 0:
            pts[:, 2] = 0.005                      # zero extent along one axis
        idx = np.zeros(n, dtype=np.int64)
        go.lib().og_kd_indices(np.ascontiguousarray(pts).ctypes.data, n, idx.ctypes.data)
        tree = cKDTree(pts)
        assert np.array_equal(tree.indices, idx)
        # ball results come back in ascending position of tree.indices
        rank = np.empty(n, dtype=np.int64)
        rank[idx] = np.arange(n)
        q = pts[rng.randint(n)]
        hits = tree.query_ball_point(q, 0.08)
        assert list(hits) == sorted(hits, key=lambda i: rank[i])


@pytest.mark.parametrize("seed", [0, 1])
def test_small_config_byte_exact(seed):
    got, _ = oracle_csv(small_config(), seed)
    assert got == open(os.path.join(GOLDEN, "graph_small_s%d.csv" % seed), "rb").read()


@pytest.mark.parametrize("seed", [0, 1])
def test_fixed_geometry_byte_exact(seed):
    """SimulationSpace.oxygen_sample_geometry_path (simulation_space.py:26-34,69-76,95-96): wall positions through
    random.choice over the mask's wall plane, sampling from argwhere(mask), mask lookup in is_valid_position."""
    cfg = small_config()
    cfg["Greenhouse"]["SimulationSpace"]["oxygen_sample_geometry_path"] = os.path.join(GOLDEN, "geometry_mask.npy")
    got, _ = oracle_csv(cfg, seed)
    assert got == open(os.path.join(GOLDEN, "graph_geom_s%d.csv" % seed), "rb").read()


def geom3d_config():
    """oracle/make_golden.py geom3d_config(): 3-D mask [40, 84, 8], trees rooted on x0, y0, y1, z0, z1."""
    cfg = docker_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (25, 15)):
        m["I"], m["N"] = i, 500
    cfg["Greenhouse"]["SimulationSpace"]["oxygen_sample_geometry_path"] = os.path.join(GOLDEN, "geometry_mask_3d.npy")
    cfg["Forest"]["source_walls"] = {"x0": True, "x1": False, "y0": True, "y1": True, "z0": True, "z1": True}
    cfg["Forest"]["N_trees"] = 6
    return cfg


@pytest.mark.parametrize("seed", [0, 1])
def test_3d_geometry_and_z_walls_byte_exact(seed):
    """A 3-D sampling mask (simulation_space.py:29-34: any .npy; geometry_size 84, argwhere triples, 3-D mask lookup) and
    the z0 / z1 source walls that only work with a geometry file (forest.py:152-176, simulation_space.py:69-76)."""
    got, _ = oracle_csv(geom3d_config(), seed)
    assert got == open(os.path.join(GOLDEN, "graph_geom3d_s%d.csv" % seed), "rb").read()


def test_save_stats_data_vs_reference():
    """The data Greenhouse.save_stats plots (greenhouse.py:401-441), as the unmodified reference held it after growing
    graph_small_s0 (oracle/make_golden.py): final oxygen-sink / CO2-source lists in list order and the per-iteration counts."""
    tr = []
    go.run(small_config(), 0, trace=lambda t, a, o, v, c, pd, nd: tr.append((a, o, v, c)))
    oxy, co2 = go.last_sinks()
    gold = np.load(os.path.join(GOLDEN, "stats_small_s0.npz"))
    assert np.array_equal(oxy, gold["oxys"]) and np.array_equal(co2, gold["co2s"])
    assert not gold["per_step"][0].any() and np.array_equal(np.array(tr), gold["per_step"][1:])


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_docker_config_byte_exact(seed):
    """BASELINE config #1: docker/vessel_graph_gen_docker_config.yml, fixed seed."""
    got, st = oracle_csv(docker_config(), seed)
    dig = json.load(open(os.path.join(GOLDEN, "graph_docker_digests.json")))["docker_s%d" % seed]
    assert len(got) == dig["bytes"] and hashlib.sha256(got).hexdigest() == dig["sha256"]
    p = os.path.join(GOLDEN, "graph_docker_s%d.csv.gz" % seed)
    if os.path.exists(p):
        assert got == gzip.open(p, "rb").read()
    if seed == 0:   # counts measured on the reference itself
        assert st["py_draws"] == 14778 and st["nn_queries"] == 1820486 and st["bifurcations"] == 41
        assert st["n_art_nodes"] == 9033 and st["n_ven_nodes"] == 3965 and st["n_oxy_left"] == 12127 and st["n_co2_left"] == 3567


def repeated_name_config():
    """Three modes named SVC, DVC, SVC: the reference re-initialises its parameters only for a mode whose name differs from the
    first mode's (greenhouse.py:84-85), so the third mode runs with EVERYTHING of the second (I, N, eps, delta, gamma, phi, omega,
    kappa) -- the values written in its own block are never read."""
    import copy
    cfg = small_config()
    modes = cfg["Greenhouse"]["modes"]
    third = copy.deepcopy(modes[0])
    third.update(I=7, N=300, gamma_art=20, gamma_ven=20, phi=40, omega=0.9, kappa=3.7, eps_k=0.09)
    modes.append(third)
    modes[1]["I"] = 6
    return cfg


@pytest.mark.skipif(not os.path.isdir("/root/reference/vessel_graph_generation"), reason="needs /root/reference (build container)")
def test_mode_with_the_first_modes_name_keeps_previous_parameters_like_the_reference():
    from oracle import ref_harness as rh
    cfg = repeated_name_config()
    art, ven, _ = rh.run_growth(cfg, 3)
    got, _ = oracle_csv(cfg, 3)
    assert got == rh.csv_bytes(art, ven)
    # and it is NOT what taking the third block's own values would give
    import copy
    other = copy.deepcopy(cfg)
    other["Greenhouse"]["modes"][2]["name"] = "third"
    assert oracle_csv(other, 3)[0] != got


@pytest.mark.skipif(not os.path.isdir("/root/reference/vessel_graph_generation"), reason="needs /root/reference (build container)")
def test_random_3d_masks_and_wall_sets_vs_the_reference_itself(tmp_path):
    """Random sampling geometries (flat, thick, longer than the default 76 voxels) with random source-wall sets, tree counts,
    schedules and seeds: CSV bytes and final sink lists of the oracle against the unmodified reference run here."""
    from oracle import ref_harness as rh
    rng = np.random.default_rng(11)
    for case in range(6):
        shape = tuple(int(x) for x in rng.integers(3, 50, size=3))
        if case % 3 == 0:
            shape = (shape[0], shape[1], 1)
        if case == 4:
            shape = (int(rng.integers(80, 120)), shape[1], int(rng.integers(1, 6)))
        g = rng.random(shape) > 0.25
        g[0, :, :] |= rng.random(shape[1:]) > 0.5          # keep the wall planes populated (an empty plane raises in random.choice)
        g[:, 0, :] |= rng.random((shape[0], shape[2])) > 0.5
        g[:, :, 0] |= rng.random(shape[:2]) > 0.5
        np.save(tmp_path / "mask.npy", g)
        cfg = docker_config()
        for m, i in zip(cfg["Greenhouse"]["modes"], (int(rng.integers(8, 16)), int(rng.integers(5, 12)))):
            m["I"], m["N"] = i, int(rng.integers(200, 500))
        cfg["Greenhouse"]["SimulationSpace"]["oxygen_sample_geometry_path"] = str(tmp_path / "mask.npy")
        walls = {k: bool(rng.random() > 0.4) for k in ("x0", "x1", "y0", "y1", "z0", "z1")}
        walls["x0"] = walls["x0"] or not any(walls.values())
        cfg["Forest"]["source_walls"] = walls
        cfg["Forest"]["N_trees"] = int(rng.integers(2, 9))
        seed = int(rng.integers(0, 1000))
        art, ven, gh = rh.run_growth(cfg, seed)
        got, _ = oracle_csv(cfg, seed)
        assert got == rh.csv_bytes(art, ven), (case, shape, walls)
        oxy, co2 = go.last_sinks()
        assert np.array_equal(oxy, np.array(gh.oxy_mesh.get_all_elements()).reshape(-1, 3)), (case, shape)
        assert np.array_equal(co2, np.array(gh.co2_mesh.get_all_elements()).reshape(-1, 3)), (case, shape)
