"""GPU parity tests of the growth kernels through the C ABI (octa_grow_batch_host) against the CPU
oracle (oracle/growth_oracle.cpp, itself byte-identical to the reference on the committed goldens)
and against the reference's own seeded CSVs.

Bar: topology, row order and radii bit-exact; positions equal to the digits the CSV prints (the
device's acos/sin/cos/exp differ from glibc/SVML by <= 2 ULP, far below the 8 printed digits)."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def growth():
    from octa_autosegmentation_b200 import _lib, growth
    assert _lib.lib().octa_device_count() > 0
    return growth


def small_config():
    from octa_autosegmentation_b200.config import default_config
    cfg = default_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (12, 12)):
        m["I"], m["N"] = i, 400
    return cfg


def numpy_csv(e7):
    """The product's own CSV writer (C ABI octa_format_csv), cross-checked against numpy/csv formatting."""
    from octa_autosegmentation_b200 import graph_io
    from oracle import growth_oracle as go
    out = graph_io.csv_bytes(e7)
    assert out == go.csv_bytes(e7)
    return out


def compare_with_oracle(growth, cfg, seeds):
    from oracle import growth_oracle as go
    graphs, stats, extra = growth.grow_batch(cfg, seeds, trace=True)
    for i, s in enumerate(seeds):
        art, ven = graphs[i]
        oa, ov, ost = go.run(cfg, s)                     # exact oracle: cKDTree ball-result order
        if len(art) != len(oa) or len(ven) != len(ov):
            # locate the first diverging iteration for the failure message
            tr = []
            go.run(cfg, s, trace=lambda t, a, o, v, c, pd, nd: tr.append((a, o, v, c)))
            mine = extra["trace"][i]
            first = next((k for k in range(len(tr)) if tuple(mine[k]) != tr[k]), None)
            raise AssertionError("seed %d: edge counts differ (%d,%d) vs oracle (%d,%d); first diverging iteration %s: %s vs %s"
                                 % (s, len(art), len(ven), len(oa), len(ov), first,
                                    None if first is None else tuple(mine[first]), None if first is None else tr[first]))
        e, o = np.concatenate([art, ven]), np.concatenate([oa, ov])
        assert np.array_equal(e[:, 6], o[:, 6]), "radii must be bit-exact (seed %d)" % s
        assert np.abs(e[:, :6] - o[:, :6]).max() < 1e-11, "positions (seed %d)" % s
        assert stats[i]["py_draws"] == ost["py_draws"]
        assert (stats[i]["n_oxy_left"], stats[i]["n_co2_left"]) == (ost["n_oxy_left"], ost["n_co2_left"])
    return graphs, stats


def test_small_config_vs_oracle_and_reference_csv(growth):
    graphs, _ = compare_with_oracle(growth, small_config(), [0, 1, 2, 3, 4, 5, 6, 7])
    for seed in (0, 1):
        got = numpy_csv(np.concatenate(graphs[seed]))
        assert got == open(os.path.join(GOLDEN, "graph_small_s%d.csv" % seed), "rb").read()


def test_docker_config_vs_reference_csv(growth):
    """BASELINE config #1/#2: docker config verbatim; seeds 0-3 must reproduce the reference's CSV bytes."""
    from octa_autosegmentation_b200.config import default_config
    graphs, stats = compare_with_oracle(growth, default_config(), [0, 1, 2, 3])
    dig = json.load(open(os.path.join(GOLDEN, "graph_docker_digests.json")))
    for seed in range(4):
        got = numpy_csv(np.concatenate(graphs[seed]))
        d = dig["docker_s%d" % seed]
        assert len(got) == d["bytes"] and hashlib.sha256(got).hexdigest() == d["sha256"], "seed %d" % seed
    assert stats[0]["py_draws"] == 14778 and stats[0]["n_art_nodes"] == 9033 and stats[0]["n_ven_nodes"] == 3965


def test_batch_is_order_independent(growth):
    """Multi-GPU contract: a sample's output depends only on its seed, not on the batch it is in."""
    cfg = small_config()
    a, _, _ = growth.grow_batch(cfg, [5, 3, 9])
    b, _, _ = growth.grow_batch(cfg, [9, 5])
    assert np.array_equal(a[0][0], b[1][0]) and np.array_equal(a[0][1], b[1][1])
    assert np.array_equal(a[2][0], b[0][0]) and np.array_equal(a[2][1], b[0][1])


def test_nerve_forest_12x12_variant_vs_reference_csv(growth):
    """Forest.type 'nerve', param_scale 12, 16 trees (example_custom_vessel_simulation.ipynb :137-160, shortened to
    I = 70 + 50, N = 1500): the nerve disk is carved out of the sampling mask and the stumps start at the optic nerve."""
    from octa_autosegmentation_b200.config import default_config
    cfg = default_config()
    g = cfg["Greenhouse"]
    g["param_scale"] = 12
    cfg["Forest"]["type"] = "nerve"
    cfg["output"]["image_scale_factor"] = 1216
    g["SimulationSpace"]["no_voxel_z"] = 0.0033
    g["d"] = 0.15
    for m, i in zip(g["modes"], (70, 50)):
        m["I"], m["N"], m["delta_sigma"] = i, 1500, 0.002222
    cfg["Forest"]["N_trees"] = 16
    graphs, _ = compare_with_oracle(growth, cfg, [0, 1, 2])
    assert numpy_csv(np.concatenate(graphs[0])) == open(os.path.join(GOLDEN, "graph_nerve_s0.csv"), "rb").read()


def test_stress_config_vs_oracle(growth):
    """BASELINE config #4 growth part: 4x attraction points (N = 8000 in both modes).  The reference needs ~4 min per
    sample for this, so the pin is the oracle (byte-identical to the reference on every committed golden)."""
    from octa_autosegmentation_b200.config import default_config
    from oracle import growth_oracle as go
    cfg = default_config()
    for m in cfg["Greenhouse"]["modes"]:
        m["N"] = 8000
    graphs, stats, _ = growth.grow_batch(cfg, [0, 1], cap_edges=60000)
    for seed, (art, ven) in zip((0, 1), graphs):
        oa, ov, _ = go.run(cfg, seed)
        e, o = np.concatenate([art, ven]), np.concatenate([oa, ov])
        assert e.shape == o.shape and np.array_equal(e[:, 6], o[:, 6]) and np.abs(e[:, :6] - o[:, :6]).max() < 1e-11
    assert len(graphs[0][0]) + len(graphs[0][1]) == 16351      # edge count measured on the reference


def test_round_robin_sharding_gives_identical_files(growth):
    """BASELINE config #3 contract: sample i -> rank i mod world; outputs must be byte-identical to the 1-GPU run."""
    from octa_autosegmentation_b200 import graph_io
    from octa_autosegmentation_b200.pipeline import shard_seeds
    cfg = small_config()
    base, n = 40, 10
    single, _, _ = growth.grow_batch(cfg, [base + i for i in range(n)])
    ref = {base + i: graph_io.csv_bytes(np.concatenate(single[i])) for i in range(n)}
    for world in (2, 4):
        for rank in range(world):
            seeds = shard_seeds(base, n, rank, world)
            got, _, _ = growth.grow_batch(cfg, seeds)
            for s, g in zip(seeds, got):
                assert graph_io.csv_bytes(np.concatenate(g)) == ref[s]


def test_commit_global_tree_view_fallback(monkeypatch):
    """A tree that does not fit k_commit's shared-memory mirror runs the same replay on the global arrays: force that path
    with a 1 KB budget (decision records and RNG window go to global memory too) and compare with the oracle."""
    monkeypatch.setenv("OCTA_COMMIT_SMEM", "1024")
    from octa_autosegmentation_b200 import growth as gmod
    compare_with_oracle(gmod, small_config(), [0, 1, 2])


def test_fixed_geometry_vs_reference_csv(growth):
    """f-4: SimulationSpace.oxygen_sample_geometry_path.  Goldens written by the unmodified reference with the synthetic mask
    tests/golden/geometry_mask.npy (oracle/make_golden.py)."""
    cfg = small_config()
    cfg["Greenhouse"]["SimulationSpace"]["oxygen_sample_geometry_path"] = os.path.join(GOLDEN, "geometry_mask.npy")
    graphs, _ = compare_with_oracle(growth, cfg, [0, 1, 2, 3])
    for seed in (0, 1):
        got = numpy_csv(np.concatenate(graphs[seed]))
        assert got == open(os.path.join(GOLDEN, "graph_geom_s%d.csv" % seed), "rb").read()


def test_3d_geometry_and_z_walls_vs_reference_csv(growth, tmp_path):
    """f-4: a 3-D sampling mask ([40, 84, 8], geometry_size 84) with trees rooted on x0, y0, y1, z0 and z1 -- goldens written by
    the unmodified reference (oracle/make_golden.py geom3d_config) -- and, against the oracle, a thick slab with trees on all six walls in which the
    ball queries and the cKDTree order are genuinely three-dimensional."""
    from test_oracle_growth import geom3d_config
    graphs, _ = compare_with_oracle(growth, geom3d_config(), [0, 1, 2, 3])
    for seed in (0, 1):
        got = numpy_csv(np.concatenate(graphs[seed]))
        assert got == open(os.path.join(GOLDEN, "graph_geom3d_s%d.csv" % seed), "rb").read()
    g = np.ones((30, 30, 5), dtype=bool)      # a slab 1/6 as thick as it is wide (the default space: 1/76)
    ii, jj, kk = np.ogrid[:30, :30, :5]
    g &= (ii - 15) ** 2 + (jj - 14) ** 2 + (kk - 2) ** 2 > (30 / 7) ** 2
    np.save(tmp_path / "cube.npy", g)
    cfg = small_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (30, 20)):
        m["I"], m["N"] = i, 300
    cfg["Greenhouse"]["SimulationSpace"]["oxygen_sample_geometry_path"] = str(tmp_path / "cube.npy")
    cfg["Forest"]["source_walls"] = {"x0": True, "x1": True, "y0": True, "y1": True, "z0": True, "z1": True}
    compare_with_oracle(growth, cfg, [0, 1, 2])


def test_final_sink_lists_and_trace_vs_oracle(growth):
    """octa_grow_sinks + trace: what Greenhouse.save_stats plots (greenhouse.py:401-441).  The oracle's lists equal the
    reference's (tests/test_oracle_growth.py::test_save_stats_data_vs_reference); here the GPU against the oracle, list order
    included, for the small config and a longer run inside the fixed sampling geometry."""
    from oracle import growth_oracle as go
    geom = small_config()
    geom["Greenhouse"]["SimulationSpace"]["oxygen_sample_geometry_path"] = os.path.join(GOLDEN, "geometry_mask.npy")
    for m, i in zip(geom["Greenhouse"]["modes"], (30, 25)):
        m["I"], m["N"] = i, 600
    for cfg in (small_config(), geom):
        ctx = growth.GrowContext(cfg, 3)
        try:
            _, stats, extra = ctx.run([0, 1, 2], trace=True)
            for i, seed in enumerate((0, 1, 2)):
                tr = []
                go.run(cfg, seed, trace=lambda t, a, o, v, c, pd, nd: tr.append((a, o, v, c)))
                oxy, co2 = go.last_sinks()
                got_oxy, got_co2 = ctx.sinks(i)
                assert np.array_equal(got_oxy, oxy) and np.array_equal(got_co2, co2), seed
                assert np.array_equal(extra["trace"][i], np.array(tr)), seed
            with pytest.raises(Exception):
                ctx.sinks(3)
        finally:
            ctx.close()


def test_docker_config_24_seeds_vs_oracle(growth):
    """the GPU equals the exact-order oracle on docker-config seeds 0-23 (the oracle is byte-identical to the
    reference on the committed goldens).  Runs with the default on-demand cKDTree order (OCTA_BALL_ORDER unset)."""
    from octa_autosegmentation_b200.config import default_config
    graphs, stats = compare_with_oracle(growth, default_config(), list(range(4, 24)))
    # the on-demand path must actually have been exercised: some graph-iterations needed the exact order, most did not
    need = [s["replay_detail"][0] for s in stats]
    assert all(0 < n < 200 for n in need), need


def test_graph_replay_equals_stream_launches(monkeypatch):
    """OCTA_GROW_GRAPH=1: the growth loop is issued as stream launches for a context's first batch and as ONE captured CUDA
    graph from the second batch on (octa_grow_host.cu): same seeds -> same bytes, whichever way the loop was issued."""
    import subprocess
    import sys
    from conftest import ROOT
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')\n"
            "import test_growth_gpu as t\nfrom octa_autosegmentation_b200 import growth\nt._graph_replay(growth)\nprint('GRAPH_OK')\n") % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, OCTA_GROW_GRAPH="1"))
    assert r.returncode == 0 and "GRAPH_OK" in r.stdout, r.stdout + r.stderr


def _graph_replay(growth):
    cfg = small_config()
    ctx = growth.GrowContext(cfg, 4)
    try:
        first, _, _ = ctx.run([0, 1, 2, 3])          # stream launches (no node-count history yet)
        other, _, _ = ctx.run([7, 6, 5, 4])          # captures the graph
        again, _, _ = ctx.run([0, 1, 2, 3])          # re-launches it
        part, _, _ = ctx.run([2, 3])                 # smaller batch: re-captured
    finally:
        ctx.close()
    one_shot, _, _ = growth.grow_batch(cfg, [4, 5, 6, 7])
    for i in range(4):
        assert np.array_equal(first[i][0], again[i][0]) and np.array_equal(first[i][1], again[i][1])
        assert np.array_equal(other[3 - i][0], one_shot[i][0]) and np.array_equal(other[3 - i][1], one_shot[i][1])
    for i in range(2):
        assert np.array_equal(part[i][0], first[2 + i][0]) and np.array_equal(part[i][1], first[2 + i][1])


@pytest.mark.parametrize("mode", ["always", "graph2"])
def test_ball_order_modes_agree(monkeypatch, mode):
    """OCTA_BALL_ORDER=always builds the cKDTree permutation for every graph in every iteration (round-1 behaviour); the
    default builds it only for the graph-iterations whose CO2 order could depend on it.  Both must give the oracle's bytes.
    `graph2`: CUDA graph from the very first batch (OCTA_GROW_GRAPH=2)."""
    if mode == "always":
        monkeypatch.setenv("OCTA_BALL_ORDER", "always")
    else:
        monkeypatch.setenv("OCTA_GROW_GRAPH", "2")
    import subprocess
    import sys
    from conftest import ROOT
    # the switches are read once per process: run the comparison in a child
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')\n"
            "import test_growth_gpu as t\n"
            "from octa_autosegmentation_b200 import growth\n"
            "from octa_autosegmentation_b200.config import default_config\n"
            "t.compare_with_oracle(growth, default_config(), [0, 5])\n"
            "t.compare_with_oracle(growth, t.small_config(), [0, 1, 2])\nprint('MODES_OK')\n") % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ))
    assert r.returncode == 0 and "MODES_OK" in r.stdout, r.stdout + r.stderr


def test_repeated_mode_name_vs_oracle(growth):
    """A later mode that reuses the first mode's name keeps the parameters of the mode initialised last (greenhouse.py:84-85);
    the oracle is pinned to the reference on this config by tests/test_oracle_growth.py."""
    from test_oracle_growth import repeated_name_config
    compare_with_oracle(growth, repeated_name_config(), [3, 4])


def test_kill_scan_fallback_equals_hit_list_path(monkeypatch):
    """k_kill keeps the hit positions of a call as a sorted list (<= 4096); a longer list falls back to block scans over the sink
    list.  OCTA_KILL_RCAP=2 sends nearly every call down the fallback: same graphs, byte for byte."""
    from octa_autosegmentation_b200 import graph_io, growth
    from octa_autosegmentation_b200.config import default_config
    cfg = default_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (30, 40)):
        m["I"] = i
    seeds = [3, 4, 5, 6]
    want, _, _ = growth.grow_batch(cfg, seeds)
    monkeypatch.setenv("OCTA_KILL_RCAP", "2")
    got, _, _ = growth.grow_batch(cfg, seeds)
    for a, b in zip(want, got):
        assert graph_io.csv_bytes(np.concatenate(a)) == graph_io.csv_bytes(np.concatenate(b))


def test_split_kd_build_equals_one_kernel_build(monkeypatch):
    """OCTA_KD_SPLIT=1: top levels of the cKDTree permutation in one CTA, the four subtrees below them in a CTA each (lower latency
    of a single loop).  Same graphs as the one-kernel build, byte for byte (the growth contexts read the switch per process, so the
    split runs in a child process)."""
    import subprocess, sys, os, hashlib
    from octa_autosegmentation_b200 import graph_io, growth
    from octa_autosegmentation_b200.config import default_config
    cfg = default_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (40, 50)):
        m["I"] = i
    seeds = [11, 12, 13, 14, 15, 16]
    want, stats, _ = growth.grow_batch(cfg, seeds)
    assert sum(s["replay_detail"][0] for s in stats) > 0            # the exact order was needed somewhere
    digest = hashlib.sha256(b"".join(graph_io.csv_bytes(np.concatenate(g)) for g in want)).hexdigest()
    code = ("import sys, hashlib, numpy as np; sys.path.insert(0, %r)\n"
            "from octa_autosegmentation_b200 import graph_io, growth\n"
            "from octa_autosegmentation_b200.config import default_config\n"
            "cfg = default_config()\n"
            "for m, i in zip(cfg['Greenhouse']['modes'], (40, 50)): m['I'] = i\n"
            "g, _, _ = growth.grow_batch(cfg, %r)\n"
            "print(hashlib.sha256(b''.join(graph_io.csv_bytes(np.concatenate(x)) for x in g)).hexdigest())\n"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), seeds))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, OCTA_KD_SPLIT="1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1] == digest
