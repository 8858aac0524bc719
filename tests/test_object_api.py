"""CPU: the reference's object API over the engine (octa_autosegmentation_b200/vessel_graph_generation/: Greenhouse, Forest,
ArterialTree, Node, SimulationSpace -- SURVEY 8b "Python API used by main").  The engine call is replaced by the CPU oracle here
(tests may use it), so what is checked is the facade: attributes, tree reconstruction from the edge tables, iteration order,
per-step lists, sink lists.  tests/test_zz_object_api_gpu.py runs the same script on the real engine."""
import csv
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN
from octa_autosegmentation_b200 import growth
from octa_autosegmentation_b200.config import default_config


class OracleContext:
    """Stand-in for growth.GrowContext backed by oracle/growth_oracle (same results by the GPU parity tests)."""

    def __init__(self, config, max_graphs, *a, **k):
        self.config = config

    def run(self, seeds, trace=False, copy=True):
        from oracle import growth_oracle as go
        assert len(seeds) == 1
        tr = []
        art, ven, st = go.run(self.config, seeds[0], trace=lambda t, a, o, v, c, pd, nd: tr.append((a, o, v, c)))
        self._sinks = go.last_sinks()
        return [(art, ven)], [st], {"device_ms": 12.0, "trace": np.array(tr, dtype=np.int32)[None]}

    def sinks(self, i):
        return self._sinks

    def close(self):
        pass


def small_config():
    cfg = default_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (12, 12)):
        m["I"], m["N"] = i, 400
    return cfg


def reference_main_csv(config, seed):
    """generate_vessel_graph.py:24-66, verbatim in structure, on the facade classes."""
    from octa_autosegmentation_b200.vessel_graph_generation.forest import Forest
    from octa_autosegmentation_b200.vessel_graph_generation.greenhouse import Greenhouse
    greenhouse = Greenhouse(config["Greenhouse"], seed=seed)
    arterial_forest = Forest(config["Forest"], greenhouse.d, greenhouse.r, greenhouse.simspace, nerve_center=greenhouse.nerve_center,
                             nerve_radius=greenhouse.nerve_radius)
    venous_forest = Forest(config["Forest"], greenhouse.d, greenhouse.r, greenhouse.simspace, arterial=False,
                           nerve_center=greenhouse.nerve_center, nerve_radius=greenhouse.nerve_radius)
    greenhouse.set_forests(arterial_forest, venous_forest)
    greenhouse.develop_forest()
    art_edges = [{"node1": n.position, "node2": n.get_proximal_node().position, "radius": n.radius}
                 for tree in arterial_forest.get_trees() for n in tree.get_tree_iterator(exclude_root=True, only_active=False)]
    ven_edges = [{"node1": n.position, "node2": n.get_proximal_node().position, "radius": n.radius}
                 for tree in venous_forest.get_trees() for n in tree.get_tree_iterator(exclude_root=True, only_active=False)]
    buf = io.StringIO(newline="")
    writer = csv.writer(buf)
    writer.writerow(["node1", "node2", "radius"])
    for row in art_edges + ven_edges:
        writer.writerow([row["node1"], row["node2"], row["radius"]])
    return buf.getvalue().encode(), greenhouse, arterial_forest, venous_forest, art_edges, ven_edges


def check_against_goldens(data, greenhouse, art, ven):
    assert data == open(os.path.join(GOLDEN, "graph_small_s0.csv"), "rb").read()
    gold = np.load(os.path.join(GOLDEN, "stats_small_s0.npz"))
    got = np.stack([greenhouse.art_nodes_per_step, greenhouse.oxys_per_step, greenhouse.ven_nodes_per_step, greenhouse.co2_per_step], 1)
    assert np.array_equal(got, gold["per_step"]) and len(greenhouse.time_per_step) == len(got) - 1
    assert np.array_equal(np.array(greenhouse.oxy_mesh.get_all_elements()), gold["oxys"])
    assert np.array_equal(np.array(greenhouse.co2_mesh.get_all_elements()), gold["co2s"])
    cfg = small_config()
    assert len(art.get_trees()) == len(ven.get_trees()) == cfg["Forest"]["N_trees"]
    assert [t.name for t in art.get_trees()][:2] == ["ArterialTree1", "ArterialTree2"] and ven.get_trees()[0].name == "VenousTree1"
    # node counts of the last step = nodes of the forests (roots included)
    assert sum(1 for _ in art.get_nodes()) == got[-1, 0] and sum(1 for _ in ven.get_nodes()) == got[-1, 2]
    for tree in art.get_trees():
        assert tree.root.is_root and tree.root.position[0] in (0.0, 1 - 1e-6) or tree.root.position[1] in (0.0, 1 - 1e-6)
        for n in tree.get_tree_iterator(exclude_root=True):
            assert n.get_proximal_node() is n.parent and n in n.parent.children and len(n.children) <= 2
            assert n.proximal_num_segments == n.parent.proximal_num_segments + 1
    with pytest.raises(RuntimeError):
        art.get_trees()[0].root.get_proximal_node()


def test_reference_main_on_the_object_api(monkeypatch, tmp_path):
    monkeypatch.setattr(growth, "GrowContext", OracleContext)
    data, greenhouse, art, ven, art_edges, ven_edges = reference_main_csv(small_config(), 0)
    check_against_goldens(data, greenhouse, art, ven)
    assert greenhouse.d == 0.1 / 3 and greenhouse.r == 0.0025 / 3 and tuple(greenhouse.simspace.shape) == (1, 1, 0.0131)
    # Forest.save writes the forest's own CSV (forest.py:196-207); save_stats the four plots
    art.save(str(tmp_path))
    rows = list(csv.reader(open(tmp_path / "ArterialForest.csv", newline="")))
    assert len(rows) == 1 + len(art_edges) and rows[1][0] == str(art_edges[0]["node1"])
    greenhouse.save_stats(str(tmp_path))
    assert all(os.path.exists(tmp_path / (n + ".png")) for n in ("oxy_distribution", "co2_distribution", "time_per_step", "growth_over_time"))
    # only_active = inside the simulation space and outside the FAZ test of simulation_space.py:89-98
    n_act = sum(1 for t in art.get_trees() for _ in t.get_tree_iterator(only_active=True))
    n_all = sum(1 for _ in art.get_nodes())
    assert 0 < n_act <= n_all


def test_unseeded_greenhouse_draws_its_seed_from_python_random(monkeypatch):
    import random
    from octa_autosegmentation_b200.vessel_graph_generation.greenhouse import Greenhouse
    cfg = small_config()
    random.seed(5)
    a = Greenhouse(cfg["Greenhouse"])
    random.seed(5)
    b = Greenhouse(cfg["Greenhouse"])
    assert a.seed == b.seed and a.FAZ_radius == b.FAZ_radius
    g = Greenhouse(cfg["Greenhouse"], seed=0)
    with pytest.raises(RuntimeError):
        g.develop_forest()                      # no forests yet
    from octa_autosegmentation_b200.vessel_graph_generation.forest import Forest
    with pytest.raises(NotImplementedError):
        Forest({"type": "bushes"}, g.d, g.r, g.simspace)
    g.set_forests(Forest(cfg["Forest"], g.d, g.r, g.simspace))
    with pytest.raises(NotImplementedError):
        g.develop_forest()                      # arterial forest alone: not offered by the engine


@pytest.mark.skipif(not os.path.isdir("/root/reference/vessel_graph_generation"), reason="needs /root/reference (build container)")
@pytest.mark.parametrize("kind", ["stumps", "nerve"])
def test_object_api_against_the_reference_objects(monkeypatch, kind):
    """Attributes and iteration of the facade against the unmodified reference's own objects after the same seeded run."""
    from oracle import ref_harness as rh
    monkeypatch.setattr(growth, "GrowContext", OracleContext)
    cfg = small_config() if kind == "stumps" else rh.nerve_config(I=(30, 20), N=800)
    _, gh, art, ven, _, _ = reference_main_csv(cfg, 1)
    _, _, rgh = rh.run_growth(cfg, 1)
    # (the reference rescales its own .d while it grows, greenhouse.py:139-147; at construction both are config d / param_scale)
    assert gh.FAZ_radius == rgh.FAZ_radius and gh.r == rgh.r and gh.rotation_radius == rgh.rotation_radius
    assert np.array_equal(gh.nerve_center, rgh.nerve_center) and np.array_equal(gh.simspace.shape, rgh.simspace.shape)
    assert np.array_equal(gh.simspace.geometry, rgh.simspace.geometry) and np.array_equal(gh.simspace.valid_voxels, rgh.simspace.valid_voxels)
    for mine, ref in ((art, rgh.arterial_forest), (ven, rgh.venous_forest)):
        assert len(mine.get_trees()) == len(ref.get_trees())
        for tm, tr in zip(mine.get_trees(), ref.get_trees()):
            assert tm.name == tr.name
            for only_active in (False, True):
                nm = list(tm.get_tree_iterator(exclude_root=False, only_active=only_active))
                nr = list(tr.get_tree_iterator(exclude_root=False, only_active=only_active))
                assert len(nm) == len(nr)
                for a, b in zip(nm, nr):
                    assert np.array_equal(a.position, b.position) and a.active == b.active and a.is_leaf == b.is_leaf
                    assert a.is_inter_node == b.is_inter_node and a.is_bifurcation_node == b.is_bifurcation_node
                    assert a.proximal_num_segments == b.proximal_num_segments
                    if not a.is_root:
                        assert a.radius == b.radius
    assert gh.art_nodes_per_step == rgh.art_nodes_per_step and gh.co2_per_step == rgh.co2_per_step
