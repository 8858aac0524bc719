"""Generates the fixtures under tests/golden/ from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py [--full]

  graph_small_s{0,1}.csv      seeded growth, docker config cut to I=(12,12), N=400     (ref_harness.run_growth)
  graph_docker_s0.csv.gz      seeded growth, docker config verbatim, seed 0  (needs --full, ~100 s)
  graph_geom_s{0,1}.csv       the same with SimulationSpace.oxygen_sample_geometry_path = tests/golden/geometry_mask.npy
                              (a synthetic 76x76x1 mask written by geometry_mask(); fixed-geometry branch of
                              simulation_space.py:26-34,69-76,95-96)
  stats_small_s0.npz          the data Greenhouse.save_stats plots for graph_small_s0: oxy_mesh / co2_mesh .get_all_elements()
                              and the *_per_step lists (incl. their leading 0 entry)
  graph_geom3d_s{0,1}.csv     a 3-D mask (tests/golden/geometry_mask_3d.npy, [40, 84, 8]) with source walls x0, y0, y1, z0, z1
                              (geometry_mask_3d(), geom3d_config(): 3-D argwhere sampling, z walls of forest.py:152-176)
  vox_small_s0_*.npz          tree2img.voxelize_forest of graph_small_s0.csv for several requests
  vox_docker_s0.json          sha256 / non-zero count of voxelize_forest(graph_docker_s0, [304,304,4]) and,
                              with --full, of the [1216,1216,16] request (77 s in the reference)
  graph_nerve_s0.csv          Forest.type 'nerve', 12x12 mm^2 variant (ref_harness.nerve_config, I = 70 + 50, N = 1500)
  graph_docker_s1.csv.gz, graph_docker_digests.json   (--full) docker config seeds 0-3: bytes + sha256 of the reference's CSV
  shipped_<name>.csv.gz, shipped_<name>_label.npz     8 of the 500 (graph csv -> 1216^2 1-bit label) pairs the reference ships
  r2d_small_s0.npz            tree2img.rasterize_forest of the unmodified reference (matplotlib calls served by
                              oracle/shims/matplotlib -> oracle/agg_oracle.c) for several option sets
"""
import argparse
import gzip
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def digest(a: np.ndarray) -> dict:
    return {"shape": list(a.shape), "dtype": str(a.dtype), "nonzero": int((a > 0).sum()), "sum": int(a.astype(np.int64).sum()),
            "sha256": hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()}


def geometry_mask() -> np.ndarray:
    """Synthetic oxygen-sample geometry (bool [76, 76, 1]): a hole, a blocked bar and partly blocked wall planes."""
    g = np.ones((76, 76, 1), dtype=bool)
    yy, xx = np.ogrid[:76, :76]
    g[:, :, 0] &= (xx - 38) ** 2 + (yy - 37) ** 2 > 6.5 ** 2
    g[10:13, 58:71, 0] = False
    g[0, :7, 0] = False          # wall plane used by x0 / x1
    g[:5, 0, 0] = False          # wall plane used by y0 / y1
    return g


def geometry_mask_3d() -> np.ndarray:
    """Synthetic 3-D sampling geometry (bool [40, 84, 8], geometry_size 84 > the default 76, x and z shorter than y): an
    ellipsoidal hole, a blocked slab and partly blocked x / y / z wall planes."""
    g = np.ones((40, 84, 8), dtype=bool)
    ii, jj, kk = np.ogrid[:40, :84, :8]
    g &= (ii - 20) ** 2 + (jj - 45) ** 2 + 4 * (kk - 4) ** 2 > 6.0 ** 2
    g[30:34, 10:30, 2:] = False
    g[0, :9, :] = False          # wall plane used by x0 / x1
    g[:6, 0, :3] = False         # wall plane used by y0 / y1
    g[:10, :12, 0] = False       # wall plane used by z0 / z1
    return g


def geom3d_config(mask_path: str) -> dict:
    """small_config grown inside the 3-D mask, trees rooted on x0, y0, y1, z0 and z1 (the z walls only work with a geometry file)."""
    cfg = rh.small_config(I=(25, 15), N=500)
    cfg["Greenhouse"]["SimulationSpace"]["oxygen_sample_geometry_path"] = mask_path
    cfg["Forest"]["source_walls"] = {"x0": True, "x1": False, "y0": True, "y1": True, "z0": True, "z1": True}
    cfg["Forest"]["N_trees"] = 6
    return cfg


def shipped_pair(name="20230216_232653"):
    """One of the 500 (graph csv -> 1216^2 1-bit label) pairs the reference ships under datasets/ -- the only pin
    available for the matplotlib/Agg 2-D path.  Stored as csv.gz + bit-packed label."""
    from PIL import Image
    raw = open(os.path.join(rh.REFERENCE_ROOT, "datasets", "vessel_graphs", name + ".csv"), "rb").read()
    with gzip.GzipFile(os.path.join(GOLD, "shipped_%s.csv.gz" % name), "wb", mtime=0) as f:
        f.write(raw)
    lab = np.array(Image.open(os.path.join(rh.REFERENCE_ROOT, "datasets", "labels", name + ".png")))
    np.savez_compressed(os.path.join(GOLD, "shipped_%s_label.npz" % name), packed=np.packbits(lab), shape=np.array(lab.shape))


SHIPPED = ["20230216_232653", "20230216_235406", "20230217_010748", "20230217_015939", "20230217_031216", "20230217_043443",
           "20230217_050150", "20230217_060539"]      # incl. every kind of near-axis-aligned (snapped) stroke found in the set


def raster_goldens():
    """tree2img.rasterize_forest of the UNMODIFIED reference (through oracle/shims/matplotlib -> oracle/agg_oracle.c) for
    graph_small_s0.csv: resolutions, MIP axes, radius filter, subtree dropout with a seeded Python RNG, shared blackdict."""
    import random
    rows = rh.read_csv_rows(os.path.join(GOLD, "graph_small_s0.csv"))
    out = {}
    out["a_304x304_mip2"], _ = rh.rasterize(rows, [304, 304], 2)
    out["b_200x120_mip0_minr"], _ = rh.rasterize(rows, [200, 120], 0, min_radius=0.001)
    out["c_96x160_mip1_maxr"], _ = rh.rasterize(rows, [96, 160], 1, max_radius=0.002)
    random.seed(153)
    rl = []
    out["d_304x304_dropout"], bd = rh.rasterize(rows, [304, 304], 2, radius_list=rl, max_dropout_prob=0.3)
    out["d_next_random"] = np.array([random.random()])
    out["d_radius_list"] = np.array(rl)
    out["d_blackdict"] = np.array(sorted(bd.keys()))
    out["e_1216x1216_blackdict"], _ = rh.rasterize(rows, [1216, 1216], 2, min_radius=0.0009, blackdict=bd)
    np.savez_compressed(os.path.join(GOLD, "r2d_small_s0.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    for seed in (0, 1):
        art, ven, _ = rh.run_growth(rh.small_config(), seed)
        with open(os.path.join(GOLD, "graph_small_s%d.csv" % seed), "wb") as f:
            f.write(rh.csv_bytes(art, ven))
    # what Greenhouse.save_stats plots (greenhouse.py:401-441) for graph_small_s0: final sink lists, per-iteration counts
    _, _, gh = rh.run_growth(rh.small_config(), 0)
    np.savez_compressed(os.path.join(GOLD, "stats_small_s0.npz"), oxys=np.array(gh.oxy_mesh.get_all_elements()),
                        co2s=np.array(gh.co2_mesh.get_all_elements()),
                        per_step=np.stack([gh.art_nodes_per_step, gh.oxys_per_step, gh.ven_nodes_per_step, gh.co2_per_step], 1))
    mask_path = os.path.join(GOLD, "geometry_mask.npy")
    np.save(mask_path, geometry_mask())
    for seed in (0, 1):
        cfg = rh.small_config()
        cfg["Greenhouse"]["SimulationSpace"]["oxygen_sample_geometry_path"] = mask_path
        art, ven, _ = rh.run_growth(cfg, seed)
        with open(os.path.join(GOLD, "graph_geom_s%d.csv" % seed), "wb") as f:
            f.write(rh.csv_bytes(art, ven))
    mask3_path = os.path.join(GOLD, "geometry_mask_3d.npy")
    np.save(mask3_path, geometry_mask_3d())
    for seed in (0, 1):
        art, ven, _ = rh.run_growth(geom3d_config(mask3_path), seed)
        with open(os.path.join(GOLD, "graph_geom3d_s%d.csv" % seed), "wb") as f:
            f.write(rh.csv_bytes(art, ven))
    rows = rh.read_csv_rows(os.path.join(GOLD, "graph_small_s0.csv"))
    cases = {"304x304x4": ([304, 304, 4], {}), "304x304x4_ignz": ([304, 304, 4], {"ignore_z": True}),
             "100x80x30_minr": ([100, 80, 30], {"min_radius": 0.001}), "64x64x64": ([64, 64, 64], {}),
             "96x96x130": ([96, 96, 130], {})}
    for name, (dims, kw) in cases.items():
        vol, _ = rh.voxelize(rows, dims, **kw)
        np.savez_compressed(os.path.join(GOLD, "vox_small_s0_%s.npz" % name), vol=vol, dims=np.array(dims),
                            kw=json.dumps(kw))
    art, ven, _ = rh.run_growth(rh.nerve_config(I=(70, 50), N=1500), 0)
    with open(os.path.join(GOLD, "graph_nerve_s0.csv"), "wb") as f:
        f.write(rh.csv_bytes(art, ven))
    for name in SHIPPED:
        shipped_pair(name)
    raster_goldens()
    if a.full:
        # docker config verbatim: ~100 s per seed in the reference.  Seeds 0 and 1 are stored, 0-3 pinned by sha256.
        dig = {}
        for seed in range(4):
            art, ven, _ = rh.run_growth(rh.load_config(), seed)
            data = rh.csv_bytes(art, ven)
            dig["docker_s%d" % seed] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest(), "rows": len(art) + len(ven)}
            if seed < 2:
                with gzip.GzipFile(os.path.join(GOLD, "graph_docker_s%d.csv.gz" % seed), "wb", mtime=0) as f:
                    f.write(data)
        with open(os.path.join(GOLD, "graph_docker_digests.json"), "w") as f:
            json.dump(dig, f, indent=1)
        p = os.path.join(GOLD, "graph_docker_s0.csv.gz")
        import csv, io
        rows = list(csv.DictReader(io.StringIO(gzip.open(p, "rt", newline="").read(), newline="")))
        out = {}
        for dims in ([304, 304, 4], [1216, 1216, 16]):
            vol, _ = rh.voxelize(rows, dims)
            out["x".join(map(str, dims))] = digest(vol)
            print(dims, out["x".join(map(str, dims))])
        with open(os.path.join(GOLD, "vox_docker_s0.json"), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

