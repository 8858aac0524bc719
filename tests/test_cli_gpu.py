"""GPU: the two drop-in CLIs write the reference's files (SURVEY 8 rows a-L, f-2) -- names, dtypes and BITS checked against the
oracles: visualize_vessel_graphs.py:69-104 (`_label.png` through PIL convert("1"), `.npy` bool volumes, --ignore_z, subtree
dropout + blackdict pickle, natural sort, --num_samples) and generate_vessel_graph.py:59-86 (csv bytes, art/ven max image, npy
volume)."""
import glob
import os
import pickle
import random
import shutil

import numpy as np
import pytest
import yaml

from conftest import GOLDEN, load_graph_rows, rows_to_edges7

pytestmark = pytest.mark.gpu


@pytest.fixture()
def source_dir(tmp_path):
    d = tmp_path / "graphs" / "nested"
    d.mkdir(parents=True)
    # natural order: 1.csv, 2.csv, 10.csv (a plain sort would put 10 before 2)
    shutil.copy(os.path.join(GOLDEN, "graph_small_s0.csv"), d / "2.csv")
    shutil.copy(os.path.join(GOLDEN, "graph_small_s1.csv"), d / "10.csv")
    shutil.copy(os.path.join(GOLDEN, "graph_geom_s0.csv"), d / "1.csv")
    return tmp_path / "graphs"


def rows_of(path):
    import csv
    with open(path, newline="") as f:
        return list(csv.DictReader(f))


def test_visualize_cli_labels_and_bool_volumes(source_dir, tmp_path):
    from PIL import Image
    from octa_autosegmentation_b200 import visualize_vessel_graphs as cli
    from oracle import agg_oracle, vox_oracle
    out = tmp_path / "out"
    random.seed(2)                                   # first draw 0.956: p = 0.956**10 * 0.3 = 0.19
    assert cli.main(["--source_dir", str(source_dir), "--out_dir", str(out), "--resolution", "304,304,4", "--save_3d",
                     "--save_3d_as", ".npy", "--binarize", "--max_dropout_prob", "0.3", "--num_samples", "2", "--batch", "2"]) == 0
    # --num_samples 2 after the natural sort -> 1.csv and 2.csv; with --save_3d every later name carries the 3-D suffix
    assert sorted(os.listdir(out)) == sorted(["1_3d_label.npy", "1_3d_label_blackdict.pkl", "1_3d_label_label.png",
                                              "2_3d_label.npy", "2_3d_label_blackdict.pkl", "2_3d_label_label.png"])
    random.seed(2)                                   # the reference in one process: per file voxelize_forest, then rasterize_forest
    dropped = 0
    for name in ("1", "2"):
        rows = rows_of(source_dir / "nested" / (name + ".csv"))
        vol, bd = vox_oracle.voxelize_forest(rows, [304, 304, 4], max_dropout_prob=0.3)
        img, _ = agg_oracle.rasterize_forest(rows, [304, 304], 2)            # no dropout on the 2-D image (:95), one draw for p
        got = np.load(out / (name + "_3d_label.npy"))
        assert got.dtype == np.bool_ and got.shape == vol.shape and np.array_equal(got, vol >= 0.1)
        assert pickle.load(open(out / (name + "_3d_label_blackdict.pkl"), "rb")) == bd
        dropped += len(bd)
        lab = Image.open(out / (name + "_3d_label_label.png"))
        assert lab.mode == "1" and np.array_equal(np.array(lab), agg_oracle.to_label(img.astype(np.uint8)))
    assert dropped > 0          # the dropout really removed subtrees (p of the first file = 0.19)


def test_visualize_cli_gray_images_ignore_z_and_2d_only(source_dir, tmp_path):
    from PIL import Image
    from octa_autosegmentation_b200 import visualize_vessel_graphs as cli
    from oracle import agg_oracle, vox_oracle
    out = tmp_path / "o2"
    assert cli.main(["--source_dir", str(source_dir), "--out_dir", str(out), "--resolution", "96,160,8", "--mip_axis", "0",
                     "--save_3d", "--save_3d_as", ".npy", "--ignore_z"]) == 0
    assert sorted(os.listdir(out)) == sorted(["%s_3d%s" % (n, e) for n in ("1", "2", "10") for e in (".npy", ".png")])
    rows = rows_of(source_dir / "nested" / "10.csv")
    vol, _ = vox_oracle.voxelize_forest(rows, [96, 160, 8], ignore_z=True)
    assert np.array_equal(np.load(out / "10_3d.npy"), vol.astype(np.bool_))                 # .npy is ALWAYS written as bool (:94)
    img = np.array(Image.open(out / "10_3d.png"))
    ref, _ = agg_oracle.rasterize_forest(rows, [160, 8], 0)                                  # resolution with the MIP axis removed (:57-59)
    assert img.dtype == np.uint8 and np.array_equal(img, ref.astype(np.uint8))
    # 2-D only (the default): plain names, label PNG
    out3 = tmp_path / "o3"
    assert cli.main(["--source_dir", str(source_dir), "--out_dir", str(out3), "--resolution", "1216,1216,16", "--binarize", "--num_samples", "1"]) == 0
    assert os.listdir(out3) == ["1_label.png"]
    rows = rows_of(source_dir / "nested" / "1.csv")
    ref, _ = agg_oracle.rasterize_forest(rows, [1216, 1216], 2)
    assert np.array_equal(np.array(Image.open(out3 / "1_label.png")), agg_oracle.to_label(ref.astype(np.uint8)))
    with pytest.raises(AssertionError):
        cli.main(["--source_dir", str(source_dir), "--out_dir", str(out3), "--no_save_2d"])
    # the DEFAULT 3-D container is NIfTI (:39, :85-87): uint16 volume, binarised to 0 / 1 with --binarize, identity affine
    from octa_autosegmentation_b200 import graph_io
    out4 = tmp_path / "o4"
    assert cli.main(["--source_dir", str(source_dir), "--out_dir", str(out4), "--resolution", "64,64,4", "--save_3d", "--no_save_2d",
                     "--num_samples", "1"]) == 0
    assert cli.main(["--source_dir", str(source_dir), "--out_dir", str(out4), "--resolution", "64,64,4", "--save_3d", "--no_save_2d",
                     "--num_samples", "1", "--binarize"]) == 0
    assert sorted(os.listdir(out4)) == ["1_3d.nii.gz", "1_3d_label.nii.gz"]
    vol, _ = vox_oracle.voxelize_forest(rows, [64, 64, 4])
    got = graph_io.load_nifti(str(out4 / "1_3d.nii.gz"))
    assert got.dtype == np.uint16 and np.array_equal(got, vol)
    assert np.array_equal(graph_io.load_nifti(str(out4 / "1_3d_label.nii.gz")), (vol >= 0.1).astype(np.uint16))


def small_cfg(tmp_path, **out_kw):
    from octa_autosegmentation_b200.config import default_config
    cfg = default_config()
    for m, i in zip(cfg["Greenhouse"]["modes"], (12, 12)):
        m["I"], m["N"] = i, 400
    cfg["output"].update(directory=str(tmp_path / "gen"), save_trees=True, save_2D_image=True, save_3D_volumes=None, save_stats=False)
    cfg["output"].update(out_kw)
    yml = tmp_path / "cfg.yml"
    yml.write_text(yaml.dump(cfg))
    return cfg, yml


def test_generate_cli_writes_reference_files(tmp_path):
    """generate_vessel_graph.py surface: YAML in, one timestamped folder per sample with config.yml, <name>.csv,
    art_ven_img_gray.png (np.maximum of the arterial and the venous raster, :80-86) and the optional uint8 volume (:69-75)."""
    from PIL import Image
    from octa_autosegmentation_b200 import generate_vessel_graph as cli
    from oracle import agg_oracle, growth_oracle, vox_oracle
    cfg, yml = small_cfg(tmp_path, save_3D_volumes="npy", image_scale_factor=152)
    assert cli.main(["--config_file", str(yml), "--num_samples", "3", "--seed", "0", "--batch", "2", "--in_flight", "2"]) == 0
    dirs = sorted(glob.glob(os.path.join(cfg["output"]["directory"], "*")))
    assert len(dirs) == 3
    by_csv = {}
    for d in dirs:
        name = os.path.basename(d)
        assert yaml.safe_load(open(os.path.join(d, "config.yml"))) == cfg
        by_csv[open(os.path.join(d, name + ".csv"), "rb").read()] = d
    dims = [152, 152, 1]                                     # int(shape * image_scale_factor), shape = (1, 1, 0.0131)
    for seed in (0, 1, 2):
        oa, ov, _ = growth_oracle.run(cfg, seed)
        data = growth_oracle.csv_bytes(np.concatenate([oa, ov]))
        assert data in by_csv, seed
        if seed < 2:
            assert data == open(os.path.join(GOLDEN, "graph_small_s%d.csv" % seed), "rb").read()
        d = by_csv[data]
        img = np.asarray(Image.open(os.path.join(d, "art_ven_img_gray.png")))
        ref = np.maximum(agg_oracle.raster_edges(oa, [152, 152]), agg_oracle.raster_edges(ov, [152, 152]))
        assert img.dtype == np.uint8 and np.array_equal(img, ref)
        vol = np.load(os.path.join(d, "art_ven_img_gray.npy"))
        want = np.maximum(vox_oracle.voxelize_edges(oa, dims), vox_oracle.voxelize_edges(ov, dims)).astype(np.uint8)
        assert vol.dtype == np.uint8 and np.array_equal(vol, want)


def test_generate_cli_save_stats(tmp_path):
    """output.save_stats (on in the reference's main config, configs/dataset_18_June_2023.yml:57): every sample folder also gets
    the five statistic plots of generate_vessel_graph.py:40-41,88-89 -- and the data behind them is the reference's (the sink
    lists and per-iteration counts of graph_small_s0 as the unmodified reference held them, tests/golden/stats_small_s0.npz)."""
    from PIL import Image
    from octa_autosegmentation_b200 import generate_vessel_graph as cli
    from octa_autosegmentation_b200.pipeline import Pipeline
    cfg, yml = small_cfg(tmp_path, save_stats=True, image_scale_factor=152)
    assert cli.main(["--config_file", str(yml), "--num_samples", "2", "--seed", "0", "--batch", "2", "--in_flight", "1"]) == 0
    dirs = sorted(glob.glob(os.path.join(cfg["output"]["directory"], "*")))
    assert len(dirs) == 2
    for d in dirs:
        for name, size in (("oxy_distribution", (600, 600)), ("co2_distribution", (600, 600)), ("time_per_step", (600, 600)),
                           ("growth_over_time", (600, 600)), ("hist", (640, 480))):
            im = Image.open(os.path.join(d, name + ".png"))
            assert im.size == size and len(im.getcolors(1 << 16)) > 2, name
    pipe = Pipeline(cfg, volume_dims=[152, 152, 1], label_res=None, image_res=[152, 152], voxelize=False, growth_stats=True)
    gs = pipe.run([0, 1], d2h=True, csv=False)["growth_stats"]
    gold = np.load(os.path.join(GOLDEN, "stats_small_s0.npz"))
    assert gs["iterations"] == 24 and np.array_equal(gs["per_step"][0], gold["per_step"][1:])
    assert np.array_equal(gs["sinks"][0][0], gold["oxys"]) and np.array_equal(gs["sinks"][0][1], gold["co2s"])
