"""GPU parity tests of K7 (csrc/octa_raster2d.cu) through the C ABI.

The reference's 2-D path (tree2img.py:12-114) is matplotlib/Agg.  oracle/agg_oracle.c restates that pipeline and reproduces
all 500 label PNGs the reference ships bit for bit (tests/test_oracle_raster2d.py); the CUDA kernel evaluates the same integer
scanline arithmetic row-parallel.  Bar: the gray image equals the oracle's PIXEL FOR PIXEL (integer work: bit-exact), the
labels equal the reference's shipped labels bit for bit, and rasterize_forest equals the goldens written by the unmodified
reference code (tests/golden/r2d_small_s0.npz)."""
import os
import pickle
import random
import shutil

import numpy as np
import pytest

from conftest import GOLDEN, load_graph_rows, rows_to_edges7
from test_oracle_raster2d import SHIPPED, hard_cases, shipped_label

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def t2i():
    from octa_autosegmentation_b200 import _lib, tree2img
    assert _lib.lib().octa_device_count() > 0
    return tree2img


def test_labels_equal_the_shipped_reference_labels_bit_for_bit(t2i):
    """visualize_vessel_graphs.py:69-101 for the committed shipped pairs: rasterize_forest at [1216, 1216] -> img<0.1 -> 0 ->
    PIL convert("1") must give the reference's own label PNG (recorded: mismatching pixels per pair, IoU)."""
    from PIL import Image
    report = []
    for name in SHIPPED:
        rows = load_graph_rows("shipped_%s.csv.gz" % name)
        img, bd = t2i.rasterize_forest(rows, [1216, 1216], 2)
        assert img.dtype == np.uint16 and img.shape == (1216, 1216) and bd == {} and img.max() == 255
        img[img < 0.1] = 0
        lab = np.array(Image.fromarray(img.astype(np.uint8)).convert("1"))
        ref = shipped_label(name)
        mism = int((lab != ref).sum())
        iou = float((lab & ref).sum() / (lab | ref).sum())
        report.append((name, mism, iou, float(lab.mean()), float(ref.mean())))
    print("\nlabel parity (name, mismatching pixels, IoU, vessel fraction ours / reference):")
    for r in report:
        print("  %s  %d  %.6f  %.4f / %.4f" % r)
    assert all(r[1] == 0 for r in report), report


def test_gray_image_equals_the_agg_oracle_pixel_for_pixel(t2i):
    from oracle import agg_oracle
    e7 = rows_to_edges7(load_graph_rows("shipped_%s.csv.gz" % SHIPPED[0]))
    for res, mip in (([1216, 1216], 2), ([304, 304], 2), ([200, 120], 0), ([96, 160], 1), ([33, 31], 2)):
        a, b = t2i.raster_edges(e7, res, mip), agg_oracle.raster_edges(e7, res, MIP_axis=mip)
        assert a.shape == b.shape and np.array_equal(a, b), (res, mip, int((a != b).sum()))
    a = t2i.raster_edges(e7, [608, 608], 2, min_radius=0.001, max_radius=0.004)
    assert np.array_equal(a, agg_oracle.raster_edges(e7, [608, 608], min_radius=0.001, max_radius=0.004))
    # strokes leaving the canvas, axis-aligned (snapped) ones, zero-length ones, hair-thin and very thick ones
    for trial in range(6):
        h = hard_cases(trial)
        for res, mip in (([160, 120], 2), ([64, 200], 0), ([333, 333], 1)):
            a, b = t2i.raster_edges(h, res, mip, 0.0003, 0.1), agg_oracle.raster_edges(h, res, MIP_axis=mip, min_radius=0.0003, max_radius=0.1)
            assert np.array_equal(a, b), (trial, res, mip, int((a != b).sum()))
    # more than 1024 strokes in one tile (global-memory ordering path) and an edge spanning more than 64 tiles ("big" list)
    rng = np.random.default_rng(5)
    p = rng.uniform(0.40, 0.46, (1500, 3))
    dense = np.concatenate([p, p + rng.normal(0, 0.01, p.shape), rng.uniform(0.0003, 0.002, (1500, 1))], 1)
    dense[700] = [0.02, 0.03, 0, 0.97, 0.95, 0, 0.004]
    a, b = t2i.raster_edges(dense, [512, 512]), agg_oracle.raster_edges(dense, [512, 512])
    assert np.array_equal(a, b), int((a != b).sum())


def test_rasterize_forest_equals_goldens_of_the_unmodified_reference(t2i):
    """Host loop of tree2img.py:58-86 (radius filter, legacy string rows, subtree dropout, blackdict, radius_list, RNG
    consumption) + the kernel vs tests/golden/r2d_small_s0.npz (written by the reference's rasterize_forest)."""
    rows = load_graph_rows("graph_small_s0.csv")
    gold = np.load(os.path.join(GOLDEN, "r2d_small_s0.npz"))
    for res, mip, kw, key in (([304, 304], 2, {}, "a_304x304_mip2"), ([200, 120], 0, {"min_radius": 0.001}, "b_200x120_mip0_minr"),
                              ([96, 160], 1, {"max_radius": 0.002}, "c_96x160_mip1_maxr")):
        img, bd = t2i.rasterize_forest(rows, res, mip, **kw)
        assert img.dtype == np.uint16 and bd == {} and np.array_equal(img, gold[key]), key
    random.seed(153)
    rl = []
    img, bd = t2i.rasterize_forest(rows, [304, 304], 2, radius_list=rl, max_dropout_prob=0.3)
    assert np.array_equal(img, gold["d_304x304_dropout"]) and random.random() == float(gold["d_next_random"][0])
    assert np.array_equal(np.array(rl), gold["d_radius_list"]) and np.array_equal(np.array(sorted(bd.keys())), gold["d_blackdict"])
    img, bd2 = t2i.rasterize_forest(rows, [1216, 1216], 2, min_radius=0.0009, blackdict=bd)
    assert np.array_equal(img, gold["e_1216x1216_blackdict"]) and bd2 is bd
    # empty forest and out-of-canvas edges
    assert t2i.raster_edges(np.zeros((0, 7)), [32, 48]).sum() == 0
    assert t2i.raster_edges(np.array([[2.0, 2.0, 0, 3.0, 3.0, 0, 0.01]]), [32, 32]).sum() == 0
    with pytest.raises(NotImplementedError):
        t2i.rasterize_forest(rows, [64, 64], colorize="continous")


def test_batch_device_matches_single(t2i):
    import torch
    g0 = rows_to_edges7(load_graph_rows("graph_small_s0.csv"))
    g1 = rows_to_edges7(load_graph_rows("graph_small_s1.csv"))
    graphs = [g0, np.zeros((0, 7)), g1]
    offs = np.cumsum([0] + [len(g) for g in graphs])
    out = t2i.raster_batch_device(torch.from_numpy(np.concatenate(graphs)).cuda(), offs, [304, 304])
    torch.cuda.synchronize()
    host = out.cpu().numpy()
    for i, g in enumerate(graphs):
        assert np.array_equal(host[i], t2i.raster_edges(g, [304, 304])), i


def test_training_transform_vs_oracle_restatement(t2i, tmp_path):
    """LoadGraphAndFilterByRandomRadiusd (data_transforms.py:358-387) against the oracle-side restatement of the
    same class (oracle/transforms_oracle.py, built on the Agg oracle): paired keys with a shared blackdict, the blackdict
    pickle path, allow_missing_keys, RNG consumption."""
    import torch
    from octa_autosegmentation_b200.data_transforms import LoadGraphAndFilterByRandomRadiusd
    from oracle import transforms_oracle
    p = tmp_path / "g.csv"
    shutil.copy(os.path.join(GOLDEN, "graph_small_s0.csv"), p)
    kw = dict(keys=["real_A", "real_B"], image_resolutions=[[304, 304], [1216, 1216]], min_radius=[0, 0.001], max_dropout_prob=0.3)
    ours, theirs = LoadGraphAndFilterByRandomRadiusd(**kw), transforms_oracle.LoadGraphAndFilterByRandomRadiusd(**kw)
    random.seed(7)
    a = ours({"real_A": str(p), "real_B": str(p)})
    na = random.random()
    random.seed(7)
    b = theirs({"real_A": str(p), "real_B": str(p)})
    nb = random.random()
    assert na == nb
    for k, shape in (("real_A", (304, 304)), ("real_B", (1216, 1216))):
        assert isinstance(a[k], torch.Tensor) and a[k].dtype == torch.float32 and tuple(a[k].shape) == shape
        assert torch.equal(a[k], b[k]), k
    # a pickled blackdict: nothing else is dropped, no draw for p (tree2img.py:58-62)
    bd = {(0.5, 0.5, 0.0): True}
    rows = load_graph_rows("graph_small_s0.csv")
    for r in rows[40:60]:
        bd[tuple(float(c) for c in r["node2"][1:-1].split(" ") if c)] = True
    pk = tmp_path / "bd.pkl"
    with open(pk, "wb") as f:
        pickle.dump(bd, f)
    one = dict(keys="real_A", image_resolutions=[[200, 100]], min_radius=[0.0009], max_dropout_prob=0.5, MIP_axis=1, allow_missing_keys=True)
    random.seed(9)
    a = LoadGraphAndFilterByRandomRadiusd(**one)({"real_A": str(p), "blackdict": str(pk)})
    b = transforms_oracle.LoadGraphAndFilterByRandomRadiusd(**one)({"real_A": str(p), "blackdict": str(pk)})
    assert torch.equal(a["real_A"], b["real_A"]) and tuple(a["real_A"].shape) == (100, 200)
    full = transforms_oracle.LoadGraphAndFilterByRandomRadiusd(**dict(one, max_dropout_prob=0))({"real_A": str(p)})
    assert not torch.equal(full["real_A"], b["real_A"])        # the blackdict really removed subtrees
    c = LoadGraphAndFilterByRandomRadiusd(keys=["x", "real_A"], image_resolutions=[[8, 8], [200, 100]], min_radius=[0, 0.0009], MIP_axis=1,
                                          allow_missing_keys=True)({"real_A": str(p), "blackdict": str(pk)})
    assert "x" not in c and torch.equal(c["real_A"], b["real_A"])
