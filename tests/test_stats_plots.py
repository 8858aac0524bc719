"""CPU: the PIL statistic plots behind output.save_stats (octa_autosegmentation_b200/stats_plots.py; reference:
greenhouse.py:401-441, tree2img.py:294-314) -- files, sizes, and that the data lands where the axes say."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from octa_autosegmentation_b200 import stats_plots as sp


def test_save_stats_files_and_marker_positions(tmp_path):
    from PIL import Image
    gold = np.load(os.path.join(GOLDEN, "stats_small_s0.npz"))
    sp.save_stats(str(tmp_path), gold["oxys"], gold["co2s"], gold["per_step"][1:], np.full(24, 0.002))
    for name in ("oxy_distribution", "co2_distribution", "time_per_step", "growth_over_time"):
        assert Image.open(tmp_path / (name + ".png")).size == (600, 600)
    # one sink at a known place: x = pos[1], y = 1 - pos[0] (greenhouse.py:405), axes [0, 1] x [0, 1]
    d = tmp_path / "one"
    d.mkdir()
    sp.save_stats(str(d), np.array([[0.25, 0.75, 0.0]]), None, np.zeros((3, 4), dtype=int), [0.1, 0.2, 0.3])
    a = np.asarray(Image.open(d / "oxy_distribution.png").convert("RGB")).astype(int)
    red = (a[:, :, 0] > 200) & (a[:, :, 1] < 60) & (a[:, :, 2] < 60)
    ys, xs = np.nonzero(red)
    ax = sp.Axes((600, 600))
    assert len(xs) > 5 and abs(xs.mean() - float(ax.px(0.75))) < 1.5 and abs(ys.mean() - float(ax.py(0.75))) < 1.5
    b = np.asarray(Image.open(d / "co2_distribution.png").convert("RGB")).astype(int)
    assert not ((b[:, :, 2] > 200) & (b[:, :, 0] < 60) & (b[:, :, 1] < 60)).any()      # no venous forest: an empty plot


def test_plot_vessel_radii(tmp_path):
    from PIL import Image
    r = np.concatenate([np.full(1000, 0.001), np.full(10, 0.004), np.linspace(0.001, 0.004, 50)])
    sp.plot_vessel_radii(str(tmp_path), r)
    assert Image.open(tmp_path / "hist.png").size == (640, 480)
    with pytest.raises(ValueError):            # min() of an empty list in the reference (tree2img.py:305)
        sp.plot_vessel_radii(str(tmp_path), [])
