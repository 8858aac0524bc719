"""The statistic plots behind `output.save_stats` (Greenhouse.save_stats, greenhouse.py:401-441, and
tree2img.plot_vessel_radii, tree2img.py:294-314): same file names, same data, same axes -- drawn with PIL, because the
plots are matplotlib figures in the reference and matplotlib is not part of this stack.  The DATA behind them is what is
checked against the reference (final sink lists and per-iteration counts: tests/test_growth_gpu.py, test_oracle_growth.py);
the pixels are not meant to match matplotlib's.

    oxy_distribution.png   final oxygen sinks, x = pos[1], y = 1 - pos[0], red dots, both axes [0, 1]
    co2_distribution.png   final CO2 sources, blue dots
    time_per_step.png      seconds per iteration (here: the device time of the batch's growth loop spread evenly over
                           iterations and samples -- the loop runs all samples of a batch together)
    growth_over_time.png   arterial nodes / oxygen sinks / venous nodes / CO2 sources after every iteration
    hist.png               histogram of the drawn radii (39 bins between min and max, log counts)
"""
from __future__ import annotations

import math
import os
import time
from typing import Sequence

import numpy as np

C = {"r": (255, 0, 0), "b": (0, 0, 255), "C0": (31, 119, 180), "C1": (255, 127, 14), "C2": (44, 160, 44), "C3": (214, 39, 40)}


def _nice_ticks(lo: float, hi: float, n: int = 6):
    if not (hi > lo):
        hi = lo + 1.0
    raw = (hi - lo) / max(n, 1)
    mag = 10.0 ** math.floor(math.log10(raw))
    step = min((m for m in (1, 2, 2.5, 5, 10) if m * mag >= raw), default=10) * mag
    t0 = math.ceil(lo / step - 1e-9) * step
    ticks = []
    while t0 <= hi + 1e-9 * step:
        ticks.append(0.0 if abs(t0) < 1e-12 * step else t0)
        t0 += step
    return ticks


def _fmt(v: float) -> str:
    return ("%.6g" % v)


class Axes:
    """A single set of axes on a white canvas: data box, ticks with labels, title, axis labels, optional legend."""

    def __init__(self, size=(600, 600), xlim=(0.0, 1.0), ylim=(0.0, 1.0), title="", xlabel="", ylabel="", logy=False):
        from PIL import Image, ImageDraw, ImageFont
        self.Image = Image
        self.img = Image.new("RGB", size, (255, 255, 255))
        self.d = ImageDraw.Draw(self.img)
        self.font = ImageFont.load_default()
        self.W, self.H = size
        self.l, self.r, self.t, self.b = 70, self.W - 20, 36, self.H - 52
        self.logy = logy
        self.xlim = (float(xlim[0]), float(xlim[1]) if xlim[1] > xlim[0] else float(xlim[0]) + 1.0)
        y0, y1 = float(ylim[0]), float(ylim[1])
        if logy:
            y0, y1 = math.log10(max(y0, 1e-300)), math.log10(max(y1, 1e-300))
        self.ylim = (y0, y1 if y1 > y0 else y0 + 1.0)
        self._frame(title, xlabel, ylabel)

    def px(self, x):
        return self.l + (np.asarray(x, dtype=np.float64) - self.xlim[0]) / (self.xlim[1] - self.xlim[0]) * (self.r - self.l)

    def py(self, y):
        y = np.asarray(y, dtype=np.float64)
        if self.logy:
            y = np.log10(np.maximum(y, 1e-300))
        return self.b - (y - self.ylim[0]) / (self.ylim[1] - self.ylim[0]) * (self.b - self.t)

    def _text(self, xy, s, anchor="la", fill=(0, 0, 0)):
        try:
            self.d.text(xy, s, fill=fill, font=self.font, anchor=anchor)
        except (ValueError, TypeError):      # bitmap fonts of older PIL builds know no anchors
            w = self.d.textlength(s, font=self.font)
            x, y = xy
            x -= w / 2 if anchor[0] == "m" else (w if anchor[0] == "r" else 0)
            y -= 5 if anchor[1] == "m" else (10 if anchor[1] in "sb" else 0)
            self.d.text((x, y), s, fill=fill, font=self.font)

    def _frame(self, title, xlabel, ylabel):
        d = self.d
        d.rectangle([self.l, self.t, self.r, self.b], outline=(0, 0, 0))
        for v in _nice_ticks(*self.xlim):
            x = float(self.px(v))
            d.line([x, self.b, x, self.b + 4], fill=(0, 0, 0))
            self._text((x, self.b + 7), _fmt(v), "ma")
        if self.logy:
            for e in range(int(math.ceil(self.ylim[0] - 1e-9)), int(math.floor(self.ylim[1] + 1e-9)) + 1):
                y = float(self.py(10.0 ** e))
                d.line([self.l - 4, y, self.l, y], fill=(0, 0, 0))
                self._text((self.l - 7, y), "1e%d" % e, "rm")
        else:
            for v in _nice_ticks(*self.ylim):
                y = float(self.py(v))
                d.line([self.l - 4, y, self.l, y], fill=(0, 0, 0))
                self._text((self.l - 7, y), _fmt(v), "rm")
        if title:
            self._text(((self.l + self.r) / 2, self.t - 20), title.replace("₂", "2"), "ma")
        if xlabel:
            self._text(((self.l + self.r) / 2, self.b + 26), xlabel, "ma")
        if ylabel:
            w = int(self.d.textlength(ylabel, font=self.font)) + 4
            lab = self.Image.new("RGB", (w, 14), (255, 255, 255))
            from PIL import ImageDraw
            ImageDraw.Draw(lab).text((2, 1), ylabel, fill=(0, 0, 0), font=self.font)
            lab = lab.rotate(90, expand=True)
            self.img.paste(lab, (6, int((self.t + self.b) / 2 - w / 2)))

    def _inside(self, x, y):
        return (x >= self.l) & (x <= self.r) & (y >= self.t) & (y <= self.b)

    def dots(self, x, y, color):
        """matplotlib's '.' marker at default size: a filled disc of about 5 px."""
        X, Y = self.px(x), self.py(y)
        keep = self._inside(X, Y)
        for a, b in zip(X[keep], Y[keep]):
            self.d.ellipse([a - 2.5, b - 2.5, a + 2.5, b + 2.5], fill=color)

    def line(self, y, color, x=None):
        y = np.asarray(y, dtype=np.float64)
        if len(y) == 0:
            return
        x = np.arange(len(y)) if x is None else np.asarray(x, dtype=np.float64)
        pts = list(zip(self.px(x).tolist(), self.py(y).tolist()))
        if len(pts) == 1:
            pts = pts * 2
        self.d.line(pts, fill=color, width=2)

    def bars(self, edges, counts, color, alpha=0.5):
        fill = tuple(int(255 - alpha * (255 - c)) for c in color)      # alpha over white
        for k, n in enumerate(counts):
            if n <= 0:
                continue
            x0, x1 = float(self.px(edges[k])), float(self.px(edges[k + 1]))
            y = max(float(self.py(n)), self.t)
            self.d.rectangle([x0, y, max(x1, x0 + 1), self.b], fill=fill)
        self.d.rectangle([self.l, self.t, self.r, self.b], outline=(0, 0, 0))

    def legend(self, names: Sequence[str], colors):
        w = max(int(self.d.textlength(n.replace("₂", "2"), font=self.font)) for n in names) + 44
        x0, y0 = self.l + 8, self.t + 8
        self.d.rectangle([x0, y0, x0 + w, y0 + 16 * len(names) + 6], fill=(255, 255, 255), outline=(200, 200, 200))
        for i, (n, c) in enumerate(zip(names, colors)):
            y = y0 + 11 + 16 * i
            self.d.line([x0 + 6, y, x0 + 30, y], fill=c, width=2)
            self._text((x0 + 36, y), n.replace("₂", "2"), "lm")

    def save(self, path):
        self.img.save(path)


def _autoscale(*series):
    vals = [np.asarray(s, dtype=np.float64) for s in series if len(s)]
    if not vals:
        return 0.0, 1.0
    lo, hi = min(float(v.min()) for v in vals), max(float(v.max()) for v in vals)
    pad = 0.05 * (hi - lo) if hi > lo else 0.5
    return lo - pad, hi + pad


def save_stats(out_dir: str, oxys: np.ndarray, co2s, per_step: np.ndarray, time_per_step: Sequence[float]):
    """Greenhouse.save_stats (greenhouse.py:401-441).  oxys / co2s: final sink positions [n, 3] (co2s None: no venous forest);
    per_step: int [iterations, 4] = arterial nodes, oxygen sinks, venous nodes, CO2 sources after every iteration (the
    reference's lists start with a 0 entry, which is added here); time_per_step: seconds per iteration."""
    for name, pts, col, title in (("oxy_distribution", oxys, C["r"], "Final Oxygen Sink Distribution"),
                                  ("co2_distribution", co2s, C["b"], "Final CO₂ Sink Distribution")):
        ax = Axes((600, 600), (0, 1), (0, 1), title)
        pts = np.zeros((0, 3)) if pts is None else np.asarray(pts, dtype=np.float64).reshape(-1, 3)
        if len(pts) > 0:
            ax.dots(pts[:, 1], 1 - pts[:, 0], col)
        ax.save(os.path.join(out_dir, name + ".png"))
    tps = np.asarray(time_per_step, dtype=np.float64)
    total = time.strftime("%H:%M:%S", time.gmtime(float(tps.sum())))
    ax = Axes((600, 600), _autoscale(np.arange(len(tps))), _autoscale(tps), "Runtime Per Iteration (Total=%s)" % total,
              "Iterations", "Seconds")
    ax.line(tps, C["C0"])
    ax.save(os.path.join(out_dir, "time_per_step.png"))
    ps = np.asarray(per_step, dtype=np.int64).reshape(-1, 4)
    ps = np.concatenate([np.zeros((1, 4), dtype=np.int64), ps])
    venous = co2s is not None
    cols = [0, 1, 2, 3] if venous else [0, 1]
    ax = Axes((600, 600), _autoscale(np.arange(len(ps))), _autoscale(*[ps[:, c] for c in cols]), "Growth Over Time",
              "Iterations", "Amount")
    for c in cols:
        ax.line(ps[:, c], C["C%d" % c])
    ax.legend(["Arterial Nodes", "Oxygen Sinks", "Venous Nodes", "CO₂ Sources"] if venous else ["Nodes", "Oxygen Sinks"],
              [C["C%d" % c] for c in cols])
    ax.save(os.path.join(out_dir, "growth_over_time.png"))


def plot_vessel_radii(out_dir: str, radius_list: Sequence[float] = ()):
    """tree2img.plot_vessel_radii (tree2img.py:294-314): 39 equal bins between the smallest and the largest radius, log counts."""
    r = np.asarray(radius_list, dtype=np.float64)
    if len(r) == 0:
        raise ValueError("min() arg is an empty sequence")      # what the reference raises for an empty list
    bins = np.linspace(r.min(), r.max(), 40)
    counts, edges = np.histogram(r, bins=bins)
    top = max(int(counts.max()), 1)
    ax = Axes((640, 480), (r.min(), r.max()), (0.7, top * 1.5), "Vessel Radii Distribution", "Radius", "Count", logy=True)
    ax.bars(edges, counts, C["C0"], 0.5)
    ax.save(os.path.join(out_dir, "hist.png"))
