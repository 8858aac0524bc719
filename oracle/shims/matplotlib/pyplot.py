"""pyplot stand-in (see matplotlib/__init__.py)."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
_current = [None]


class _Patch:
    def __init__(self):
        self.facecolor = "white"

    def set_facecolor(self, c):
        self.facecolor = c


class _Axes:
    def __init__(self, rect, frameon=True, xticks=None, yticks=None):
        if list(rect) != [0.0, 0.0, 1.0, 1.0] or frameon or xticks or yticks:
            raise NotImplementedError("the matplotlib stand-in only provides the full-bleed, frameless axes of tree2img.py:55")
        self.inverted = False
        self.collections = []

    def invert_yaxis(self):
        self.inverted = not self.inverted

    def add_collection(self, c):
        self.collections.append(c)


class _Canvas:
    def __init__(self, fig):
        self.fig = fig
        self._rgba = None

    def get_width_height(self):
        return int(self.fig.figsize[0] * self.fig.dpi), int(self.fig.figsize[1] * self.fig.dpi)

    def draw(self):
        if _ROOT not in sys.path:
            sys.path.insert(0, _ROOT)
        from oracle import agg_oracle
        fig = self.fig
        if fig.patch.facecolor != "black" or fig.axes is None or not fig.axes.inverted:
            raise NotImplementedError("the matplotlib stand-in renders the black, y-inverted figure of tree2img.py:53-56 only")
        W, H = self.get_width_height()
        segs, lws = [], []
        for c in fig.axes.collections:
            for (a, b), lw in zip(c.segments, c.linewidths):
                segs.append((a[0], a[1], b[0], b[1]))
                lws.append(lw)
        gray = agg_oracle.raster_segments(np.array(segs, dtype=np.float64).reshape(-1, 4), np.array(lws, dtype=np.float64), H, W)
        self._rgba = np.dstack([gray, gray, gray, np.full_like(gray, 255)])

    def buffer_rgba(self):
        return memoryview(np.ascontiguousarray(self._rgba))


class _Figure:
    def __init__(self, figsize, dpi=100):
        self.figsize = figsize
        self.dpi = dpi
        self.patch = _Patch()
        self.axes = None
        self.canvas = _Canvas(self)


def figure(figsize=(6.4, 4.8), dpi=100, **kw):
    _current[0] = _Figure(figsize, dpi)
    return _current[0]


def axes(rect, **kw):
    ax = _Axes(rect, **kw)
    _current[0].axes = ax
    return ax


def close(fig=None):
    _current[0] = None
