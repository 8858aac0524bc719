"""Training-time consumer of the graph CSVs: drop-in for the reference's
data/data_transforms.py:358-387 `LoadGraphAndFilterByRandomRadiusd`, which re-rasterizes every CSV per sample per
epoch through `rasterize_forest`.  Same constructor arguments and dictionary semantics (the MONAI `MapTransform` base is
not needed: only `keys` / `allow_missing_keys` are used), same blackdict pickle handling, GPU rasterization."""
from __future__ import annotations

import csv
import pickle
from typing import Sequence

import numpy as np

from .tree2img import rasterize_forest


class LoadGraphAndFilterByRandomRadiusd:
    """Given a graph csv file, only load edges with radius larger than the given threshold, then turn the graph into a
    grayscale image of the given shape (data_transforms.py:362-387).  `data[key]` (a csv path) is replaced by a
    float32 torch tensor [H, W]; a shared `data["blackdict"]` pickle path keeps paired renderings consistent."""

    def __init__(self, keys: Sequence[str], allow_missing_keys: bool = False, image_resolutions=[[304, 304]],
                 min_radius=[0], max_dropout_prob=0, MIP_axis=2) -> None:
        self.keys = [keys] if isinstance(keys, str) else list(keys)
        self.allow_missing_keys = allow_missing_keys
        self.min_radius = min_radius
        self.image_resolutions = image_resolutions
        self.max_dropout_prob = max_dropout_prob
        self.MIP_axis = MIP_axis

    def __call__(self, data):
        import torch

        if "blackdict" in data:
            with open(data["blackdict"], mode="rb") as file:
                blackdict = pickle.load(file)
        else:
            blackdict = None
        for i, key in enumerate(self.keys):
            if key not in data and self.allow_missing_keys:
                continue
            with open(data[key], newline="") as csvfile:
                f = list(csv.DictReader(csvfile))
            img, blackdict = rasterize_forest(f, self.image_resolutions[i], self.MIP_axis, min_radius=self.min_radius[i],
                                              max_dropout_prob=self.max_dropout_prob, blackdict=blackdict)
            data[key] = torch.tensor(img.astype(np.float32))
        return data
