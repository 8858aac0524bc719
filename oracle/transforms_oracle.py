"""TEST INFRASTRUCTURE ONLY: oracle-side restatement of data/data_transforms.py:358-387 `LoadGraphAndFilterByRandomRadiusd`
(the training-time consumer of the graph CSVs) on top of oracle.agg_oracle.rasterize_forest -- independent of
the product's tree2img / data_transforms.  MONAI's MapTransform base only contributes `keys` / `allow_missing_keys`."""
import csv
import pickle

import numpy as np

from . import agg_oracle


class LoadGraphAndFilterByRandomRadiusd:
    def __init__(self, keys, allow_missing_keys=False, image_resolutions=[[304, 304]], min_radius=[0], max_dropout_prob=0, MIP_axis=2):
        self.keys = (keys,) if isinstance(keys, str) else tuple(keys)      # monai.transforms.MapTransform.__init__ (ensure_tuple)
        self.allow_missing_keys = allow_missing_keys
        self.min_radius = min_radius
        self.image_resolutions = image_resolutions
        self.max_dropout_prob = max_dropout_prob
        self.MIP_axis = MIP_axis

    def __call__(self, data):
        import torch
        if "blackdict" in data:                                           # data_transforms.py:370-373
            with open(data["blackdict"], mode="rb") as file:
                blackdict = pickle.load(file)
        else:
            blackdict = None
        for i, key in enumerate(self.keys):                               # :376-386
            if key not in data and self.allow_missing_keys:
                continue
            f = list()
            with open(data[key], newline='') as csvfile:
                for row in csv.DictReader(csvfile):
                    f.append(row)
            img, blackdict = agg_oracle.rasterize_forest(f, self.image_resolutions[i], self.MIP_axis, min_radius=self.min_radius[i],
                                                         max_dropout_prob=self.max_dropout_prob, blackdict=blackdict)
            data[key] = torch.tensor(img.astype(np.float32))
        return data
