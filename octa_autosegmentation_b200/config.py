"""Configuration surface of the generator: the same three-section YAML the reference reads
(`Greenhouse`, `output`, `Forest`; reference: vessel_graph_generation/utilities.py:25-35 read_config)
plus the dotted command-line overrides of utils/config_overrides.py:18-62
(`--A.b.c value`, `--A.b=value`, bare `--A.b` meaning true; values typed through YAML).

`default_config()` returns the parameter values of the reference's shipped 3x3 mm^2 macular setup
(docker/vessel_graph_gen_docker_config.yml and vessel_graph_generation/configs/dataset_18_June_2023.yml
carry identical growth values)."""
from __future__ import annotations

import copy
import os
from typing import Any

import yaml

_MODE_KEYS = ("name", "I", "N", "eps_n", "eps_s", "eps_k", "delta_art", "delta_ven", "gamma_art", "gamma_ven",
              "phi", "omega", "kappa", "delta_sigma")
_SVC = ("SVC", 100, 2000, 0.18, 0.135, 0.135, 0.2925, 0.2925, 50, 50, 15, 0.3, 2.55, 0.02)
_DVC = ("DVC", 150, 2000, 0.09, 0.0675, 0.0675, 0.14625, 0.14625, 90, 90, 15, 0, 2.9, 0.02)

_DEFAULT = {
    "Greenhouse": {
        "SimulationSpace": {"no_voxel_x": 1, "no_voxel_y": 1, "no_voxel_z": 0.0131},
        "d": 0.1, "r": 0.0025, "FAZ_radius_bound": [0.44, 0.04], "rotation_radius": 1.05,
        "FAZ_center": [0.5, 0.5], "nerve_center": [10.56, 5.16], "nerve_radius": 0.3, "param_scale": 3,
        "modes": [dict(zip(_MODE_KEYS, _SVC)), dict(zip(_MODE_KEYS, _DVC))],
    },
    "output": {"directory": "./vessel_graphs", "image_scale_factor": 304, "save_trees": True,
               "save_3D_volumes": None, "save_2D_image": True, "proj_axis": 2, "save_stats": False},
    "Forest": {"type": "stumps", "N_trees": 8,
               "source_walls": {"x0": True, "x1": True, "y0": True, "y1": True, "z0": False, "z1": False}},
}


def default_config() -> dict:
    return copy.deepcopy(_DEFAULT)


def read_config(configpath: str) -> dict:
    path = os.path.abspath(configpath)
    with open(path, "r") as f:
        try:
            return yaml.safe_load(f)
        except Exception:
            print("Your provided config file at %s is not a valid yaml file!" % path)
            raise


def parse_cli_overrides(unknown_args: list) -> list:
    out, i = [], 0
    while i < len(unknown_args):
        tok = unknown_args[i]
        if not isinstance(tok, str) or not tok.startswith("--"):
            i += 1
            continue
        body = tok[2:]
        if "=" in body:
            k, v = body.split("=", 1)
            i += 1
        elif i + 1 < len(unknown_args) and isinstance(unknown_args[i + 1], str) and not unknown_args[i + 1].startswith("--"):
            k, v = body, unknown_args[i + 1]
            i += 2
        else:
            k, v = body, "true"
            i += 1
        out.append((k, v))
    return out


def apply_cli_overrides_from_unknown_args(config: dict, unknown_args: list) -> None:
    """In-place dotted overrides; keys without a dot are ignored (they are ordinary flags)."""
    for key, raw in parse_cli_overrides(unknown_args):
        if "." not in key:
            continue
        parts = key.split(".")
        d: Any = config
        for p in parts[:-1]:
            if p not in d or not isinstance(d[p], dict):
                d[p] = {}
            d = d[p]
        try:
            val = yaml.safe_load(raw)
        except Exception:
            val = raw
        d[parts[-1]] = val
