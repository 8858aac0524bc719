"""SimulationSpace of the reference (simulation_space.py:10-54,89-112) as far as callers outside the growth loop use it: the
shape of the space and `is_valid_position`.  Candidate sampling (`get_candidate_sinks`, `get_random_valid_position`) happens on
the device (csrc/octa_grow_kernels.cu k_prepare, octa_grow_host.cu init_graph)."""
from __future__ import annotations

import warnings
from math import ceil, sqrt

import numpy as np

GEOMETRY_SIZE = 76


class SimulationSpace:
    def __init__(self, config: dict, FAZ_center=None, FAZ_radius=None, nerve_center=None, nerve_radius=None):
        self.fixed_geometry = config.get("oxygen_sample_geometry_path") is not None
        if self.fixed_geometry:                                                  # :29-34
            self.geometry = np.load(config["oxygen_sample_geometry_path"])
            self.geometry_size = max(self.geometry.shape)
            self.shape = np.array(self.geometry.shape) / self.geometry_size
            self.size_x, self.size_y, self.size_z = self.shape
            self.valid_voxels = np.argwhere(self.geometry)
        else:                                                                    # :36-54
            self.size_x, self.size_y, self.size_z = config["no_voxel_x"], config["no_voxel_y"], config["no_voxel_z"]
            self.shape = np.array([self.size_x, self.size_y, self.size_z])
            assert all(self.shape > 0), "The simulation space dimensions must be postive!"
            if any(self.shape > 1) or all(self.shape != 1):
                warnings.warn("Warning: The largest dimension of the simulation space should be exactly one.")
            self.geometry_size = GEOMETRY_SIZE
            self.FAZ_center = np.array(FAZ_center) * self.geometry_size
            self.FAZ_radius = np.array(FAZ_radius) * self.geometry_size * 0.5
            y_coords, x_coords = np.ogrid[:ceil(self.size_x * self.geometry_size), :ceil(self.size_y * self.geometry_size)]
            self.geometry = (x_coords - self.FAZ_center[0]) ** 2 + (y_coords - self.FAZ_center[1]) ** 2 > self.FAZ_radius ** 2
            if all(np.asarray(nerve_center) - nerve_radius <= 1):
                self.nerve_center = np.array(nerve_center) * self.geometry_size
                self.nerve_radius = np.array(nerve_radius) * self.geometry_size
                self.geometry &= (x_coords - self.nerve_center[0]) ** 2 + (y_coords - self.nerve_center[1]) ** 2 > self.nerve_radius ** 2
            else:
                self.nerve_radius = None
                self.nerve_center = None
            self.geometry = np.expand_dims(self.geometry, -1)
            self.valid_voxels = np.argwhere(self.geometry)

    def is_valid_position(self, pos) -> bool:
        """:89-98, including the reference's mixed units (a unit-cube position against the FAZ centre in voxel units, the
        distance taken over the first two coordinates only: zip() stops at the shorter sequence)."""
        pos = np.asarray(pos, dtype=np.float64)
        if any(pos >= self.shape) or any(pos < 0):
            return False
        if self.fixed_geometry:
            return bool(self.geometry[tuple((pos * self.geometry_size).astype(np.uint16))] > 0)
        return sqrt(sum((a - b) ** 2 for a, b in zip(pos, self.FAZ_center))) > self.FAZ_radius
