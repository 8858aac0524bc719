"""Forest of the reference (forest.py:13-207): a handle the Greenhouse fills from the engine's edge table once
`develop_forest()` has run.  The stumps of forest.py:38-181 are placed by the engine (octa_grow_host.cu init_graph) with the
seeded streams, so a Forest holds no trees before the growth."""
from __future__ import annotations

import csv
import os

import numpy as np

from .arterial_tree import ArterialTree


class Forest:
    def __init__(self, config: dict, d_0: float, r_0: float, sim_space, arterial=True, nerve_center=None, nerve_radius=0):
        if config["type"] not in ("nerve", "stumps"):
            raise NotImplementedError(f"The Forest initialization type '{config['type']}' is not implemented. Try 'stump' or 'nerve' instead.")
        self.config = config
        self.d_0, self.r_0 = d_0, r_0
        self.trees = []
        self.sim_space = sim_space
        self.size_x, self.size_y, self.size_z = self.sim_space.shape
        self.arterial = arterial

    def _fill(self, edges7: np.ndarray):
        """Rows (node xyz, proximal node xyz, radius) in the reference's export order -- per tree, level order without the root
        (generate_vessel_graph.py:45-56) -- back into trees: a row whose proximal position is not a node of the tree being built
        opens the next tree at that position (a root).  A root carries no row; its radius is that of its first segment (what
        the Murray update of arterial_tree.py:174-184 gives a node with one child)."""
        self.trees = []
        nodes = {}
        tree = None
        for row in np.asarray(edges7, dtype=np.float64).reshape(-1, 7):
            key = row[3:6].tobytes()
            parent = nodes.get(key)
            if parent is None:
                # tree names count from 1 for 'stumps' (forest.py:87-88) and from 0 for 'nerve' (:49-50)
                name = f'{"Arterial" if self.arterial else "Venous"}Tree{len(self.trees) + (0 if self.config["type"] == "nerve" else 1)}'
                tree = ArterialTree(name, row[3:6].copy(), float(row[6]), self.size_x, self.size_y, self.size_z, self)
                self.trees.append(tree)
                nodes = {key: tree.root}
                parent = tree.root
            nodes[row[0:3].tobytes()] = tree._attach(row[0:3].copy(), float(row[6]), parent)

    def get_trees(self):
        return self.trees

    def get_nodes(self):
        for tree in self.trees:
            for node in tree.get_tree_iterator(exclude_root=False, only_active=False):
                yield node

    def get_node_coords(self):
        for node in self.get_nodes():
            yield node.position

    def save(self, save_directory="."):
        name = f'{"Arterial" if self.arterial else "Venous"}Forest'
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, name + ".csv"), "w+") as file:
            writer = csv.writer(file)
            writer.writerow(["node1", "node2", "radius"])
            for tree in self.get_trees():
                for current_node in tree.get_tree_iterator(exclude_root=True, only_active=False):
                    writer.writerow([current_node.position, current_node.get_proximal_node().position, current_node.radius])
