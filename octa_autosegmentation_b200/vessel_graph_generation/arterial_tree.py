"""Node / ArterialTree of the reference (arterial_tree.py:7-228) as read-only views of a grown forest: the attributes and the
accessors callers use after `develop_forest()` (positions, radii, parents / children, level-order iteration).  The growth
operations themselves (`add_node`, `optimize_edge_radius_to_root`) run on the device (csrc/octa_grow_kernels.cu k_commit)."""
from __future__ import annotations

from collections import deque

import numpy as np


class Node:
    def __init__(self, tree, name, position, radius, parent=None, kappa=4):
        self.tree = tree
        self.name = name
        self.position = np.array(position)
        self.radius = radius
        self.kappa = kappa
        self.parent = parent
        self.children = []
        self.active = tree.forest.sim_space.is_valid_position(self.position)        # arterial_tree.py:33,70-71
        self.proximal_num_segments = 0 if parent is None else parent.proximal_num_segments + 1
        if parent is not None:
            parent.children.append(self)

    @property
    def is_root(self):
        return self.parent is None

    @property
    def is_leaf(self):
        return len(self.children) == 0

    @property
    def is_inter_node(self):
        return self.parent is not None and len(self.children) == 1

    @property
    def is_bifurcation_node(self):
        return len(self.children) == 2

    def __repr__(self):
        return "{} (position: {}, radius: {}, active: {})".format(self.name, self.position, self.radius, self.active)

    def _distal(self, child_index):
        if self.is_leaf:
            raise RuntimeError("Unable to analyze distal part. This node does not have any children.")
        if self.is_bifurcation_node:
            if child_index is None:
                raise RuntimeError("Unable to analyze distal part. Unclear which branch to return.")
            return self.children[child_index]
        return self.children[0]

    def get_distal_node(self, child_index=None):
        return self._distal(child_index)

    def get_distal_position(self, child_index=None):
        return self._distal(child_index).position

    def get_distal_radius(self, child_index=None):
        return self._distal(child_index).radius

    def get_distal_segment(self, child_index=None):
        return self._distal(child_index).position - self.position

    def get_proximal_node(self):
        if self.is_root:
            raise RuntimeError("Unable to analyze proximal part. This node is the root.")
        return self.parent

    def get_proximal_position(self):
        return self.get_proximal_node().position

    def get_proximal_radius(self):
        self.get_proximal_node()
        return self.radius

    def get_proximal_segment(self):
        return self.get_proximal_node().position - self.position


class ArterialTree:
    def __init__(self, name, root_position, r_0, size_x, size_y, size_z, forest):
        self.name = name
        self.init_size_x = self.size_x = size_x
        self.init_size_y = self.size_y = size_y
        self.init_size_z = self.size_z = size_z
        self.r_0 = r_0
        self.scaling_factor = 1.0
        self.forest = forest
        self.root = Node(self, "Root", position=root_position, radius=r_0)
        self.name_counter = 1

    def _attach(self, position, radius, parent) -> Node:
        name = "Node" + str(self.name_counter)
        self.name_counter += 1
        return Node(self, name, position=position, radius=radius, parent=parent)

    def add_node(self, position, radius, parent, kappa=4):
        raise NotImplementedError("trees are grown on the device (Greenhouse.develop_forest); this view is read-only")

    def get_tree_iterator(self, exclude_root=False, only_active=False):
        """anytree.LevelOrderIter(self.root, filter_) of arterial_tree.py:226-229: breadth first, children in creation order."""
        queue = deque([self.root])
        while queue:
            n = queue.popleft()
            queue.extend(n.children)
            if (n.parent is not None or not exclude_root) and (n.active or not only_active):
                yield n
