"""Minimal stand-in for the `anytree` package (absent from this image; the reference pins
anytree>=2.13.0, pyproject.toml:8).  TEST INFRASTRUCTURE ONLY: it exists so that the unmodified
reference modules under /root/reference can be imported by oracle/ref_harness.py.

Semantics honoured (only what vessel_graph_generation/arterial_tree.py uses):
  * NodeMixin.parent setter appends to the parent's children in attach order and fires
    _pre_attach/_post_attach (and detach hooks on re-parenting);
  * children is a tuple in attach order; is_leaf = no children; is_root = parent is None;
  * LevelOrderIter(root, filter_) = breadth-first by level, children in attach order, with
    filter_ applied to *yielding* only (children of filtered nodes are still traversed).
"""


class NodeMixin:
    @property
    def parent(self):
        return self.__dict__.get("_NodeMixin__parent")

    @parent.setter
    def parent(self, value):
        old = self.__dict__.get("_NodeMixin__parent")
        if old is value:
            if "_NodeMixin__parent" not in self.__dict__:
                self.__dict__["_NodeMixin__parent"] = None
            return
        if old is not None:
            self._pre_detach(old)
            old._NodeMixin__children_list().remove(self)
            self.__dict__["_NodeMixin__parent"] = None
            self._post_detach(old)
        if value is not None:
            self._pre_attach(value)
            value._NodeMixin__children_list().append(self)
            self.__dict__["_NodeMixin__parent"] = value
            self._post_attach(value)
        else:
            self.__dict__["_NodeMixin__parent"] = None

    def _NodeMixin__children_list(self):
        lst = self.__dict__.get("_NodeMixin__children")
        if lst is None:
            lst = []
            self.__dict__["_NodeMixin__children"] = lst
        return lst

    @property
    def children(self):
        return tuple(self._NodeMixin__children_list())

    @property
    def is_leaf(self):
        return len(self._NodeMixin__children_list()) == 0

    @property
    def is_root(self):
        return self.parent is None

    def _pre_attach(self, parent):
        pass

    def _post_attach(self, parent):
        pass

    def _pre_detach(self, parent):
        pass

    def _post_detach(self, parent):
        pass


class LevelOrderIter:
    def __init__(self, node, filter_=None, stop=None, maxlevel=None):
        self.node = node
        self.filter_ = filter_ or (lambda n: True)

    def __iter__(self):
        level = [self.node]
        while level:
            nxt = []
            for n in level:
                if self.filter_(n):
                    yield n
                nxt.extend(n.children)
            level = nxt


class RenderTree:
    def __init__(self, node, *a, **k):
        self.node = node

    def __str__(self):
        return "RenderTree(%r)" % (self.node,)
