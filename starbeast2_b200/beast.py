"""A minimal stand-in for the slice of BEAST 2's object model the reference's hot-path classes live in
(beast.base.core.BEASTObject / Input, beast.base.inference.{StateNode, CalculationNode, Distribution,
CompoundDistribution, State}, parameter.{RealParameter, IntegerParameter}, evolution.tree.{Tree, TreeParser}).

BEAST.base is not vendored under /root/reference and there is no JVM in this image, so the host side of the
drop-in boundary is mirrored in Python (where the reference is Java): same class names, same Input names
(initByName), same protocol (store / restore / requiresRecalculation / calculateLogP, -inf short-circuit).
This is host plumbing only -- no likelihood arithmetic happens here; that is libsb2gpu's job.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .newick import leaf_sets, parse_newick
from .problem import TreeArrays

IS_CLEAN, IS_DIRTY, IS_FILTHY = 0, 1, 2   # beast.base.evolution.tree.Tree dirtiness levels


class BEASTObject:
    """initByName("name", value, ...) then initAndValidate(), like beast.base.core.BEASTObject."""
    INPUTS: Dict[str, Any] = {}     # name -> default (REQUIRED marks Validate.REQUIRED)
    REQUIRED = object()

    def __init__(self, **kw):
        self._inputs: Dict[str, Any] = {}
        for name, default in self._all_inputs().items():
            self._inputs[name] = [] if isinstance(default, list) else default
        self.id: Optional[str] = None
        if kw:
            self.initByName(*[x for kv in kw.items() for x in kv])

    @classmethod
    def _all_inputs(cls) -> Dict[str, Any]:
        out: Dict[str, Any] = {}
        for k in reversed(cls.__mro__):
            out.update(getattr(k, "INPUTS", {}))
        return out

    def initByName(self, *args):
        if len(args) % 2:
            raise ValueError("initByName expects name/value pairs")
        for name, value in zip(args[::2], args[1::2]):
            key = name.replace(".", "_")
            if key not in self._inputs:
                raise ValueError(f"{type(self).__name__} has no input '{name}'")
            if isinstance(self._inputs[key], list) and not isinstance(value, (list, tuple)):
                self._inputs[key].append(value)
            elif isinstance(self._inputs[key], list):
                self._inputs[key] = list(value)
            else:
                self._inputs[key] = value
        for name, v in self._inputs.items():
            if v is BEASTObject.REQUIRED:
                raise ValueError(f"Input '{name}' of {type(self).__name__} must be specified")   # Validate.REQUIRED
        self.initAndValidate()
        return self

    def get(self, name: str):
        return self._inputs[name.replace(".", "_")]

    def initAndValidate(self):
        pass


class StateNode(BEASTObject):
    def __init__(self, **kw):
        self._dirty = True
        self._version = 0   # bumped on every edit / restore: lets the GPU session see what changed since its last push
        super().__init__(**kw)

    def _touch(self):
        self._version += 1
        self._dirty = True

    def somethingIsDirty(self) -> bool:
        return self._dirty

    def setEverythingDirty(self, d: bool):
        self._dirty = d

    def store(self):
        raise NotImplementedError

    def restore(self):
        raise NotImplementedError


class RealParameter(StateNode):
    INPUTS = {"value": BEASTObject.REQUIRED, "dimension": 1, "lower": -math.inf, "upper": math.inf, "estimate": True}
    DTYPE = np.float64

    def initAndValidate(self):
        v = self.get("value")
        if isinstance(v, str):
            v = [float(x) for x in v.split()]
        vals = np.atleast_1d(np.asarray(v, dtype=self.DTYPE))
        dim = int(self.get("dimension"))
        if dim > len(vals):
            vals = np.resize(vals, dim)
        self.values = vals.copy()
        self._stored = vals.copy()
        self._dirty_idx = np.ones(len(vals), bool)

    def getDimension(self):
        return len(self.values)

    def setDimension(self, dim: int):
        if dim != len(self.values):
            self._version += 1
            self.values = np.resize(self.values, dim)
            self._stored = np.resize(self._stored, dim)
            self._dirty_idx = np.ones(dim, bool)

    def getValue(self, i: int = 0):
        return self.values[i]

    def getValues(self):
        return self.values

    def setValue(self, i, v=None):
        if v is None:
            i, v = 0, i
        self.values[i] = v
        self._dirty_idx[i] = True
        self._touch()

    def isDirty(self, i: int) -> bool:
        return bool(self._dirty_idx[i])

    def getLower(self):
        return self.get("lower")

    def getUpper(self):
        return self.get("upper")

    def setLower(self, v):
        self._inputs["lower"] = v

    def setUpper(self, v):
        self._inputs["upper"] = v

    def store(self):
        self._stored = self.values.copy()

    def restore(self):
        self.values, self._stored = self._stored, self.values
        self._dirty_idx[:] = False
        self._version += 1
        self._dirty = False

    def accept(self):
        self._dirty_idx[:] = False
        self._dirty = False


class IntegerParameter(RealParameter):
    DTYPE = np.int64


class BooleanParameter(RealParameter):
    DTYPE = np.bool_

    def initAndValidate(self):
        v = self.get("value")
        if isinstance(v, str):
            self._inputs["value"] = [x.lower() in ("true", "1") for x in v.split()]
        super().initAndValidate()


class Tree(StateNode):
    """beast.base.evolution.tree.Tree as flat arrays by node number (leaves first, root last)."""
    INPUTS = {"newick": None, "IsLabelledNewick": True, "taxonset": None, "adjustTipHeights": True, "arrays": None, "labels": None}

    def initAndValidate(self):
        if self.get("arrays") is not None:
            self.arrays: TreeArrays = self.get("arrays").copy()
            self.labels: List[str] = list(self.get("labels") or [str(i) for i in range(self.arrays.n_leaves)])
        else:
            taxa = self.get("taxonset")
            names = taxa.asStringList() if taxa is not None else None
            self.arrays, self.labels = parse_newick(self.get("newick"), taxa=names, adjust_tip_heights=bool(self.get("adjustTipHeights")))
        self._stored = self.arrays.copy()
        self.node_dirty = np.full(self.arrays.n_nodes, IS_FILTHY, np.int8)

    # -- BEAST Tree API (the part the hot path reads) --
    def getNodeCount(self):
        return self.arrays.n_nodes

    def getLeafNodeCount(self):
        return self.arrays.n_leaves

    def getRootNr(self):
        return self.arrays.root

    def getTaxaNames(self):
        return self.labels

    def findNode(self, leaf_labels: Sequence[str]) -> int:
        target = frozenset(leaf_labels)
        for i, s in enumerate(leaf_sets(self.arrays, self.labels)):
            if s == target:
                return i
        raise KeyError(leaf_labels)

    # -- edits (operators) --
    def setHeight(self, node: int, h: float):
        """Node.setHeight: marks the node and its children IS_DIRTY (their branch lengths changed)."""
        self.arrays.height[node] = h
        self.node_dirty[node] |= IS_DIRTY
        for c in (self.arrays.left[node], self.arrays.right[node]):
            if c >= 0:
                self.node_dirty[c] |= IS_DIRTY
        self._touch()

    def assignFrom(self, arrays: TreeArrays):
        self.arrays = arrays.copy()
        self.node_dirty[:] = IS_FILTHY
        self._touch()

    def scale(self, f: float):
        self.arrays.height *= f
        self.node_dirty[:] |= IS_DIRTY
        self._touch()

    def store(self):
        self._stored = self.arrays.copy()

    def restore(self):
        self.arrays, self._stored = self._stored, self.arrays
        self.node_dirty[:] = IS_CLEAN
        self._version += 1
        self._dirty = False

    def accept(self):
        self.node_dirty[:] = IS_CLEAN
        self._dirty = False


class TreeParser(Tree):
    pass


class Taxon(BEASTObject):
    def __init__(self, id: str):
        super().__init__()
        self.id = id


class TaxonSet(BEASTObject):
    """Either a list of Taxon (a species) or a list of TaxonSet (the species superset)."""

    def __init__(self, id_or_list, taxa: Optional[list] = None):
        super().__init__()
        if taxa is None:
            self.id, self.taxa = None, list(id_or_list)
        else:
            self.id, self.taxa = id_or_list, list(taxa)

    def asStringList(self) -> List[str]:
        return [t.id for t in self.taxa]


class CalculationNode(BEASTObject):
    def store(self):
        pass

    def restore(self):
        pass

    def accept(self):
        pass

    def requiresRecalculation(self) -> bool:
        return True


class Distribution(CalculationNode):
    def __init__(self, **kw):
        self.logP = 0.0
        self.storedLogP = 0.0
        super().__init__(**kw)

    def calculateLogP(self) -> float:
        raise NotImplementedError

    def getCurrentLogP(self) -> float:
        return self.logP

    def isDirtyCalculation(self) -> bool:
        return True

    def store(self):
        self.storedLogP = self.logP

    def restore(self):
        self.logP = self.storedLogP


class CompoundDistribution(Distribution):
    """beast.base.inference.CompoundDistribution: sum of children, early return on +-inf / NaN."""
    INPUTS = {"distribution": []}

    def calculateLogP(self) -> float:
        self.logP = 0.0
        for d in self.get("distribution"):
            self.logP += d.calculateLogP() if d.isDirtyCalculation() else d.getCurrentLogP()
            if math.isinf(self.logP) or math.isnan(self.logP):
                return self.logP
        return self.logP

    def store(self):
        super().store()
        for d in self.get("distribution"):
            d.store()

    def restore(self):
        super().restore()
        for d in self.get("distribution"):
            d.restore()

    def accept(self):
        for d in self.get("distribution"):
            d.accept()
