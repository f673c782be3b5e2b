"""Minimal Newick reader producing BEAST2-style node arrays.

Only what the parity tests need to restate the reference's JUnit fixtures, which build their
trees with TreeParser / SpeciesTreeParser from Newick strings (e.g.
src/sb2tests/ConstantPopulationTest.java:46-48).  Conventions follow BEAST2's TreeParser:
leaves are numbered 0..n-1 (in `taxa` order when given, else sorted by label), internal nodes
in post-order so that the root is the last node (quirk Q8), heights are measured from the
deepest tip and, with adjust_tip_heights (TreeParser's default), every tip is snapped to 0.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .problem import TreeArrays


class _N:
    __slots__ = ("label", "length", "children", "nr", "depth")

    def __init__(self):
        self.label = None
        self.length = 0.0
        self.children = []
        self.nr = -1
        self.depth = 0.0


def _parse(s: str) -> _N:
    s = s.strip()
    if s.endswith(";"):
        s = s[:-1]
    pos = 0

    def node() -> _N:
        nonlocal pos
        n = _N()
        if s[pos] == "(":
            pos += 1
            while True:
                n.children.append(node())
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == ")":
                    pos += 1
                    break
                raise ValueError(f"bad newick at {pos}")
        start = pos
        while pos < len(s) and s[pos] not in ",():;":
            pos += 1
        if pos > start:
            n.label = s[start:pos]
        if pos < len(s) and s[pos] == ":":
            pos += 1
            start = pos
            while pos < len(s) and s[pos] not in ",();":
                pos += 1
            n.length = float(s[start:pos])
        return n

    root = node()
    return root


def parse_newick(s: str, taxa: Optional[Sequence[str]] = None,
                 adjust_tip_heights: bool = True) -> Tuple[TreeArrays, List[str]]:
    """Returns (tree arrays, leaf labels by node number)."""
    root = _parse(s)
    leaves: List[_N] = []
    order: List[_N] = []

    def walk(n: _N, depth: float):
        n.depth = depth
        for c in n.children:
            walk(c, depth + c.length)
        if not n.children:
            leaves.append(n)
        order.append(n)

    walk(root, 0.0)
    labels = list(taxa) if taxa is not None else sorted(l.label for l in leaves)
    index: Dict[str, int] = {lab: i for i, lab in enumerate(labels)}
    n_leaves = len(leaves)
    if taxa is not None and len(labels) != n_leaves:
        # a taxon superset may name more leaves than the tree holds (not used by the fixtures)
        labels = [lab for lab in labels if any(l.label == lab for l in leaves)]
        index = {lab: i for i, lab in enumerate(labels)}
    for l in leaves:
        l.nr = index[l.label]
    nxt = n_leaves
    for n in order:
        if n.children:
            if len(n.children) != 2:
                raise ValueError("only binary trees")
            n.nr = nxt
            nxt += 1
    G = nxt
    parent = np.full(G, -1, np.int32)
    left = np.full(G, -1, np.int32)
    right = np.full(G, -1, np.int32)
    height = np.zeros(G)
    max_depth = max(l.depth for l in leaves)
    for n in order:
        height[n.nr] = max_depth - n.depth
        if n.children:
            left[n.nr] = n.children[0].nr
            right[n.nr] = n.children[1].nr
            parent[n.children[0].nr] = n.nr
            parent[n.children[1].nr] = n.nr
    if adjust_tip_heights:
        height[:n_leaves] = 0.0
    return TreeArrays(parent, left, right, height), labels


def leaf_sets(tree: TreeArrays, labels: Sequence[str]) -> List[frozenset]:
    """Leaf-label set under every node (the fixtures address nodes this way, StarbeastClockTest.java:164-189)."""
    G = tree.n_nodes
    out: List[Optional[frozenset]] = [None] * G

    def rec(i: int) -> frozenset:
        if tree.left[i] < 0:
            out[i] = frozenset([labels[i]])
        else:
            out[i] = rec(int(tree.left[i])) | rec(int(tree.right[i]))
        return out[i]

    rec(tree.root)
    return out  # type: ignore[return-value]
