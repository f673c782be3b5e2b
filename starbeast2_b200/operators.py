"""Host-side mirror of the reference's species-tree operators whose per-proposal cost is a walk over EVERY gene tree
(SURVEY.md 8f rank 2): CoordinatedUniform, CoordinatedExponential (Jones 2015 co-ordinated moves) and NodeReheight2.

Same Input names and proposal() contract as /root/reference/src/starbeast2/{CoordinatedOperator, CoordinatedUniform,
CoordinatedExponential, NodeReheight2}.java.  The gene-tree walks themselves -- getConnectingNodes / findConnectingNodes
(CoordinatedUniform.java:94-177, CoordinatedExponential.java:92-160) and recalculateMaxHeight / recurseMaxHeight
(NodeReheight2.java:150-203) -- are NOT done here: they are one libsb2gpu query each (sb2_connecting_nodes,
sb2_max_height) answered from the trees already resident on the device.  Everything else (choice of node, random draws,
canonical order, rebuilding the species tree) is the cheap O(S) host part and stays on the host, like in the Java twins.

Operators run between accept()/restore() of one step and store() of the next, so the device's current state is the chain's
current state; the operators make sure of it by evaluating the (cached) coalescent phase first.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np

from .beast import BEASTObject, IS_FILTHY
from .plugins import GeneTree, GpuSession, SpeciesTree

REQ = BEASTObject.REQUIRED
__all__ = ["Randomizer", "Operator", "CoordinatedUniform", "CoordinatedExponential", "NodeReheight2"]


class Randomizer:
    """beast.base.util.Randomizer's call surface over a numpy Generator (not the same stream as the Java MersenneTwister)."""

    def __init__(self, seed: int = 0):
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def nextInt(self, n: int) -> int:
        return int(self.rng.integers(0, n))

    def nextDouble(self) -> float:
        return float(self.rng.random())

    def nextBoolean(self) -> bool:
        return bool(self.rng.integers(0, 2))

    def nextExponential(self, lam: float) -> float:
        return float(self.rng.exponential(1.0 / lam))


class Operator(BEASTObject):
    INPUTS = {"weight": 1.0, "random": None}

    def randomizer(self) -> Randomizer:
        r = self.get("random")
        if r is None:
            r = Randomizer(0)
            self._inputs["random"] = r
        return r

    def proposal(self) -> float:   # log Hastings ratio, -inf = reject outright
        raise NotImplementedError


def _is_fake(sp, node: int) -> bool:
    """Node.isFake(): an internal node one of whose children is a sampled ancestor (a leaf on a zero-length branch)."""
    l, r = sp.left[node], sp.right[node]
    if l < 0:
        return False
    return any(sp.left[c] < 0 and sp.height[c] == sp.height[node] for c in (l, r))


class CoordinatedOperator(Operator):
    """CoordinatedOperator.java:15-25: inputs speciesTree, geneTree (list of gene Trees)."""
    INPUTS = {"speciesTree": REQ, "geneTree": []}

    def initAndValidate(self):
        self.speciesTree: SpeciesTree = self.get("speciesTree")
        self.geneTrees = list(self.get("geneTree"))
        self.nGeneTrees = len(self.geneTrees)
        self.session = GpuSession.of(self.speciesTree)

    def _connecting(self, species_node: int):
        """getConnectingNodes for every gene tree: (per gene tree arrays of node numbers, tipward, rootward)."""
        s = self.session
        s.coalescent()   # brings the device to the chain's current state (cached when nothing changed)
        # the operator's geneTree list may be a subset / reordering of the session's loci (it is the same list in every shipped XML)
        mask, tipward, rootward, _ = s.engine.connecting_nodes(species_node)
        off = np.concatenate([[0], np.cumsum([gt.get("tree").getNodeCount() for gt in s.gene_trees])])
        nodes = []
        for tree in self.geneTrees:
            l = s.locus_of_tree(tree)
            if l is None:
                raise ValueError("every geneTree of a co-ordinated operator must belong to a GeneTree of the session")
            nodes.append(np.flatnonzero(mask[off[l]:off[l + 1]]))
        if len(self.geneTrees) != len(s.gene_trees):
            raise ValueError("co-ordinated operators must list every gene tree of the species tree")
        return nodes, tipward, rootward

    def _shift(self, species_node: int, nodes, shift: float):
        sp = self.speciesTree
        sp.setHeight(species_node, sp.arrays.height[species_node] + shift)
        for tree, idx in zip(self.geneTrees, nodes):
            for g in idx:
                tree.setHeight(int(g), tree.arrays.height[g] + shift)


class CoordinatedUniform(CoordinatedOperator):
    """CoordinatedUniform.java:41-90."""

    def proposal(self) -> float:
        sp = self.speciesTree.arrays
        rnd = self.randomizer()
        cands = [i for i in range(sp.n_leaves, sp.n_nodes) if not (_is_fake(sp, i) or sp.parent[i] < 0)]   # :52-58
        if not cands:
            return -math.inf
        node = cands[rnd.nextInt(len(cands))]
        h = sp.height[node]
        nodes, twf, rwf = self._connecting(node)
        twf = min(twf, h - sp.height[sp.left[node]], h - sp.height[sp.right[node]])   # :76-80
        rwf = min(rwf, sp.height[sp.parent[node]] - h)                                 # :78,:81
        shift = (rnd.nextDouble() * (twf + rwf)) - twf                                  # :85
        self._shift(node, nodes, shift)
        return 0.0


class CoordinatedExponential(CoordinatedOperator):
    """CoordinatedExponential.java:56-87, 169-192."""
    INPUTS = {"beta": 1.0, "optimise": True}
    waitingTimeScale = 1.4426950408889634

    def initAndValidate(self):
        super().initAndValidate()
        self.beta = float(self.get("beta"))
        self.lam = 1.0 / self.beta
        self.optimise = bool(self.get("optimise"))
        self.waitingTime = 0.0
        self.nAccepted = self.nRejected = 0

    def proposal(self) -> float:
        sp = self.speciesTree.arrays
        root = sp.root
        if _is_fake(sp, root):
            return -math.inf
        h = sp.height[root]
        nodes, twf, _ = self._connecting(root)
        twf = min(twf, h - sp.height[sp.left[root]], h - sp.height[sp.right[root]])
        self.waitingTime = twf * self.waitingTimeScale
        shift = self.randomizer().nextExponential(self.lam) - twf
        self._shift(root, nodes, shift)
        return self.lam * shift

    def getCoercableParameterValue(self) -> float:
        return self.beta

    def setCoercableParameterValue(self, v: float):
        self.beta = v
        self.lam = 1.0 / v

    def optimize(self, logAlpha: float):   # :184-191
        if self.optimise:
            count = self.nRejected + self.nAccepted + 1.0
            self.setCoercableParameterValue(self.beta + (self.waitingTime - self.beta) / count)


class NodeReheight2(Operator):
    """NodeReheight2.java:23-27 inputs (tree, taxonset, geneTree: List<GeneTree>, window = 10, origin); proposal :79-147."""
    INPUTS = {"tree": REQ, "taxonset": None, "geneTree": [], "window": 10.0, "origin": None}

    def initAndValidate(self):
        self.tree: SpeciesTree = self.get("tree")
        self.geneTrees: List[GeneTree] = list(self.get("geneTree"))
        self.window = float(self.get("window"))
        self.session = GpuSession.of(self.tree)
        self.maxHeight = math.inf

    # -- chooseCanonicalOrder (:209-262): in-order traversal with random left/right flips ---------------------------------
    def _canonical(self, sp, rnd):
        n = sp.n_nodes
        order = np.zeros(n, np.int64)
        cmap = np.zeros(n, np.int32)
        heights = np.zeros(n)
        true_bif: List[int] = []
        nxt = [0]

        def place(node):
            i = nxt[0]
            nxt[0] += 1
            cmap[node] = i
            order[i] = node
            heights[i] = sp.height[node]
            return i

        def rec(node) -> float:
            a, b = (sp.left[node], sp.right[node]) if rnd.nextBoolean() else (sp.right[node], sp.left[node])
            if sp.left[a] < 0:
                place(a)
                ha = sp.height[a]
            else:
                ha = rec(a)
            i = place(node)
            if sp.left[b] < 0:
                place(b)
                hb = sp.height[b]
            else:
                hb = rec(b)
            if sp.height[node] > ha and sp.height[node] > hb:
                true_bif.append(i)
            return sp.height[node]

        rec(sp.root)
        return order, cmap, heights, true_bif

    def recalculateMaxHeight(self, cmap, center: int) -> float:
        """:150-173 -- the walk over all gene trees is one device query."""
        origin = self.get("origin")
        self.maxHeight = float(origin.getValue()) if origin is not None else math.inf
        s = self.session
        s.coalescent()
        if len(self.geneTrees) != len(s.gene_trees):
            raise ValueError("NodeReheight2 must list every GeneTree of the species tree")
        n_leaves = self.tree.arrays.n_leaves
        self.maxHeight = min(self.maxHeight, s.engine.max_height(cmap[:n_leaves] >= center))
        return self.maxHeight

    def proposal(self) -> float:
        sp = self.tree.arrays
        rnd = self.randomizer()
        n = sp.n_nodes
        order, cmap, heights, true_bif = self._canonical(sp, rnd)
        if not true_bif:
            return -math.inf
        chosen = true_bif[rnd.nextInt(len(true_bif))]
        original = heights[chosen]
        lo = max(heights[chosen - 1], heights[chosen + 1])            # :101
        hi = self.recalculateMaxHeight(cmap, chosen)
        delta = math.fmod(self.window * (rnd.nextDouble() - 0.5), 2.0 * (hi - lo))   # :108 (Java % on doubles = fmod)
        new = original + delta
        while new < lo or new > hi:                                   # reflection, :110-117
            if new < lo:
                new = lo + lo - new
            if new > hi:
                new = hi + hi - new
        heights[chosen] = new

        parent = np.full(n, -1, np.int64)     # by canonical index
        left = np.full(n, -1, np.int64)
        right = np.full(n, -1, np.int64)
        superimposed = [False]

        def rebuild(frm: int, to: int) -> int:   # :265-308, inclusive bounds, internal nodes sit at odd indices
            best, idx = 0.0, -1
            for i in range(frm + 1, to, 2):
                if heights[i] > best:
                    best, idx = heights[i], i
                elif heights[i] == best:
                    superimposed[0] = True
            li = frm if frm == idx - 1 else rebuild(frm, idx - 1)
            parent[li] = idx
            left[idx] = li
            ri = to if idx + 1 == to else rebuild(idx + 1, to)
            parent[ri] = idx
            right[idx] = ri
            return idx

        root_idx = rebuild(0, n - 1)
        if superimposed[0]:
            return -math.inf

        # write back by node number; Tree.setRoot keeps the root in the last slot by swapping node numbers (quirk Q8)
        new_arrays = sp.copy()
        for i in range(n):
            node = order[i]
            new_arrays.parent[node] = order[parent[i]] if parent[i] >= 0 else -1
            if i % 2 == 1:
                new_arrays.left[node] = order[left[i]]
                new_arrays.right[node] = order[right[i]]
        new_arrays.height[order[chosen]] = new
        new_root, last = int(order[root_idx]), n - 1
        if new_root != last:
            _swap_numbers(new_arrays, new_root, last)
        self.tree.assignFrom(new_arrays)
        self.tree.node_dirty[:] = IS_FILTHY
        return 0.0


def _swap_numbers(t, a: int, b: int):
    """Exchange the node numbers of two internal nodes of a flat tree."""
    for arr in (t.parent, t.left, t.right):
        ia, ib = arr == a, arr == b
        arr[ia], arr[ib] = b, a
    for arr in (t.parent, t.left, t.right, t.height):
        arr[a], arr[b] = arr[b], arr[a]
