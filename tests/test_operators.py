"""the per-proposal walks over every gene tree of CoordinatedUniform / CoordinatedExponential
(getConnectingNodes, CoordinatedUniform.java:94-177) and NodeReheight2 (recalculateMaxHeight, NodeReheight2.java:150-203).

The reference has no test for these operators (src/sb2tests covers CoordinatedExchange only), so the oracle's literal
restatement of the Java recursions is pinned by properties: a hand-worked case, an independent set-based formulation, and
the operators' own contract (a proposal inside the returned freedoms preserves every topology and keeps all gene trees
compatible with the species tree).  The GPU queries must then equal the oracle bit for bit."""
import math

import numpy as np
import pytest

import oracle
from starbeast2_b200 import synth
from starbeast2_b200.problem import Locus, Problem, TreeArrays


def leaf_sets(t: TreeArrays):
    sets = [None] * t.n_nodes
    order = np.argsort(t.height, kind="stable")
    for i in order:   # children are never above their parent
        sets[i] = frozenset([int(i)]) if t.left[i] < 0 else None
    def rec(i):
        if sets[i] is None:
            sets[i] = rec(t.left[i]) | rec(t.right[i])
        return sets[i]
    rec(t.root)
    return sets


def species_leaves_under(sp: TreeArrays, node: int):
    return leaf_sets(sp)[node]


def connecting_by_sets(pb, species_node):
    """Independent formulation: a gene node is connecting iff its tips all sit under `species_node` and touch both children."""
    sp = pb.species
    sl = leaf_sets(sp)
    L, R = sl[sp.left[species_node]], sl[sp.right[species_node]]
    masks, tip, root = [], math.inf, math.inf
    for loc in pb.loci:
        t = loc.tree
        gs = leaf_sets(t)
        cls = []
        for g in range(t.n_nodes):
            spp = {int(loc.tip_species[x]) for x in gs[g]}
            cls.append("N" if not spp <= (L | R) else "B" if (spp & L and spp & R) else "L" if spp & L else "R")
        m = np.array([c == "B" and t.left[g] >= 0 for g, c in enumerate(cls)], np.uint8)
        for g in np.flatnonzero(m):
            for c in (t.left[g], t.right[g]):
                if cls[c] != "B":
                    tip = min(tip, t.height[g] - t.height[c])
            p = t.parent[g]
            if p >= 0 and cls[p] == "N":
                root = min(root, t.height[p] - t.height[g])
        masks.append(m)
    return np.concatenate(masks), tip, root


def test_oracle_connecting_nodes_hand_worked():
    # species ((A,B)ab,C)root ; gene tree tips a1 a2 b1 c1: ((a1,a2)x,(b1,c1)y)r -- no node descends through both A and B only,
    # except none; then gene tree ((a1,b1)x,(a2,c1)y)r: x = {A,B} -> connecting for species node ab
    sp = TreeArrays([3, 3, 4, 4, -1], [-1, -1, -1, 0, 3], [-1, -1, -1, 1, 2], [0, 0, 0, 1.0, 2.0])
    g = TreeArrays([4, 4, 5, 5, 6, 6, -1], [-1, -1, -1, -1, 0, 2, 4], [-1, -1, -1, -1, 1, 3, 5], [0, 0, 0, 0, 1.5, 2.5, 3.0])
    def pb_with(tip_species):
        loc = Locus(tree=g, tip_species=np.array(tip_species, np.int32), tip_states=np.zeros((4, 1), np.uint8), weights=np.ones(1, np.int32))
        return Problem(species=sp, loci=[loc])
    # tips 0,1 under x; 2,3 under y
    mask, tip, root, n = oracle.connecting_nodes(pb_with([0, 0, 1, 2]), 3)      # x = {A}, y = {B,C}: nothing connects
    assert n == 0 and mask.sum() == 0 and tip == math.inf and root == math.inf
    mask, tip, root, n = oracle.connecting_nodes(pb_with([0, 1, 0, 2]), 3)      # x = {A,B} connects, y = {A,C} does not
    assert n == 1 and list(np.flatnonzero(mask)) == [4]
    assert tip == 1.5 and root == 1.5                                             # x above its tips: 1.5; root r above x: 3.0 - 1.5
    mask, tip, root, n = oracle.connecting_nodes(pb_with([0, 1, 0, 2]), 4)      # species root: left side {A,B}, right side {C}
    assert list(np.flatnonzero(mask)) == [5, 6] and root == math.inf             # x = {A,B} is LEFT_ONLY; y = {A,C} and r connect
    assert tip == 1.5                                                             # y over a2 / c1: 2.5; r over the pure-left x: 3.0 - 1.5
    _, tip_e, root_e, _ = oracle.connecting_nodes(pb_with([0, 1, 0, 2]), 3, exponential=True)
    assert tip_e == 1.5 and root_e == math.inf                                    # CoordinatedExponential never sets the rootward freedom


@pytest.mark.parametrize("name,kw", [("C2", dict(n_loci=6, n_sites=20)), ("C3", dict(n_loci=3, n_sites=20)), ("C4", dict(n_loci=8, n_sites=20))])
def test_oracle_recursion_equals_set_formulation(name, kw):
    pb = synth.make_problem(name, **kw)
    sp = pb.species
    for node in range(sp.n_leaves, sp.n_nodes):
        mask, tip, root, n = oracle.connecting_nodes(pb, node)
        m2, t2, r2 = connecting_by_sets(pb, node)
        assert np.array_equal(mask, m2) and tip == t2 and root == r2 and n == int(m2.sum()), node


def canonical_order(sp: TreeArrays, rng):
    cmap = np.zeros(sp.n_nodes, np.int32)
    heights = []
    def rec(n):
        if sp.left[n] < 0:
            cmap[n] = len(heights); heights.append(sp.height[n]); return
        a, b = (sp.left[n], sp.right[n]) if rng.integers(0, 2) else (sp.right[n], sp.left[n])
        rec(a); cmap[n] = len(heights); heights.append(sp.height[n]); rec(b)
    rec(sp.root)
    return cmap, np.array(heights)


def max_height_by_sets(pb, cmap, center):
    out = math.inf
    for loc in pb.loci:
        t = loc.tree
        gs = leaf_sets(t)
        side = lambda g: {bool(cmap[int(loc.tip_species[x])] >= center) for x in gs[g]}
        for g in range(loc.n_tips, t.n_nodes):
            a, b = side(t.left[g]), side(t.right[g])
            if len(a) == 1 and len(b) == 1 and a != b:
                out = min(out, t.height[g])
    return out


@pytest.mark.parametrize("name,kw", [("C2", dict(n_loci=6, n_sites=20)), ("C4", dict(n_loci=8, n_sites=20))])
def test_oracle_max_height_equals_set_formulation_and_bounds_the_species_node(name, kw):
    pb = synth.make_problem(name, **kw)
    rng = np.random.Generator(np.random.PCG64(5))
    cmap, heights = canonical_order(pb.species, rng)
    for center in range(1, pb.species.n_nodes, 2):
        mh = oracle.max_height(pb, cmap, center)
        assert mh == max_height_by_sets(pb, cmap, center)
        assert mh >= heights[center]          # gene trees are compatible: no cross coalescence below the species node
        assert oracle.max_height(pb, cmap, center, origin=0.5 * mh) == 0.5 * mh


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------

def _engine(pb):
    from starbeast2_b200.engine import Engine
    e = Engine.from_problem(pb)
    e.eval_posterior()
    return e


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("C2", dict(n_loci=50, n_sites=40)), ("C3", dict(n_loci=12, n_sites=40)), ("C4", dict(n_loci=40, n_sites=40)),
                                     ("C5", dict(n_loci=16, n_sites=40))])
def test_gpu_connecting_nodes_bit_identical(name, kw):
    pb = synth.make_problem(name, **kw)
    e = _engine(pb)
    sp = pb.species
    for node in range(sp.n_leaves, sp.n_nodes):
        want = oracle.connecting_nodes(pb, node)
        mask, tip, root, n = e.connecting_nodes(node)
        assert np.array_equal(mask, want[0]), node
        assert (tip, root, n) == want[1:], node
        _, tip2, root2, n2 = e.connecting_nodes(node, want_mask=False)
        assert (tip2, root2, n2) == want[1:]
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("C2", dict(n_loci=50, n_sites=40)), ("C4", dict(n_loci=40, n_sites=40)), ("C5", dict(n_loci=16, n_sites=40))])
def test_gpu_max_height_bit_identical(name, kw):
    pb = synth.make_problem(name, **kw)
    e = _engine(pb)
    rng = np.random.Generator(np.random.PCG64(11))
    for rep in range(3):
        cmap, _ = canonical_order(pb.species, rng)
        for center in range(1, pb.species.n_nodes, 2):
            assert e.max_height(cmap[:pb.species.n_leaves] >= center) == oracle.max_height(pb, cmap, center), (rep, center)
    e.close()


@pytest.mark.gpu
def test_gpu_queries_follow_store_restore_and_refuse_unevaluated_trees():
    from starbeast2_b200._lib import Sb2Error
    from starbeast2_b200.engine import Engine
    pb = synth.make_problem("C2", n_loci=8, n_sites=40)
    e = Engine.from_problem(pb)
    node = pb.species.n_leaves + 2
    with pytest.raises(Sb2Error):
        e.connecting_nodes(node)                       # nothing evaluated yet
    e.eval_posterior()
    base = oracle.connecting_nodes(pb, node)
    e.store()
    rng = np.random.Generator(np.random.PCG64(3))
    t, _ = synth.perturb_gene_tree(rng, pb.loci[2].tree, pb.loci[2].n_tips)
    old = pb.loci[2].tree
    e.set_gene_tree(2, t)
    with pytest.raises(Sb2Error):
        e.connecting_nodes(node)                       # staged, not evaluated
    e.eval_posterior()
    pb.loci[2].tree = t
    got = e.connecting_nodes(node)
    want = oracle.connecting_nodes(pb, node)
    assert np.array_equal(got[0], want[0]) and got[1:] == want[1:]
    e.restore()
    pb.loci[2].tree = old
    got = e.connecting_nodes(node)
    assert np.array_equal(got[0], base[0]) and got[1:] == base[1:]
    with pytest.raises(Sb2Error):
        e.connecting_nodes(0)                          # a leaf
    e.close()


def _topology(t: TreeArrays):
    return sorted(tuple(sorted(s)) for s in leaf_sets(t))


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("C2", dict(n_loci=6, n_sites=60)), ("C3", dict(n_loci=3, n_sites=60))])
def test_operators_preserve_topologies_and_chain_matches_oracle(name, kw):
    """CoordinatedUniform / CoordinatedExponential / NodeReheight2 driven through the plugin protocol: every proposal keeps
    all gene trees compatible (finite coalescent), co-ordinated moves keep every topology, and the evaluated posterior of
    each proposed state equals the oracle's from-scratch value (1e-9), including the accept/reject decision."""
    from test_mcmc_trace import build
    from starbeast2_b200.operators import CoordinatedExponential, CoordinatedUniform, NodeReheight2, Randomizer
    from starbeast2_b200.problem import POP_CONSTANT
    pb = synth.make_problem(name, **kw)
    speciesTree, trees, params, posterior = build(pb)
    msc = posterior.get("distribution")[0]
    geneTrees = msc.get("distribution")
    rnd = Randomizer(99)
    ops = [CoordinatedUniform(speciesTree=speciesTree, geneTree=trees, random=rnd),
           CoordinatedExponential(speciesTree=speciesTree, geneTree=trees, random=rnd, beta=0.05),
           NodeReheight2(tree=speciesTree, geneTree=geneTrees, random=rnd, window=0.5)]
    state_nodes = [speciesTree] + trees + params
    cur = posterior.calculateLogP()
    assert math.isfinite(cur)
    rng = np.random.Generator(np.random.PCG64(4))
    accepted = {type(o).__name__: 0 for o in ops}
    for step in range(90):
        for sn in state_nodes:
            sn.store()
        posterior.store()
        op = ops[step % 3]
        topo_before = [_topology(t.arrays) for t in trees]
        sp_topo_before = _topology(speciesTree.arrays)
        hr = op.proposal()
        assert math.isfinite(hr)
        new = posterior.calculateLogP()
        assert math.isfinite(new), (step, type(op).__name__)          # the freedoms / max height keep every gene tree compatible
        assert [_topology(t.arrays) for t in trees] == topo_before     # gene topologies never change
        if not isinstance(op, NodeReheight2):
            assert _topology(speciesTree.arrays) == sp_topo_before
        sp = speciesTree.arrays
        assert sp.root == sp.n_nodes - 1                               # quirk Q8
        for t in [sp] + [t.arrays for t in trees]:
            internal = t.left >= 0
            assert np.all(t.height[internal] >= t.height[t.left[internal]]) and np.all(t.height[internal] >= t.height[t.right[internal]])
        pb.species = speciesTree.arrays
        for l, t in enumerate(trees):
            pb.loci[l].tree = t.arrays
        want = oracle.posterior(pb).total
        assert new == pytest.approx(want, rel=1e-9), (step, type(op).__name__)
        logu = math.log(rng.uniform())
        assert (logu < new - cur + hr) == (logu < want - cur + hr)
        if logu < new - cur + hr:
            for sn in state_nodes:
                sn.accept()
            posterior.accept()
            cur = new
            accepted[type(op).__name__] += 1
        else:
            for sn in state_nodes:
                sn.restore()
            posterior.restore()
            assert posterior.calculateLogP() == cur
    assert all(v > 0 for v in accepted.values()), accepted
