"""The BEAST.base half of the oracle is 'parity unpinned' (no reference source, no reference test).
It is therefore checked against independent mathematics: scipy.linalg.expm for P(t), brute-force
enumeration of internal states on small trees, and the JC69 closed form."""
import itertools
import math

import numpy as np
import pytest
from scipy.linalg import expm

import oracle
from starbeast2_b200 import synth
from starbeast2_b200.newick import parse_newick
from starbeast2_b200.problem import (SUBST_EIGEN, SUBST_HKY, Locus, const_pattern_mask, gtr_eigen_params, hky_params)


def hky_q(kappa, pi):
    q = np.array([[0, 1, kappa, 1], [1, 0, 1, kappa], [kappa, 1, 0, 1], [1, kappa, 1, 0]], float) * np.asarray(pi)[None, :]
    np.fill_diagonal(q, -q.sum(1))
    return q / -(np.asarray(pi) * np.diag(q)).sum()


def gtr_q(r, pi):
    ac, ag, at, cg, ct, gt = r
    q = np.array([[0, ac, ag, at], [ac, 0, cg, ct], [ag, cg, 0, gt], [at, ct, gt, 0]], float) * np.asarray(pi)[None, :]
    np.fill_diagonal(q, -q.sum(1))
    return q / -(np.asarray(pi) * np.diag(q)).sum()


@pytest.mark.parametrize("d", [0.0, 1e-6, 0.01, 0.3, 2.5, 40.0])
def test_hky_matrix_matches_expm(d):
    pi = np.array([0.30, 0.20, 0.20, 0.30])
    got = oracle.hky_matrix(2.7, pi, d)
    np.testing.assert_allclose(got, expm(hky_q(2.7, pi) * d), rtol=0, atol=1e-13)


@pytest.mark.parametrize("d", [0.0, 0.01, 0.3, 2.5])
def test_eigen_matrix_matches_expm(d):
    pi = np.array([0.28, 0.22, 0.21, 0.29])
    r = (1, 3, 0.8, 1.2, 3.5, 1)
    sp = gtr_eigen_params(r, pi)
    got = oracle.eigen_matrix(sp[0:4], sp[4:20], sp[20:36], d)
    np.testing.assert_allclose(got, expm(gtr_q(r, pi) * d), rtol=0, atol=1e-13)


def brute_force_logl(locus, branch_rates, q):
    t = locus.tree
    G, nT = locus.n_nodes, locus.n_tips
    pi = locus.subst_params[1:5] if locus.subst_kind == SUBST_HKY else locus.subst_params[36:40]
    internals = list(range(nT, G))
    total = 0.0
    pat_ll = []
    for p in range(locus.n_pat):
        lk = 0.0
        for c in range(locus.n_cat):
            P = {i: expm(q * (t.height[t.parent[i]] - t.height[i]) * branch_rates[i] * locus.cat_rates[c])
                 for i in range(G) if t.parent[i] >= 0}
            s = 0.0
            for assign in itertools.product(range(4), repeat=len(internals)):
                st = dict(zip(internals, assign))
                w = pi[st[t.root]]
                for i in range(G):
                    if t.parent[i] < 0:
                        continue
                    if i < nT:
                        x = locus.tip_states[i, p]
                        if x < 4:
                            w *= P[i][st[t.parent[i]], x]
                    else:
                        w *= P[i][st[t.parent[i]], st[i]]
                s += w
            lk += locus.cat_props[c] * s
        pat_ll.append(math.log(lk))
        total += math.log(lk) * locus.weights[p]
    return total, np.array(pat_ll)


@pytest.mark.parametrize("kind,ncat", [("hky", 1), ("hky", 3), ("gtr", 2)])
def test_tree_likelihood_matches_brute_force(kind, ncat):
    rng = np.random.default_rng(5)
    tree, labels = parse_newick("(((a:0.1,b:0.2):0.15,c:0.3):0.05,(d:0.22,e:0.07):0.4);", adjust_tip_heights=False)
    nT = 5
    pats = rng.integers(0, 4, size=(nT, 12)).astype(np.uint8)
    pats[1, 3] = 4   # missing
    pats[:, 5] = 2   # constant column
    pats[0, 7] = 4
    pats[4, 7] = 4
    w = rng.integers(1, 5, size=12).astype(np.int32)
    pi = np.array([0.30, 0.20, 0.20, 0.30])
    if kind == "hky":
        sk, spar, q = SUBST_HKY, hky_params(2.0, pi), hky_q(2.0, pi)
    else:
        r = (1, 3, 0.8, 1.2, 3.5, 1)
        sk, spar, q = SUBST_EIGEN, gtr_eigen_params(r, pi), gtr_q(r, pi)
    cat_rates = synth.discrete_gamma_rates(0.5, ncat) * 1.3
    loc = Locus(tree=tree, tip_states=pats, weights=w, tip_species=np.zeros(nT, np.int32), subst_kind=sk,
                subst_params=spar, cat_rates=cat_rates, cat_props=np.full(ncat, 1.0 / ncat))
    br = rng.uniform(0.5, 2.0, size=tree.n_nodes)
    want, want_pat = brute_force_logl(loc, br, q)
    for mode in (0, 1, 2):
        got, det = oracle.tree_likelihood(loc, br, use_scaling=mode, want_details=True)
        assert got == pytest.approx(want, rel=1e-12)
        np.testing.assert_allclose(det["pattern_logl"], want_pat, rtol=1e-12)


def test_jc69_two_tip_closed_form():
    # two tips, distance d: same state 1/4(1/4+3/4 e^{-4d/3}), different 1/4(1/4-1/4 e^{-4d/3})
    tree, _ = parse_newick("(a:0.35,b:0.35);")
    loc = Locus(tree=tree, tip_states=np.array([[0, 0], [0, 1]], np.uint8), weights=np.array([3, 2], np.int32),
                tip_species=np.zeros(2, np.int32), subst_params=hky_params(1.0, [0.25] * 4))
    d = 0.7
    e = math.exp(-4 * d / 3)
    want = 3 * math.log(0.25 * (0.25 + 0.75 * e)) + 2 * math.log(0.25 * (0.25 - 0.25 * e))
    assert oracle.tree_likelihood(loc, np.ones(3)) == pytest.approx(want, rel=1e-13)


def test_scaling_rescues_underflow():
    # a deep caterpillar with many sites per pattern product: unscaled underflows to -inf, lazy scaling recovers
    import sys
    sys.setrecursionlimit(10000)
    rng = np.random.default_rng(11)
    n = 640
    nw = "t0:1.0"
    for i in range(1, n):
        nw = f"({nw},t{i}:{1.0 + i}):1.0"
    tree, labels = parse_newick(nw + ";", taxa=[f"t{i}" for i in range(n)], adjust_tip_heights=False)
    pats = rng.integers(0, 4, size=(n, 6)).astype(np.uint8)
    loc = Locus(tree=tree, tip_states=pats, weights=np.ones(6, np.int32), tip_species=np.zeros(n, np.int32),
                subst_params=hky_params(2.0, [0.3, 0.2, 0.2, 0.3]))
    br = np.ones(tree.n_nodes)
    assert oracle.tree_likelihood(loc, br, use_scaling=0) == -math.inf
    lazy = oracle.tree_likelihood(loc, br, use_scaling=2)
    always = oracle.tree_likelihood(loc, br, use_scaling=1)
    assert math.isfinite(lazy) and lazy == pytest.approx(always, rel=1e-12)


def test_pinv_constant_patterns():
    tree, _ = parse_newick("((a:0.1,b:0.1):0.1,c:0.2);")
    pats = np.array([[0, 1, 4], [0, 2, 3], [0, 1, 3]], np.uint8)
    assert list(const_pattern_mask(pats)) == [1, 0, 8]
    loc = Locus(tree=tree, tip_states=pats, weights=np.ones(3, np.int32), tip_species=np.zeros(3, np.int32),
                subst_params=hky_params(2.0, [0.3, 0.2, 0.2, 0.3]), p_inv=0.2,
                cat_rates=np.array([1.25]), cat_props=np.array([0.8]))
    ll, det = oracle.tree_likelihood(loc, np.ones(5), want_details=True)
    loc0 = Locus(tree=tree, tip_states=pats, weights=np.ones(3, np.int32), tip_species=np.zeros(3, np.int32),
                 subst_params=hky_params(2.0, [0.3, 0.2, 0.2, 0.3]), cat_rates=np.array([1.25]), cat_props=np.array([1.0]))
    _, det0 = oracle.tree_likelihood(loc0, np.ones(5), want_details=True)
    # pattern 0 is constant in A: L = 0.8*L0 + 0.2*pi_A ; pattern 1 is variable: L = 0.8*L0
    assert math.exp(det["pattern_logl"][0]) == pytest.approx(0.8 * math.exp(det0["pattern_logl"][0]) + 0.2 * 0.3, rel=1e-13)
    assert math.exp(det["pattern_logl"][1]) == pytest.approx(0.8 * math.exp(det0["pattern_logl"][1]), rel=1e-13)


def test_posterior_driver_consistent_with_pieces():
    pb = synth.make_problem("C3", n_loci=3, n_sites=60)
    res = oracle.posterior(pb, n_threads=2)
    assert math.isfinite(res.total)
    for l, loc in enumerate(pb.loci):
        emb = oracle.gene_tree_update(loc.tree, loc.tip_species, pb.species)
        assert emb.compatible
        rates = oracle.starbeast_clock(loc.clock_rate, pb.species_rates, emb.occupancy)
        assert oracle.tree_likelihood(loc, rates) == pytest.approx(res.per_locus_lik[l], rel=1e-14)
        coal = sum(oracle.lwcr_branch_logp(b, pb.species, pb.pop_a, pb.pop_b, loc.ploidy, emb.coalescent_times(b), emb.n[b], emb.k[b])
                   for b in range(pb.species.n_nodes))
        assert coal == pytest.approx(res.per_locus_coal[l], rel=1e-13)
    assert res.total == pytest.approx(res.coalescent + res.likelihood)


# ---------------------------------------------------------------------------------------------------------------------
# Size-independent properties of the pruning algorithm (the BEAST.base half is parity-unpinned: these pin it from the
# mathematics side at realistic tree sizes, where brute-force enumeration is impossible)
# ---------------------------------------------------------------------------------------------------------------------

def _random_locus(rng, n_tips=24, n_pat=40, kind="gtr", ncat=4, missing=0.05):
    from starbeast2_b200.problem import TreeArrays
    # random binary tree by successive joins at increasing heights (node numbers: leaves first, root last)
    G = 2 * n_tips - 1
    parent, left, right, height = np.full(G, -1), np.full(G, -1), np.full(G, -1), np.zeros(G)
    active = list(range(n_tips))
    h = 0.0
    for node in range(n_tips, G):
        i, j = rng.choice(len(active), 2, replace=False)
        a, b = active[i], active[j]
        h += rng.exponential(0.03)
        left[node], right[node], height[node] = a, b, h
        parent[a] = parent[b] = node
        active = [x for x in active if x not in (a, b)] + [node]
    tree = TreeArrays(parent, left, right, height)
    pats = rng.integers(0, 4, size=(n_tips, n_pat)).astype(np.uint8)
    pats[rng.random(pats.shape) < missing] = 4
    w = rng.integers(1, 4, size=n_pat).astype(np.int32)
    pi = np.array([0.31, 0.19, 0.22, 0.28])
    if kind == "hky":
        sk, spar = SUBST_HKY, hky_params(2.7, pi)
    else:
        sk, spar = SUBST_EIGEN, gtr_eigen_params((1, 3, 0.8, 1.2, 3.5, 1), pi)
    rates = synth.discrete_gamma_rates(0.5, ncat) * 0.9
    return Locus(tree=tree, tip_states=pats, weights=w, tip_species=np.zeros(n_tips, np.int32), subst_kind=sk, subst_params=spar,
                 cat_rates=rates, cat_props=np.full(ncat, 1.0 / ncat))


@pytest.mark.parametrize("kind,ncat", [("hky", 1), ("gtr", 4)])
def test_pulley_principle_root_position_is_irrelevant_for_reversible_models(kind, ncat):
    """Felsenstein's pulley principle: for a time-reversible model only the SUM of the two root branch lengths matters.
    Sliding the root along its branch (explicit per-branch rates stretch one root branch and shrink the other) leaves every
    pattern likelihood unchanged."""
    rng = np.random.default_rng(17)
    loc = _random_locus(rng, kind=kind, ncat=ncat)
    t = loc.tree
    br = np.ones(t.n_nodes)
    base, det = oracle.tree_likelihood(loc, br, want_details=True)
    a, b = t.left[t.root], t.right[t.root]
    la, lb = t.height[t.root] - t.height[a], t.height[t.root] - t.height[b]
    for shift in (0.3, -0.25, 0.9):
        delta = shift * min(la, lb)
        br2 = br.copy()
        br2[a] = (la + delta) / la        # length * rate: the branch above a becomes la + delta
        br2[b] = (lb - delta) / lb        # ... and the one above b becomes lb - delta
        got, det2 = oracle.tree_likelihood(loc, br2, want_details=True)
        assert got == pytest.approx(base, rel=1e-12)
        np.testing.assert_allclose(det2["pattern_logl"], det["pattern_logl"], rtol=1e-11)


def test_all_missing_tip_and_pattern_permutation_do_not_change_the_likelihood():
    rng = np.random.default_rng(23)
    loc = _random_locus(rng, n_tips=16, n_pat=30, kind="gtr", ncat=2)
    br = rng.uniform(0.6, 1.6, loc.tree.n_nodes)
    base, det = oracle.tree_likelihood(loc, br, want_details=True)
    # permuting the patterns permutes the pattern likelihoods and leaves the total unchanged up to summation order
    perm = rng.permutation(loc.n_pat)
    loc2 = Locus(tree=loc.tree, tip_states=loc.tip_states[:, perm], weights=loc.weights[perm], tip_species=loc.tip_species,
                 subst_kind=loc.subst_kind, subst_params=loc.subst_params, cat_rates=loc.cat_rates, cat_props=loc.cat_props)
    got2, det2 = oracle.tree_likelihood(loc2, br, want_details=True)
    assert np.array_equal(det2["pattern_logl"], det["pattern_logl"][perm])
    assert got2 == pytest.approx(base, rel=1e-13)
    # a tip that is missing everywhere contributes a factor 1: same as pruning it (its sibling then hangs on the summed branch)
    from starbeast2_b200.problem import TreeArrays
    t = loc.tree
    tip = 3
    blank = loc.tip_states.copy()
    blank[tip, :] = 4
    loc3 = Locus(tree=t, tip_states=blank, weights=loc.weights, tip_species=loc.tip_species, subst_kind=loc.subst_kind,
                 subst_params=loc.subst_params, cat_rates=loc.cat_rates, cat_props=loc.cat_props)
    with_blank = oracle.tree_likelihood(loc3, np.ones(t.n_nodes))
    # build the pruned tree: remove `tip` and its parent p; the sibling s attaches to p's parent (heights keep branch sums)
    p = t.parent[tip]
    s = t.right[p] if t.left[p] == tip else t.left[p]
    gp = t.parent[p]
    assert gp >= 0
    keep = [i for i in range(t.n_nodes) if i not in (tip, p)]
    renum = {old: new for new, old in enumerate(keep)}
    par = np.array([-1 if t.parent[i] < 0 else renum[gp if t.parent[i] == p else t.parent[i]] for i in keep])
    def child(c):
        return -1 if c < 0 else renum[s if c == p else c]
    lef = np.array([child(t.left[i]) for i in keep])
    rig = np.array([child(t.right[i]) for i in keep])
    pruned = TreeArrays(par, lef, rig, t.height[keep])
    loc4 = Locus(tree=pruned, tip_states=np.delete(loc.tip_states, tip, axis=0), weights=loc.weights,
                 tip_species=np.zeros(loc.n_tips - 1, np.int32), subst_kind=loc.subst_kind, subst_params=loc.subst_params,
                 cat_rates=loc.cat_rates, cat_props=loc.cat_props)
    assert oracle.tree_likelihood(loc4, np.ones(pruned.n_nodes)) == pytest.approx(with_blank, rel=1e-12)


@pytest.mark.parametrize("name,kw", [("C2", dict(n_loci=5, n_sites=120)), ("C3", dict(n_loci=3, n_sites=90))])
def test_incremental_baseline_equals_full_evaluation(name, kw):
    """The regime-B CPU baseline (embedding of every locus + incremental TreeLikelihood of the changed one) returns exactly
    what a from-scratch evaluation of the proposed state returns, move after move, and -inf for an incompatible tree."""
    import math
    pb = synth.make_problem(name, **kw)
    inc = oracle.IncrementalBaseline(pb)
    rng = np.random.default_rng(2)
    for step in range(12):
        l = int(rng.integers(0, pb.n_loci))
        t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
        pb.loci[l].tree = t
        got = inc.step(l, t)
        want = oracle.posterior(pb, use_scaling=0)
        if math.isfinite(want.coalescent):
            assert got == pytest.approx(want.total, rel=1e-13)
        else:
            assert got == -math.inf
    inc.close()
