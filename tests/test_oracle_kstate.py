"""The K-state half of the oracle (oracle/sb2_oracle.c section 1c'): generic BeerLikelihoodCore / GeneralSubstitutionModel /
ascertainment correction, restated for the morphological TreeLikelihoods of tutorial/CanisFBD/Canis-FBD.xml:964-1033.

BEAST.base and the morph-models package are not under /root/reference and no reference test builds such a likelihood, so
this half is PARITY UNPINNED against the Java; it is pinned by independent mathematics instead: the Mk closed form,
brute-force summation over internal states, the 4-state oracle (K = 4 must reproduce it bit for bit), scipy's expm."""
import itertools
import math

import numpy as np
import pytest

import oracle
from starbeast2_b200 import synth
from starbeast2_b200.problem import MorphPartition, TreeArrays, eigen_params_k, lewis_mk_eigen


def mk_closed_form(K, t):
    e = math.exp(-K * t / (K - 1.0))
    p = np.full((K, K), 1.0 / K - e / K)
    np.fill_diagonal(p, 1.0 / K + (K - 1.0) * e / K)
    return p


@pytest.mark.parametrize("K", [2, 3, 4, 5])
def test_lewis_mk_transition_matrices_match_the_closed_form(K):
    ev, vec, ivec, pi = oracle.split_eigen_k(K, lewis_mk_eigen(K))
    assert np.allclose(pi, 1.0 / K)
    for t in (0.0, 1e-6, 0.03, 0.4, 2.5, 40.0):
        p = oracle.eigen_matrix_k(K, ev, vec, ivec, t)
        assert np.allclose(p, mk_closed_form(K, t), rtol=1e-12, atol=1e-15)
        assert np.allclose(p.sum(axis=1), 1.0, rtol=1e-13)


def test_general_k_state_matrix_matches_expm():
    from scipy.linalg import expm
    rng = np.random.default_rng(3)
    K = 5
    pi = rng.dirichlet(np.ones(K) * 5)
    r = rng.uniform(0.3, 3.0, size=(K, K))
    r = (r + r.T) / 2
    q = r * pi[None, :]
    np.fill_diagonal(q, 0.0)
    np.fill_diagonal(q, -q.sum(axis=1))
    d = np.sqrt(pi)
    w, u = np.linalg.eigh((d[:, None] * q) / d[None, :])
    vec, ivec = u / d[:, None], u.T * d[None, :]
    for t in (0.01, 0.5, 3.0):
        assert np.allclose(oracle.eigen_matrix_k(K, w, vec, ivec, t), expm(q * t), rtol=1e-11, atol=1e-14)


def five_tip_tree():
    # ((0:0.1,1:0.1)5:0.25,((2:0.05,3:0.05)6:0.15,4:0.2)7:0.15)8 ; node 4 is a fossil-like tip above the present
    parent = [5, 5, 6, 6, 7, 8, 7, 8, -1]
    left = [-1, -1, -1, -1, -1, 0, 2, 6, 5]
    right = [-1, -1, -1, -1, -1, 1, 3, 4, 7]
    height = [0.0, 0.0, 0.0, 0.0, 0.07, 0.1, 0.05, 0.27, 0.42]
    return TreeArrays(parent, left, right, height)


def brute_force_pattern_logl(tree, part):
    K = part.n_states
    ev, vec, ivec, pi = oracle.split_eigen_k(K, part.subst_params)
    nT, G = part.n_tips, tree.n_nodes
    internal = list(range(nT, G))
    out = np.zeros(part.n_pat)
    for p in range(part.n_pat):
        total = 0.0
        for c in range(part.n_cat):
            mats = {i: oracle.eigen_matrix_k(K, ev, vec, ivec, (tree.height[tree.parent[i]] - tree.height[i]) * part.clock_rate * part.cat_rates[c])
                    for i in range(G) if tree.parent[i] >= 0}
            s_c = 0.0
            for assign in itertools.product(range(K), repeat=len(internal)):
                st = dict(zip(internal, assign))
                pr = pi[st[tree.root]]
                for i in range(G):
                    if tree.parent[i] < 0:
                        continue
                    a = st[int(tree.parent[i])]
                    if i < nT:
                        b = int(part.tip_states[i, p])
                        pr *= mats[i][a, b] if b < K else 1.0
                    else:
                        pr *= mats[i][a, st[i]]
                s_c += pr
            total += part.cat_props[c] * s_c
        out[p] = math.log(total)
    return out


@pytest.mark.parametrize("K", [2, 3, 5])
def test_pruning_equals_brute_force_over_internal_states(K):
    rng = np.random.default_rng(10 + K)
    tree = five_tip_tree()
    n_pat = 9
    states = rng.integers(0, K, size=(5, n_pat)).astype(np.uint8)
    states[1, 2] = K          # missing
    states[4, 5] = K + 3      # any code >= K is missing
    states[:, 8] = K          # an all-missing character: likelihood 1
    part = MorphPartition(tip_states=states, weights=rng.integers(1, 4, size=n_pat), n_states=K, subst_params=lewis_mk_eigen(K),
                          cat_rates=[0.3, 1.7], cat_props=[0.5, 0.5], clock_rate=0.8)
    want = brute_force_pattern_logl(tree, part)
    for scaling in (0, 1, 2):
        ll, pat = oracle.tree_likelihood_k(part, tree, use_scaling=scaling, want_patterns=True)
        assert np.allclose(pat, want, rtol=1e-12, atol=1e-14)
        assert ll == pytest.approx(float((want * part.weights).sum()), rel=1e-12)
    assert want[8] == pytest.approx(0.0, abs=1e-14)


def test_ascertainment_correction():
    """Lewis's Mkv: condition on the characters being variable.  The excluded (dummy constant) patterns carry weight 0; their
    probabilities are removed from the sample space (Alignment.getAscertainmentCorrection, TreeLikelihood.calcLogP)."""
    K = 2
    tree = five_tip_tree()
    rng = np.random.default_rng(5)
    real = rng.integers(0, K, size=(5, 6)).astype(np.uint8)
    dummies = np.stack([np.full(5, s, np.uint8) for s in range(K)], axis=1)
    states = np.concatenate([real, dummies], axis=1)
    w = np.concatenate([rng.integers(1, 3, size=6), np.zeros(K, int)])
    base = dict(tip_states=states, weights=w, n_states=K, subst_params=lewis_mk_eigen(K), clock_rate=1.3)
    plain = MorphPartition(**base)
    asc = MorphPartition(excluded=[6, 7], **base)
    _, pat = oracle.tree_likelihood_k(plain, tree, want_patterns=True)
    corr = math.log(1.0 - math.exp(pat[6]) - math.exp(pat[7]))
    want = float(((pat - corr) * w).sum())
    assert oracle.tree_likelihood_k(asc, tree) == pytest.approx(want, rel=1e-13)
    assert oracle.tree_likelihood_k(asc, tree) > oracle.tree_likelihood_k(plain, tree)   # conditioning can only raise it


def test_four_states_reproduce_the_nucleotide_oracle_bitwise():
    pb = synth.make_problem("C3", n_loci=2, n_sites=90)
    loc = pb.loci[0]
    sp = loc.subst_params
    part = MorphPartition(tip_states=loc.tip_states, weights=loc.weights, n_states=4,
                          subst_params=eigen_params_k(sp[0:4], sp[4:20], sp[20:36], sp[36:40]),
                          cat_rates=loc.cat_rates, cat_props=loc.cat_props, clock_rate=loc.clock_rate)
    br = np.full(loc.n_nodes, loc.clock_rate)
    for scaling in (0, 1):
        a, d = oracle.tree_likelihood(loc, br, use_scaling=scaling, want_details=True)
        b, pat = oracle.tree_likelihood_k(part, loc.tree, br, use_scaling=scaling, want_patterns=True)
        assert a == b and np.array_equal(d["pattern_logl"], pat)


def test_scaling_rescues_a_deep_binary_character_tree():
    rng = np.random.default_rng(8)
    sp = synth.yule_species_tree(rng, 400)
    states = rng.integers(0, 2, size=(sp.n_leaves, 40)).astype(np.uint8)
    part = MorphPartition(tip_states=states, weights=np.ones(40, int), n_states=2, subst_params=lewis_mk_eigen(2), clock_rate=200.0)
    unscaled = oracle.tree_likelihood_k(part, sp, use_scaling=0)
    scaled = oracle.tree_likelihood_k(part, sp, use_scaling=1)
    lazy = oracle.tree_likelihood_k(part, sp, use_scaling=2)
    assert math.isfinite(scaled) and scaled == lazy
    assert unscaled == -math.inf or unscaled == pytest.approx(scaled, rel=1e-10)


def test_posterior_adds_the_morphological_partitions():
    pb = synth.make_problem("C4", n_loci=3, n_sites=80, n_morph=3)
    assert len(pb.morph) == 3 and {m.n_states for m in pb.morph} <= {2, 3, 4, 5}
    r = oracle.posterior(pb)
    assert len(r.per_locus_lik) == 3 + 3 and np.all(r.per_locus_coal[3:] == 0.0)
    pb2 = synth.make_problem("C4", n_loci=3, n_sites=80, n_morph=0)
    r2 = oracle.posterior(pb2)
    extra = sum(oracle.tree_likelihood_k(m, pb.species) for m in pb.morph)
    assert r.likelihood == pytest.approx(r2.likelihood + extra, rel=1e-13)
    assert r.coalescent == r2.coalescent


def test_morphological_partitions_travel_with_exactly_one_shard():
    """Sharded runs (bench.py under torchrun, Engine.from_problem on a locus range): every partition of the C4 config belongs to
    exactly one shard, in partition order, whatever the number of shards -- and the shards' oracle log-likelihoods add up to the
    whole problem's."""
    from starbeast2_b200 import synth
    full = synth.make_problem("C4", n_loci=24, n_sites=40)
    assert len(full.morph) == 4
    want = oracle.posterior(full)
    for world in (1, 2, 3, 8, 24):
        cuts = [(24 * r) // world for r in range(world + 1)]
        shards = [synth.make_problem("C4", n_loci=24, n_sites=40, locus_range=(a, b)) for a, b in zip(cuts[:-1], cuts[1:])]
        got = [m for sh in shards for m in sh.morph]
        assert len(got) == 4 and all(np.array_equal(a.tip_states, b.tip_states) and a.n_states == b.n_states for a, b in zip(full.morph, got))
        owners = [r for r, sh in enumerate(shards) for _ in sh.morph]
        assert owners == sorted(owners)
        if world >= 4:
            assert max(owners.count(r) for r in set(owners)) == 1          # spread: no shard carries two of them
        lik = sum(float(np.sum(oracle.posterior(sh).per_locus_lik)) for sh in shards)
        assert lik == pytest.approx(want.likelihood, rel=1e-12)
