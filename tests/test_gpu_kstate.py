"""K-state likelihood-only loci (csrc/kstate.cu) through the C ABI against the oracle: the morphological TreeLikelihoods of the
FBD configuration (tutorial/CanisFBD/Canis-FBD.xml:964-1033) -- K = 2..5 characters on the species tree with its fossil tips,
LewisMK, ascertained -- next to the gene loci of a C4-style problem."""
import math

import numpy as np
import pytest

import oracle
from starbeast2_b200 import synth
from starbeast2_b200._lib import Sb2Error
from starbeast2_b200.engine import Engine
from starbeast2_b200.problem import MorphPartition, TreeArrays, lewis_mk_eigen

pytestmark = pytest.mark.gpu


def check(got, want, rel=1e-9):
    assert np.allclose(got.per_locus_lik, want.per_locus_lik, rtol=rel, atol=0), (got.per_locus_lik, want.per_locus_lik)
    assert np.allclose(got.per_locus_coal, want.per_locus_coal, rtol=rel, atol=0)
    assert abs(got.total - want.total) <= rel * abs(want.total)
    assert abs(got.likelihood - want.likelihood) <= rel * abs(want.likelihood)


def test_c4_with_morphology_matches_the_oracle():
    pb = synth.make_problem("C4", n_loci=6, n_sites=200, n_morph=4)
    assert sorted(m.n_states for m in pb.morph) == [2, 3, 4, 5] and all(len(m.excluded) == m.n_states for m in pb.morph)
    e = Engine.from_problem(pb)
    got = e.eval_posterior()
    want = oracle.posterior(pb, n_threads=2)
    assert len(got.per_locus_lik) == 6 + 4 and np.all(got.per_locus_coal[6:] == 0.0)
    check(got, want)
    # per-pattern values, and the ascertainment correction on top of them
    for j, m in enumerate(pb.morph):
        ll, pat = oracle.tree_likelihood_k(m, pb.species, want_patterns=True)
        assert np.allclose(e.pattern_logl(6 + j), pat, rtol=1e-11, atol=1e-13)
        assert got.per_locus_lik[6 + j] == pytest.approx(ll, rel=1e-11)


@pytest.mark.parametrize("fused", ["0", "1", None])
def test_morphology_through_mcmc_steps(fused, monkeypatch):
    """Gene-tree moves leave the morphological likelihoods cached (one-launch steps stay one launch); species-tree moves
    re-stage the tree of every partition; store / restore / accept keep the cached values straight."""
    if fused is None:
        monkeypatch.delenv("SB2_FUSED", raising=False)
    else:
        monkeypatch.setenv("SB2_FUSED", fused)
    pb = synth.make_problem("C4", n_loci=5, n_sites=150, n_morph=3)
    e = Engine.from_problem(pb)
    L, M = 5, 3
    check(e.eval_posterior(), oracle.posterior(pb))
    rng = np.random.default_rng(4)
    for step in range(10):
        e.store()
        if step % 3 == 2:
            # a species-tree height move: every gene tree re-embeds, every morphological partition sees the new tree
            sp = pb.species
            old = sp.copy()
            cand = [i for i in range(sp.n_leaves, sp.n_nodes) if sp.parent[i] >= 0]
            i = int(rng.choice(cand))
            lo = max(sp.height[sp.left[i]], sp.height[sp.right[i]])
            hi = sp.height[sp.parent[i]]
            sp.height[i] = lo + (hi - lo) * rng.uniform(0.3, 0.7)
            e.set_species_tree(sp)
            for j in range(M):
                e.set_gene_tree(L + j, sp)
            got, want = e.eval_posterior(), oracle.posterior(pb)
            fin = np.isfinite(want.per_locus_coal)
            assert np.array_equal(np.isfinite(got.per_locus_coal), fin)
            assert np.allclose(got.per_locus_lik[L:], want.per_locus_lik[L:], rtol=1e-9, atol=0)
            if fin.all():
                check(got, want)
                e.accept()
            else:
                e.restore()
                pb.species = old
                e.set_species_tree(old)   # the host mirror goes back with BEAST's own restore
        else:
            l = int(rng.integers(0, L))
            t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
            old = pb.loci[l].tree
            pb.loci[l].tree = t
            e.set_gene_tree(l, t)
            n0 = e.launch_count()
            got, want = e.eval_posterior(), oracle.posterior(pb)
            if fused != "0":
                assert e.launch_count() - n0 == 1
            if np.isfinite(want.per_locus_coal).all():
                check(got, want)
            if step % 2:
                e.restore()
                pb.loci[l].tree = old
            else:
                e.accept()
        back = e.eval_posterior()
        want = oracle.posterior(pb)
        if np.isfinite(want.per_locus_coal).all():
            check(back, want)


def test_a_morphology_only_engine_and_a_clock_rate_move():
    rng = np.random.default_rng(2)
    sp = synth.yule_species_tree(rng, 9, 5)
    parts = []
    for K in (2, 5):
        pats, w, excl = synth.simulate_morphology(rng, sp, K, 30, 2.0)
        parts.append(MorphPartition(tip_states=pats, weights=w, n_states=K, subst_params=lewis_mk_eigen(K), clock_rate=1.5, excluded=excl,
                                    cat_rates=[0.2, 0.8, 2.0], cat_props=[0.3, 0.4, 0.3]))
    e = Engine()
    for m in parts:
        e.add_locus_k(m.tip_states, m.weights, m.n_states, m.n_cat, m.excluded)
    e.finalize(sp.n_nodes, sp.n_leaves)
    e.set_species_tree(sp)
    e.set_pop_model(0, np.ones(sp.n_nodes))
    for i, m in enumerate(parts):
        e.set_gene_tree(i, sp)
        e.set_subst_model(i, 1, m.subst_params)
        e.set_site_model(i, m.cat_rates, m.cat_props, 0.0)
        e.set_clock(i, 0, m.clock_rate, None)
    r = e.eval_posterior()
    want = [oracle.tree_likelihood_k(m, sp) for m in parts]
    assert np.allclose(r.per_locus_lik, want, rtol=1e-11, atol=0) and r.coalescent == 0.0
    e.store()
    parts[1].clock_rate = 0.9
    e.set_clock(1, 0, 0.9, None)
    r2 = e.eval_posterior()
    assert r2.per_locus_lik[0] == r.per_locus_lik[0]
    assert r2.per_locus_lik[1] == pytest.approx(oracle.tree_likelihood_k(parts[1], sp), rel=1e-11)
    e.restore()
    r3 = e.eval_posterior()
    assert np.array_equal(r3.per_locus_lik, r.per_locus_lik)


def test_k_state_argument_errors():
    e = Engine()
    st = np.zeros((4, 3), np.uint8)
    w = np.ones(3, np.int32)
    with pytest.raises(Sb2Error, match="2\\.\\."):
        e.add_locus_k(st, w, 7)
    with pytest.raises(Sb2Error, match="out of range"):
        e.add_locus_k(st, w, 3, excluded=[5])
    i = e.add_locus_k(st, w, 3)
    e.finalize(7, 4)
    with pytest.raises(Sb2Error, match="eigen"):
        e.set_subst_model(i, 0, np.ones(5))
    with pytest.raises(Sb2Error, match="proportionInvariant"):
        e.set_site_model(i, [1.0], [1.0], 0.1)
    with pytest.raises(Sb2Error, match="StarBeastClock"):
        e.set_clock(i, 1, 1.0, None)


def test_a_partition_too_large_for_shared_memory_uses_the_hbm_scratch():
    """k_kstate keeps a locus's partials in shared memory up to 96 KB; a larger partition (here ~1 MB) takes the HBM scratch,
    next to a small one in the same launch."""
    rng = np.random.default_rng(6)
    sp = synth.yule_species_tree(rng, 40)
    parts = []
    for K, n_chars in ((5, 400), (3, 12)):
        pats, w, excl = synth.simulate_morphology(rng, sp, K, n_chars, 1.5)
        parts.append(MorphPartition(tip_states=pats, weights=w, n_states=K, subst_params=lewis_mk_eigen(K), clock_rate=1.2, excluded=excl,
                                    cat_rates=[0.4, 1.6], cat_props=[0.5, 0.5]))
    assert parts[0].n_pat * 39 * 2 * 44 > 96 * 1024 > parts[1].n_pat * 39 * 2 * 28
    e = Engine()
    for m in parts:
        e.add_locus_k(m.tip_states, m.weights, m.n_states, m.n_cat, m.excluded)
    e.finalize(sp.n_nodes, sp.n_leaves)
    e.set_species_tree(sp)
    e.set_pop_model(0, np.ones(sp.n_nodes))
    for i, m in enumerate(parts):
        e.set_gene_tree(i, sp)
        e.set_subst_model(i, 1, m.subst_params)
        e.set_site_model(i, m.cat_rates, m.cat_props, 0.0)
        e.set_clock(i, 0, m.clock_rate, None)
    r = e.eval_posterior()
    want = [oracle.tree_likelihood_k(m, sp) for m in parts]
    assert np.allclose(r.per_locus_lik, want, rtol=1e-11, atol=0)
