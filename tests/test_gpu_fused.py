"""The fused one-launch step (csrc/step.cu: one thread-block cluster per locus) against the three-kernel path and the oracle.

Both paths share the engine's whole state and the pruning arithmetic (walk_common.cuh), and per-locus log-likelihoods are
sums of 32-pattern chunk sums in pattern order whatever the launch geometry -- so they must agree BIT FOR BIT, not just to
1e-9: SB2_FUSED=0 / 1 pins the path per engine, SB2_FUSED_MAX_LANES=0 makes one engine alternate (single-locus steps through
k_step, all-loci evaluations through the three kernels)."""
import math
import os

import numpy as np
import pytest

import oracle
from starbeast2_b200 import synth
from starbeast2_b200.engine import Engine
from starbeast2_b200.problem import POP_ANALYTIC

pytestmark = pytest.mark.gpu


class env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def engine(pb, fused, **kw):
    with env(SB2_FUSED=fused, **kw):
        return Engine.from_problem(pb)


def same_bits(a, b):
    for f in ("per_locus_lik", "per_locus_coal", "per_branch"):
        x, y = getattr(a, f), getattr(b, f)
        assert np.array_equal(x.view(np.int64), y.view(np.int64)), (f, np.max(np.abs(x - y)))
    assert (a.coalescent, a.likelihood, a.total) == (b.coalescent, b.likelihood, b.total)


def close_to_oracle(got, pb, rel=1e-9):
    want = oracle.posterior(pb, n_threads=2)
    fin = np.isfinite(want.per_locus_coal)   # the oracle does not evaluate the likelihood of an incompatible locus (NaN)
    assert np.allclose(got.per_locus_lik[fin], want.per_locus_lik[fin], rtol=rel, atol=0)
    assert np.array_equal(np.isfinite(got.per_locus_coal), fin)
    assert np.allclose(got.per_locus_coal[fin], want.per_locus_coal[fin], rtol=rel, atol=0)
    if fin.all():
        assert abs(got.total - want.total) <= rel * abs(want.total)


CASES = [
    ("C2", dict(n_loci=7, n_sites=700)),                                   # HKY, strict clock, 3 CTAs per cluster
    ("C3", dict(n_loci=5, n_sites=300)),                                   # GTR+G4, species-tree clock (followers embed too), LWCR
    ("C3", dict(n_loci=4, n_sites=40)),                                    # single-CTA clusters
    ("C4", dict(n_loci=6, n_sites=200, n_morph=0)),                        # fossil species tree (gene loci only: K-state partitions keep the k_prepare + k_kstate path)
    ("C5", dict(n_loci=3, n_sites=260)),                                   # 120 tips, 4 categories: the large-schedule shared-memory carve
    ("C2", dict(n_loci=4, n_sites=500, missing_frac=0.2)),
    ("C2", dict(n_loci=3, n_sites=300, n_cat=3)),                          # a category count that leaves warps idle
    ("C2", dict(n_loci=3, n_sites=300, pop="analytic")),                   # analytic population sizes: closed forms in the leader
]


@pytest.mark.parametrize("name,kw", CASES)
def test_full_evaluation_fused_equals_three_kernels_bitwise(name, kw):
    pb = synth.make_problem(name, **kw)
    a, b = engine(pb, 0), engine(pb, 1)
    ra, rb = a.eval_posterior(), b.eval_posterior()
    assert b.launch_count() == 1 and a.launch_count() == 3
    same_bits(ra, rb)
    close_to_oracle(rb, pb)
    assert a.last_node_updates() == b.last_node_updates() == sum(l.n_tips - 1 for l in pb.loci)
    # nothing staged: the cached result is handed out, no launch
    rb2 = b.eval_posterior()
    same_bits(rb, rb2)
    assert b.launch_count() == 1


@pytest.mark.parametrize("name,kw", [CASES[0], CASES[1], CASES[4], CASES[7]])
def test_mcmc_steps_fused_equal_three_kernels_bitwise(name, kw):
    """store / move one gene-tree node / two-phase evaluation / accept or restore, the same moves on both paths."""
    pb = synth.make_problem(name, **kw)
    a, b = engine(pb, 0), engine(pb, 1)
    same_bits(a.eval_posterior(), b.eval_posterior())
    rng = np.random.default_rng(5)
    for step in range(16):
        l = int(rng.integers(0, pb.n_loci))
        loc = pb.loci[l]
        new_tree, _ = synth.perturb_gene_tree(rng, loc.tree, loc.n_tips)
        old_tree = loc.tree
        for e in (a, b):
            e.store()
            e.set_gene_tree(l, new_tree)
        loc.tree = new_tree
        lb = b.launch_count()
        ca, cb = a.eval_coalescent(), b.eval_coalescent()
        assert ca[0] == cb[0] and np.array_equal(ca[1].view(np.int64), cb[1].view(np.int64))
        if math.isfinite(ca[0]):
            la_, lb_ = a.eval_likelihood(), b.eval_likelihood()
            assert la_[0] == lb_[0] and np.array_equal(la_[1].view(np.int64), lb_[1].view(np.int64))
        assert b.launch_count() - lb == 1                      # ONE launch per MCMC step
        assert a.last_node_updates() == b.last_node_updates()
        ra, rb = a.eval_posterior(), b.eval_posterior()
        same_bits(ra, rb)
        close_to_oracle(rb, pb)
        if step % 3 == 0 and math.isfinite(ca[0]):
            a.accept(); b.accept()
        else:
            a.restore(); b.restore()
            loc.tree = old_tree
            same_bits(a.eval_posterior(), b.eval_posterior())
    close_to_oracle(b.eval_posterior(), pb)


def test_one_engine_alternating_between_the_paths():
    """auto mode with SB2_FUSED_MAX_LANES=0: all-loci evaluations take the three kernels, single-locus steps take k_step --
    on ONE state; a second engine that never fuses is the bitwise witness."""
    pb = synth.make_problem("C3", n_loci=5, n_sites=200)
    a = engine(pb, 0)
    b = engine(pb, None, SB2_FUSED_MAX_LANES=0)
    same_bits(a.eval_posterior(), b.eval_posterior())
    assert b.launch_count() == 3
    rng = np.random.default_rng(9)
    for step in range(12):
        l = int(rng.integers(0, pb.n_loci))
        loc = pb.loci[l]
        new_tree, _ = synth.perturb_gene_tree(rng, loc.tree, loc.n_tips)
        for e in (a, b):
            e.store()
            e.set_gene_tree(l, new_tree)
        n0 = b.launch_count()
        same_bits(a.eval_posterior(), b.eval_posterior())
        assert b.launch_count() - n0 == 1
        if step % 2:
            a.restore(); b.restore()
        else:
            a.accept(); b.accept()
            loc.tree = new_tree
        if step % 4 == 3:   # a species-tree-wide change: every locus re-embeds (three kernels in b)
            rates = pb.species_rates * (1.0 + 0.01 * (step + 1))
            rates[-1] = pb.species_rates[-1]
            for e in (a, b):
                e.store()
                e.set_species_rates(rates)
            pb.species_rates = rates
            n0 = b.launch_count()
            same_bits(a.eval_posterior(), b.eval_posterior())
            assert b.launch_count() - n0 == 3
            a.accept(); b.accept()
    close_to_oracle(b.eval_posterior(), pb)


def test_model_moves_and_two_loci_in_one_step():
    """A substitution-model move dirties every matrix of one locus (full schedule through the single-locus cluster); two loci
    staged in one step go through the all-loci form."""
    pb = synth.make_problem("C2", n_loci=5, n_sites=400)
    a, b = engine(pb, 0), engine(pb, 1)
    same_bits(a.eval_posterior(), b.eval_posterior())
    from starbeast2_b200.problem import hky_params
    p = hky_params(3.5, np.array([0.27, 0.23, 0.19, 0.31]))
    for e in (a, b):
        e.store()
        e.set_subst_model(2, pb.loci[2].subst_kind, p)
    pb.loci[2].subst_params = p
    same_bits(a.eval_posterior(), b.eval_posterior())
    assert b.last_node_updates() == pb.loci[2].n_tips - 1
    close_to_oracle(b.eval_posterior(), pb)
    rng = np.random.default_rng(3)
    for e in (a, b):
        e.accept()
        e.store()
    for l in (0, 4):
        t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
        pb.loci[l].tree = t
        a.set_gene_tree(l, t); b.set_gene_tree(l, t)
    same_bits(a.eval_posterior(), b.eval_posterior())
    close_to_oracle(b.eval_posterior(), pb)
    a.restore(); b.restore()
    same_bits(a.eval_posterior(), b.eval_posterior())


def test_shapes_beyond_the_cluster_fall_back_to_three_kernels():
    """More patterns than 16 CTAs x 64 (four categories) hold: the engine keeps the three-kernel path, silently and correctly."""
    pb = synth.make_problem("C2", n_loci=2, n_sites=12000, n_species=8, tips_per=2, n_cat=4)
    if max(l.n_pat for l in pb.loci) <= 16 * 64:
        pytest.skip("synthetic alignment compressed below the cluster limit")
    e = engine(pb, 1)
    r = e.eval_posterior()
    assert e.launch_count() == 3
    close_to_oracle(r, pb)


@pytest.mark.parametrize("fused", [0, 1])
def test_analytic_closed_form_memo_never_changes_a_bit(fused):
    """MultispeciesCoalescent.java:165-195 keeps a species branch's term while nothing it depends on changes; here the closed
    form is memoised per branch by VALUE (two entries).  An engine with the memo and one without (SB2_NO_MEMO) must agree bit
    for bit through proposals, rejections (the restored key is the other entry), acceptances and hyper-parameter moves."""
    pb = synth.make_problem("C4", n_loci=12, n_sites=120, n_morph=0)
    assert pb.pop_kind == POP_ANALYTIC
    a = engine(pb, fused, SB2_NO_MEMO=None)
    b = engine(pb, fused, SB2_NO_MEMO=1)
    same_bits(a.eval_posterior(), b.eval_posterior())
    rng = np.random.default_rng(11)
    alpha, beta = pb.alpha, pb.beta
    for step in range(24):
        l = int(rng.integers(0, pb.n_loci))
        loc = pb.loci[l]
        new_tree, _ = synth.perturb_gene_tree(rng, loc.tree, loc.n_tips)
        for e in (a, b):
            e.store()
            e.set_gene_tree(l, new_tree)
        if step % 5 == 4:   # populationShape / populationMean move on top
            alpha, beta = alpha * float(rng.uniform(0.8, 1.25)), beta * float(rng.uniform(0.8, 1.25))
            for e in (a, b):
                e.set_pop_model(pb.pop_kind, None, None, alpha, beta)
        ra, rb = a.eval_posterior(), b.eval_posterior()
        same_bits(ra, rb)
        if step % 3 == 0:   # accept
            loc.tree = new_tree
            pb.alpha, pb.beta = alpha, beta
            for e in (a, b):
                e.accept()
        else:               # reject: the state -- and the hyper-parameters -- go back
            alpha, beta = pb.alpha, pb.beta
            for e in (a, b):
                e.restore()
                e.set_pop_model(pb.pop_kind, None, None, alpha, beta)
            same_bits(a.eval_posterior(), b.eval_posterior())
    close_to_oracle(a.eval_posterior(), pb)
    a.close(); b.close()
