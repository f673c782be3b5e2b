"""Error behaviour and edge shapes of the C ABI (include/sb2gpu.h): configuration errors come back as negative codes with a
message (the Java twins turn them into IllegalArgumentException / IllegalStateException at initAndValidate), never as a crash;
numerical failure is -infinity, never an error."""
import math

import numpy as np
import pytest

import oracle
from starbeast2_b200 import synth
from starbeast2_b200._lib import Sb2Error
from starbeast2_b200.engine import Engine
from starbeast2_b200.problem import CLOCK_EXPLICIT, POP_CONSTANT, Locus, Problem, TreeArrays, hky_params

pytestmark = pytest.mark.gpu


def test_argument_and_state_errors():
    pb = synth.make_problem("C2", n_loci=2, n_sites=40)
    e = Engine(device=0)
    loc = pb.loci[0]
    with pytest.raises(Sb2Error):
        e.finalize(pb.species.n_nodes, pb.species.n_leaves)            # no loci
    with pytest.raises(Sb2Error):
        e.add_locus(loc.tip_states, loc.weights, 0, loc.tip_species)   # zero categories
    e.add_locus(loc.tip_states, loc.weights, 1, loc.tip_species, loc.ploidy)
    with pytest.raises(Sb2Error):
        e.add_locus(loc.tip_states, loc.weights, 4, loc.tip_species)   # category count differs from the first locus
    with pytest.raises(Sb2Error):
        e.set_species_tree(pb.species)                                 # not finalized yet
    bad = loc.tip_species.copy()
    bad[0] = pb.species.n_leaves                                       # an internal species node is not a valid tip mapping
    e2 = Engine(device=0)
    e2.add_locus(loc.tip_states, loc.weights, 1, bad)
    with pytest.raises(Sb2Error):
        e2.finalize(pb.species.n_nodes, pb.species.n_leaves)
    e2.close()
    e.finalize(pb.species.n_nodes, pb.species.n_leaves)
    with pytest.raises(Sb2Error):
        e.add_locus(loc.tip_states, loc.weights, 1, loc.tip_species)   # after finalize
    with pytest.raises(Sb2Error):
        e.set_gene_tree(5, loc.tree)                                   # locus out of range
    t = loc.tree.copy()
    t.parent[t.root] = 0                                               # no root
    with pytest.raises(Sb2Error):
        e.set_gene_tree(0, t)
    t = loc.tree.copy()
    t.left[loc.n_tips] = t.n_nodes + 3                                 # child index out of range
    with pytest.raises(Sb2Error):
        e.set_gene_tree(0, t)
    sp = pb.species.copy()
    sp.parent[0] = -1                                                  # two roots
    with pytest.raises(Sb2Error):
        e.set_species_tree(sp)
    with pytest.raises(Sb2Error):
        e.partials(0, 0)                                               # a tip has no partials buffer
    e.close()


def test_smallest_shapes_one_pattern_two_tips_one_species():
    """Degenerate but legal inputs: 2 tips in 1 species, one pattern; a 3-tip tree with an all-missing alignment."""
    sp = TreeArrays([-1], [-1], [-1], [0.0])
    g = TreeArrays([2, 2, -1], [-1, -1, 0], [-1, -1, 1], [0.0, 0.0, 0.7])
    loc = Locus(tree=g, tip_states=np.array([[1], [3]], np.uint8), weights=np.array([5], np.int32), tip_species=np.zeros(2, np.int32),
                subst_params=hky_params(2.0, [0.3, 0.2, 0.2, 0.3]), clock_rate=0.4)
    pb = Problem(species=sp, loci=[loc], pop_kind=POP_CONSTANT, pop_a=np.array([0.2]), species_rates=np.ones(1))
    e = Engine.from_problem(pb)
    got, want = e.eval_posterior(), oracle.posterior(pb)
    assert got.total == pytest.approx(want.total, rel=1e-12)
    e.close()
    g3 = TreeArrays([3, 3, 4, 4, -1], [-1, -1, -1, 0, 3], [-1, -1, -1, 1, 2], [0.0, 0.0, 0.0, 0.2, 0.5])
    loc3 = Locus(tree=g3, tip_states=np.full((3, 4), 4, np.uint8), weights=np.ones(4, np.int32), tip_species=np.zeros(3, np.int32))
    pb3 = Problem(species=sp, loci=[loc3], pop_kind=POP_CONSTANT, pop_a=np.array([0.2]), species_rates=np.ones(1))
    e = Engine.from_problem(pb3)
    got = e.eval_posterior()
    assert got.likelihood == 0.0                                       # every pattern is all-missing: likelihood exactly 1
    assert got.coalescent == pytest.approx(oracle.posterior(pb3).coalescent, rel=1e-12)
    e.close()


def test_numerical_failure_is_minus_infinity_or_nan_never_an_error():
    pb = synth.make_problem("C2", n_loci=3, n_sites=60)
    e = Engine.from_problem(pb)
    base = e.eval_posterior()
    assert math.isfinite(base.total)
    # a zero population size: -k log(0) - gamma/0 -> the reference returns +-inf / NaN from calculateLogP, it does not throw
    bad = pb.pop_a.copy()
    bad[0] = 0.0
    e.store()
    e.set_pop_model(POP_CONSTANT, bad)
    r = e.eval_posterior()
    pb.pop_a = bad
    want = oracle.posterior(pb)
    assert (math.isnan(r.coalescent) and math.isnan(want.coalescent)) or r.coalescent == want.coalescent
    e.restore()
    assert e.eval_posterior().total == base.total
    # zero branch rates: every transition matrix is the identity; conflicting tips give likelihood 0 -> log = -inf
    e.store()
    for l, loc in enumerate(pb.loci):
        e.set_clock(l, CLOCK_EXPLICIT, 1.0, np.zeros(loc.tree.n_nodes))
    r = e.eval_posterior()
    assert r.likelihood == -math.inf
    e.restore()
    assert e.eval_posterior().total == base.total
    e.close()


def test_malformed_trees_are_rejected_not_hung():
    """A parent array with a cycle detached from the root, or child / parent arrays that disagree, would spin k_prepare's
    rootward walks forever; the host-side check must refuse them."""
    pb = synth.make_problem("C2", n_loci=1, n_sites=40)
    e = Engine.from_problem(pb)
    loc = pb.loci[0]
    nT = loc.n_tips
    t = loc.tree.copy()
    # two internal non-root nodes that are each other's parent: indices stay in range and there is still exactly one root
    a = next(i for i in range(nT, t.n_nodes) if t.parent[i] >= 0 and t.parent[t.parent[i]] >= 0)
    b = int(t.parent[a])
    t.parent[b] = a
    with pytest.raises(Sb2Error, match="disagree|connected"):
        e.set_gene_tree(0, t)
    t = loc.tree.copy()
    t.right[t.root] = t.left[t.root]                                   # the same child twice
    with pytest.raises(Sb2Error):
        e.set_gene_tree(0, t)
    sp = pb.species.copy()
    sp.right[sp.n_nodes - 1] = -1                                      # left >= 0 but right < 0
    with pytest.raises(Sb2Error):
        e.set_species_tree(sp)
    sp = pb.species.copy()
    i = next(i for i in range(sp.n_leaves, sp.n_nodes) if sp.parent[i] >= 0 and sp.parent[sp.parent[i]] >= 0)
    sp.parent[int(sp.parent[i])] = i
    with pytest.raises(Sb2Error):
        e.set_species_tree(sp)
    # nothing was staged by the refused calls: the engine still evaluates the original state
    assert e.eval_posterior().total == pytest.approx(oracle.posterior(pb).total, rel=1e-9)
    e.close()


def test_restore_without_store_and_store_with_staged_inputs():
    pb = synth.make_problem("C2", n_loci=3, n_sites=60)
    e = Engine.from_problem(pb)
    e.eval_posterior()
    with pytest.raises(Sb2Error, match="without a prior"):
        e.restore()
    # stage a new gene tree, store WITHOUT evaluating, then propose something else and reject it:
    # the stored state is the staged one (the staged values must not be lost by the restore)
    rng = np.random.default_rng(5)
    t1, _ = synth.perturb_gene_tree(rng, pb.loci[1].tree, pb.loci[1].n_tips)
    e.set_gene_tree(1, t1)
    e.store()
    t2, _ = synth.perturb_gene_tree(rng, pb.loci[2].tree, pb.loci[2].n_tips)
    e.set_gene_tree(2, t2)
    e.eval_posterior()
    e.restore()
    pb.loci[1].tree = t1
    got, want = e.eval_posterior(), oracle.posterior(pb)
    assert got.total == pytest.approx(want.total, rel=1e-9)
    np.testing.assert_allclose(got.per_locus_lik, want.per_locus_lik, rtol=1e-9)
    e.close()


def test_strict_protocol_two_coalescent_evaluations_without_a_likelihood(monkeypatch):
    """SB2_NO_SPECULATION: eval_coalescent runs k_prepare but no pruning kernel.  A second staged input + evaluation must
    first run the pending walk, or the nodes re-scheduled by the first k_prepare would be treated as clean."""
    monkeypatch.setenv("SB2_NO_SPECULATION", "1")
    pb = synth.make_problem("C2", n_loci=3, n_sites=60)
    e = Engine.from_problem(pb)
    e.eval_posterior()
    rng = np.random.default_rng(9)
    t0, _ = synth.perturb_gene_tree(rng, pb.loci[0].tree, pb.loci[0].n_tips)
    e.set_gene_tree(0, t0)
    e.eval_coalescent()
    t1, _ = synth.perturb_gene_tree(rng, pb.loci[1].tree, pb.loci[1].n_tips)
    e.set_gene_tree(1, t1)
    e.eval_coalescent()
    pb.loci[0].tree, pb.loci[1].tree = t0, t1
    got, want = e.eval_likelihood(), oracle.posterior(pb)
    np.testing.assert_allclose(got[1], want.per_locus_lik, rtol=1e-9)
    e.close()


def test_large_shared_memory_shapes_on_every_device():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per device: an engine on EVERY visible GPU must be able to launch the
    > 48 KB instantiations (many tips: big k_prepare / tip-state tiles), driven by one host thread."""
    import torch
    pb = synth.make_problem("C2", n_loci=2, n_sites=60, n_species=80, tips_per=4)   # 320 tips: k_prepare > 48 KB
    want = oracle.posterior(pb)
    engines = [Engine.from_problem(pb, device=d) for d in range(torch.cuda.device_count())]
    for e in engines:
        assert e.eval_posterior().total == pytest.approx(want.total, rel=1e-9)
    for e in engines:
        e.close()
