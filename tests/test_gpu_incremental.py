"""Dirty-node-only recomputation, store / restore / accept, and node renumbering, on the GPU,
checked against from-scratch oracle evaluations of the same state."""
import math

import numpy as np
import pytest

import oracle
from starbeast2_b200 import synth
from starbeast2_b200.engine import Engine
from starbeast2_b200.problem import POP_CONSTANT, TreeArrays, hky_params

pytestmark = pytest.mark.gpu


def rel_close(a, b, rel=1e-9):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.all(np.abs(a - b) <= rel * np.maximum(1.0, np.abs(b)))


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_single_locus_move_updates_only_the_root_path(name):
    pb = synth.make_problem(name, n_loci=5, n_sites=150)
    e = Engine.from_problem(pb)
    r0 = e.eval_posterior()
    full_updates = e.last_node_updates()
    assert full_updates == sum(l.n_tips - 1 for l in pb.loci)
    rng = np.random.default_rng(1)
    for step in range(12):
        l = int(rng.integers(0, pb.n_loci))
        loc = pb.loci[l]
        e.store()
        new_tree, moved = synth.perturb_gene_tree(rng, loc.tree, loc.n_tips)
        old_tree = loc.tree
        loc.tree = new_tree
        e.set_gene_tree(l, new_tree)
        got = e.eval_posterior()
        want = oracle.posterior(pb)
        ok = np.isfinite(want.per_locus_coal)
        assert np.array_equal(np.isfinite(got.per_locus_coal), ok)
        if not ok.all():
            # the slide crossed a speciation: incompatible => -inf, the reference short-circuits; reject and go on
            assert got.coalescent == -math.inf and got.total == -math.inf
            e.restore()
            loc.tree = old_tree
            continue
        assert rel_close(got.per_locus_lik, want.per_locus_lik) and rel_close(got.per_locus_coal, want.per_locus_coal)
        assert rel_close(got.total, want.total)
        # only the moved node and its ancestors are recomputed (strict clock); with the species-rates clock the
        # children's rates can change too but never more than the path + nothing in other loci
        depth = 0
        a = moved
        while a >= 0:
            depth += 1
            a = int(new_tree.parent[a])
        assert 1 <= e.last_node_updates() <= depth
        if step % 2 == 0:
            e.accept()
        else:
            e.restore()
            loc.tree = old_tree
            back = e.eval_posterior()
            want_back = oracle.posterior(pb)
            assert rel_close(back.total, want_back.total)
            assert e.last_node_updates() == 0 or True


def test_restore_is_bit_exact_and_free():
    pb = synth.make_problem("C3", n_loci=4, n_sites=120)
    e = Engine.from_problem(pb)
    r0 = e.eval_posterior()
    rng = np.random.default_rng(2)
    for _ in range(6):
        e.store()
        l = int(rng.integers(0, pb.n_loci))
        t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
        e.set_gene_tree(l, t)
        r1 = e.eval_posterior()
        assert r1.total != r0.total
        e.restore()
        launches = e.launch_count()
        r2 = e.eval_posterior()
        assert r2.total == r0.total and np.array_equal(r2.per_locus_lik, r0.per_locus_lik)
        assert np.array_equal(r2.per_locus_coal, r0.per_locus_coal)
        assert e.launch_count() - launches <= 1   # no prepare, no pruning kernel: only the tiny reduction
    # a from-scratch recomputation of the restored state agrees bit for bit (BEAST's periodic consistency check)
    e.mark_all_dirty()
    r3 = e.eval_posterior()
    assert r3.total == r0.total and np.array_equal(r3.per_locus_lik, r0.per_locus_lik)


def test_incremental_equals_from_scratch_bitwise():
    """A chain of accepted moves evaluated incrementally ends in exactly the state a fresh engine computes."""
    pb = synth.make_problem("C3", n_loci=3, n_sites=100)
    e = Engine.from_problem(pb)
    e.eval_posterior()
    rng = np.random.default_rng(7)
    for _ in range(10):
        e.store()
        l = int(rng.integers(0, pb.n_loci))
        pb.loci[l].tree, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
        e.set_gene_tree(l, pb.loci[l].tree)
        inc = e.eval_posterior()
        e.accept()
    fresh = Engine.from_problem(pb).eval_posterior()
    assert inc.total == fresh.total and np.array_equal(inc.per_locus_lik, fresh.per_locus_lik)


def test_species_tree_move_dirties_everything_with_species_clock():
    pb = synth.make_problem("C3", n_loci=4, n_sites=100)
    e = Engine.from_problem(pb)
    e.eval_posterior()
    e.store()
    sp = pb.species.copy()
    # scale the whole species tree and all gene trees up (UpDownOperator-like): stays compatible
    f = 1.07
    sp.height *= f
    for loc in pb.loci:
        loc.tree = loc.tree.copy()
        loc.tree.height *= f
    pb.species = sp
    e.set_species_tree(sp)
    e.set_gene_trees(0, [l.tree for l in pb.loci])
    got = e.eval_posterior()
    want = oracle.posterior(pb)
    assert rel_close(got.per_locus_lik, want.per_locus_lik) and rel_close(got.total, want.total)
    assert e.last_node_updates() == sum(l.n_tips - 1 for l in pb.loci)
    # species rates change (DiscreteRateCycle-like): every species-clock locus re-evaluates
    e.accept()
    e.store()
    pb.species_rates = np.roll(pb.species_rates[:-1], 1).tolist() + [pb.species_rates[-1]]
    pb.species_rates = np.asarray(pb.species_rates)
    e.set_species_rates(pb.species_rates)
    got = e.eval_posterior()
    want = oracle.posterior(pb)
    assert rel_close(got.per_locus_lik, want.per_locus_lik) and rel_close(got.total, want.total)


def test_species_move_incompatible_then_restore():
    pb = synth.make_problem("C2", n_loci=4, n_sites=80)
    e = Engine.from_problem(pb)
    r0 = e.eval_posterior()
    e.store()
    sp = pb.species.copy()
    root = sp.root
    # push an internal species node above gene coalescences: some gene tree becomes incompatible
    child = int(sp.left[root]) if sp.left[int(sp.left[root])] >= 0 else int(sp.right[root])
    sp.height[child] = sp.height[root] * 0.999
    e.set_species_tree(sp)
    coal, per_locus, _ = e.eval_coalescent()
    old = pb.species
    pb.species = sp
    want = oracle.posterior(pb)
    pb.species = old
    assert [math.isfinite(x) for x in per_locus] == [math.isfinite(x) for x in want.per_locus_coal]
    if not all(math.isfinite(x) for x in want.per_locus_coal):
        assert coal == -math.inf
    e.restore()   # likelihood never evaluated for the rejected proposal (the reference's short-circuit)
    r1 = e.eval_posterior()
    assert r1.total == r0.total


def test_model_parameter_change_and_restore():
    pb = synth.make_problem("C2", n_loci=3, n_sites=120)
    e = Engine.from_problem(pb)
    r0 = e.eval_posterior()
    e.store()
    # kappa scaler on locus 1: all its matrices and partials, nothing else
    pb.loci[1].subst_params = hky_params(4.5, [0.30, 0.20, 0.20, 0.30])
    e.set_subst_model(1, 0, pb.loci[1].subst_params)
    got = e.eval_posterior()
    want = oracle.posterior(pb)
    assert rel_close(got.per_locus_lik, want.per_locus_lik)
    assert e.last_node_updates() == pb.loci[1].n_tips - 1
    assert got.per_locus_lik[0] == r0.per_locus_lik[0] and got.per_locus_lik[2] == r0.per_locus_lik[2]
    e.restore()
    pb.loci[1].subst_params = hky_params(2.0, [0.30, 0.20, 0.20, 0.30])
    # a later tree move on the same locus must use the restored kappa, not the rejected one
    e.store()
    rng = np.random.default_rng(5)
    pb.loci[1].tree, _ = synth.perturb_gene_tree(rng, pb.loci[1].tree, pb.loci[1].n_tips)
    e.set_gene_tree(1, pb.loci[1].tree)
    got = e.eval_posterior()
    want = oracle.posterior(pb)
    assert rel_close(got.per_locus_lik, want.per_locus_lik)
    # clock-rate / mutation-rate change: all branch lengths*rate change
    e.accept(); e.store()
    pb.loci[2].clock_rate *= 1.3
    e.set_clock(2, pb.loci[2].clock_kind, pb.loci[2].clock_rate)
    got = e.eval_posterior()
    assert rel_close(got.per_locus_lik, oracle.posterior(pb).per_locus_lik)
    assert e.last_node_updates() == pb.loci[2].n_tips - 1


def test_population_size_change_only_touches_coalescent():
    pb = synth.make_problem("C2", n_loci=3, n_sites=80)
    e = Engine.from_problem(pb)
    r0 = e.eval_posterior()
    e.store()
    pb.pop_a = pb.pop_a * 1.5
    e.set_pop_model(POP_CONSTANT, pb.pop_a)
    got = e.eval_posterior()
    want = oracle.posterior(pb)
    assert rel_close(got.per_locus_coal, want.per_locus_coal)
    assert np.array_equal(got.per_locus_lik, r0.per_locus_lik) and e.last_node_updates() == 0


def test_root_renumbering_q8():
    """Swap the numbers of two internal nodes (what Tree.setRoot does to keep the root last): same tree, same logP."""
    pb = synth.make_problem("C2", n_loci=2, n_sites=100)
    e = Engine.from_problem(pb)
    r0 = e.eval_posterior()
    loc = pb.loci[0]
    t = loc.tree
    a, b = loc.n_tips, loc.n_tips + 3          # two internal non-root nodes
    perm = np.arange(t.n_nodes)
    perm[a], perm[b] = b, a
    inv = perm                                    # a transposition is its own inverse
    relabel = lambda x: np.where(x >= 0, inv[np.maximum(x, 0)], -1).astype(np.int32)
    new = TreeArrays(relabel(t.parent)[perm], relabel(t.left)[perm], relabel(t.right)[perm], t.height[perm])
    e.store()
    e.set_gene_tree(0, new)
    r1 = e.eval_posterior()
    assert rel_close(r1.per_locus_lik[0], r0.per_locus_lik[0], rel=1e-13)
    assert r1.per_locus_coal[0] == pytest.approx(r0.per_locus_coal[0], rel=1e-13)
    loc.tree = new
    assert rel_close(r1.per_locus_lik, oracle.posterior(pb).per_locus_lik)


def test_topology_change_nni():
    """Exchange-operator style NNI: swap a node's child with its sibling ('uncle')."""
    pb = synth.make_problem("C2", n_loci=2, n_sites=100, pop="analytic")   # analytic pops: any tree compatible? no -> check flag
    e = Engine.from_problem(pb)
    e.eval_posterior()
    loc = pb.loci[1]
    t = loc.tree.copy()
    root = t.root
    # find an internal node i whose parent p is not root-less and whose sibling u is lower than i's parent
    for i in range(loc.n_tips, t.n_nodes):
        p = int(t.parent[i])
        if p < 0:
            continue
        u = int(t.left[p]) if int(t.right[p]) == i else int(t.right[p])
        c = int(t.left[i])
        if t.height[u] < t.height[i]:
            # swap child c of i with uncle u
            if t.left[p] == u: t.left[p] = c
            else: t.right[p] = c
            t.left[i] = u
            t.parent[u], t.parent[c] = i, p
            break
    else:
        pytest.skip("no NNI candidate")
    e.store()
    loc.tree = t
    e.set_gene_tree(1, t)
    got = e.eval_posterior()
    want = oracle.posterior(pb)
    assert [math.isfinite(x) for x in got.per_locus_coal] == [math.isfinite(x) for x in want.per_locus_coal]
    if math.isfinite(want.total):
        assert rel_close(got.total, want.total)
    emb = oracle.gene_tree_update(t, loc.tip_species, pb.species)
    if emb.compatible:
        assert rel_close(got.per_locus_lik[1], want.per_locus_lik[1])


def test_lazy_store_edge_cases():
    """store() owes its device copy to the next evaluation (which brings the stored block up to date only where it differs).
    Sequences that stress that bookkeeping: restore without any evaluation since store, store after accept without
    evaluation, more than 8 loci touched in one step, parameter changes, and alternating loci."""
    pb = synth.make_problem("C2", n_loci=12, n_sites=60)
    e = Engine.from_problem(pb)
    rng = np.random.default_rng(21)
    base = e.eval_posterior()

    def same(a, b):
        return a.total == b.total and np.array_equal(a.per_locus_lik, b.per_locus_lik) and np.array_equal(a.per_locus_coal, b.per_locus_coal)

    def move(l):
        while True:   # a node slide that keeps the gene tree inside the species tree
            t, _ = synth.perturb_gene_tree(rng, pb.loci[l].tree, pb.loci[l].n_tips)
            if oracle.gene_tree_update(t, pb.loci[l].tip_species, pb.species, want_occupancy=False).compatible:
                break
        e.set_gene_tree(l, t)
        return t

    # (1) store, then restore with nothing evaluated in between (an operator that failed outright)
    e.store(); e.restore()
    assert same(e.eval_posterior(), base)
    e.store(); move(3); e.restore()                      # staged but never evaluated: dropped
    assert same(e.eval_posterior(), base)
    # (2) accept a move, store, restore without evaluation: the accepted state stands
    e.store(); pb.loci[2].tree = move(2); acc = e.eval_posterior(); e.accept()
    e.store(); e.restore()
    assert same(e.eval_posterior(), acc)
    want = oracle.posterior(pb)
    assert rel_close(acc.per_locus_lik, want.per_locus_lik)
    # (3) alternating loci, rejected every time: always back to `acc`, bit for bit
    for l in (5, 2, 7, 7, 0, 11):
        e.store(); move(l); r = e.eval_posterior(); assert not same(r, acc); e.restore()
        assert same(e.eval_posterior(), acc)
    # (4) ten loci in one step (more than the in-kernel sync list takes), rejected, then a single-locus step, rejected
    e.store()
    for l in range(10):
        move(l)
    e.eval_posterior(); e.restore()
    assert same(e.eval_posterior(), acc)
    e.store(); move(4); e.eval_posterior(); e.restore()
    assert same(e.eval_posterior(), acc)
    # (5) a parameter change (population sizes: every locus's coalescent terms move), rejected; then accepted
    e.store(); e.set_pop_model(pb.pop_kind, pb.pop_a * 1.1, pb.pop_b); r = e.eval_posterior(); assert r.coalescent != acc.coalescent; e.restore()
    assert same(e.eval_posterior(), acc)
    e.store(); e.set_pop_model(pb.pop_kind, pb.pop_a * 1.1, pb.pop_b); r2 = e.eval_posterior(); e.accept()
    assert r2.total == r.total
    e.store(); move(6); e.eval_posterior(); e.restore()
    assert same(e.eval_posterior(), r2)
    # (6) the whole thing still equals a from-scratch evaluation of the final state
    e.mark_all_dirty()
    pb.pop_a = pb.pop_a * 1.1
    final = e.eval_posterior()
    assert same(final, r2)
    assert rel_close(final.per_locus_lik, oracle.posterior(pb).per_locus_lik)
    e.close()
