"""Pins the CPU oracle against every known-answer value the reference's own tests hold for the
hot path.  Tolerance = the reference's own (10e-6), and a much tighter
one against the full-precision values re-derived in the survey."""
import math

import numpy as np
import pytest

import oracle
import ref_fixtures as rf
from starbeast2_b200.newick import leaf_sets, parse_newick


def _four():
    return rf.load_case(rf.SPECIES_4, rf.GENES_4, ["s0", "s1", "s2", "s3"])


def test_constant_populations_golden():
    # ConstantPopulationTest.java:50-80
    sp, genes = _four()
    total = 0.0
    for g, tipsp, _ in genes:
        emb = oracle.gene_tree_update(g, tipsp, sp)
        assert emb.compatible
        for b in range(sp.n_nodes):
            total += oracle.constant_logp(0.3, 2.0, emb.coalescent_times(b), emb.n[b], emb.k[b])
    assert total == pytest.approx(rf.EXPECTED["constant"], abs=rf.ALLOWED_ERROR)
    assert total == pytest.approx(2.0784641097740977, abs=1e-13)


def test_lwcr_golden():
    # LinearWithConstantRootTest.java:53-84; initPopSizes(0.3): tip 0.3, top 0.15 (LinearWithConstantRoot.java:73-94)
    sp, genes = _four()
    tip = np.full(sp.n_leaves, 0.3)
    top = np.full(sp.n_nodes - 1, 0.15)
    total = 0.0
    for g, tipsp, _ in genes:
        emb = oracle.gene_tree_update(g, tipsp, sp)
        for b in range(sp.n_nodes):
            total += oracle.lwcr_branch_logp(b, sp, tip, top, 2.0, emb.coalescent_times(b), emb.n[b], emb.k[b])
    assert total == pytest.approx(rf.EXPECTED["lwcr"], abs=rf.ALLOWED_ERROR)
    assert total == pytest.approx(2.887976975975221, abs=1e-13)


@pytest.mark.parametrize("genes_nw,key,full", [(rf.GENES_4, "analytic", -14.523398476213451),
                                               (rf.GENES_4_MISSING, "analytic_missing", -10.956285249389675)])
def test_analytic_golden(genes_nw, key, full):
    # ConstantIOTest.java:37-39,59-65 ; MissingDataConstantIOTest.java:59-65.  alpha 1.5, beta = mean*(alpha-1) = 2.5
    sp, genes = rf.load_case(rf.SPECIES_4, genes_nw, ["s0", "s1", "s2", "s3"])
    embs = [oracle.gene_tree_update(g, ts, sp) for g, ts, _ in genes]
    total = 0.0
    for b in range(sp.n_nodes):
        total += oracle.analytical_logp(1.5, 2.5, [2.0, 2.0], [e.coalescent_times(b) for e in embs],
                                        [e.n[b] for e in embs], [e.k[b] for e in embs])
    assert total == pytest.approx(rf.EXPECTED[key], abs=rf.ALLOWED_ERROR)
    assert total == pytest.approx(full, abs=1e-12)


def test_incompatible_tree_golden():
    # IncompatibleTreeTest.java:51-76 expects -Infinity
    sp, genes = rf.load_case(rf.SPECIES_8, rf.GENES_8, [f"T{i+1}" for i in range(8)])
    flags = [oracle.gene_tree_update(g, ts, sp).compatible for g, ts, _ in genes]
    assert not all(flags)


def _ucln_rates(tree, labels, bins_by_set, n_bins, mean, estimate_root):
    sets = leaf_sets(tree, labels)
    S = tree.n_nodes
    n_est = S if estimate_root else S - 1
    idx = np.ones(n_est, np.int32)  # initialBranchRate = 1
    for i, s in enumerate(sets):
        key = tuple(sorted(s))
        if key in bins_by_set and i < n_est:
            idx[i] = bins_by_set[key]
    bin_rates = oracle.ucln_bin_rates(n_bins, 1.0)
    return oracle.uncorrelated_rates(S, idx, bin_rates, mean, estimate_root, tree.root), sets, bin_rates


def test_uncorrelated_rates_golden():
    # UncorrelatedRatesTest.java:33-52: nBins defaults to #estimated rates = 10, stdev 1, clock.rate 1.5
    tree, labels = parse_newick(rf.GENE_CLOCK)
    rates, sets, _ = _ucln_rates(tree, labels, rf.UCLN_BINS, 10, 1.5, False)
    for i, s in enumerate(sets):
        assert rates[i] == pytest.approx(rf.UCLN_EXPECTED[tuple(sorted(s))], abs=rf.ALLOWED_ERROR)


def test_ucln_bins_match_reference_comments():
    # StarbeastClockTest.java:106-110 quotes the 5-bin sigma=1 rates
    b = oracle.ucln_bin_rates(5, 1.0)
    for got, want in zip(b[[0, 1, 3, 4]], [0.1683767, 0.3590116, 1.0247006, 2.1848596]):
        assert got == pytest.approx(want, abs=1e-6)
    assert oracle.uced_bin_rates(4)[0] == pytest.approx(-math.log(1 - 0.125))


def test_starbeast_clock_golden():
    # StarbeastClockTest.java:52-80: species UCLN (5 bins: estimateRoot=true, 5 nodes), mean rate NOT given to the
    # species clock (so 1.0), gene clock.rate 1.5
    sp, genes = rf.load_case(rf.SPECIES_3, [rf.GENE_CLOCK], ["a", "b", "c"])
    sp_labels = ["a", "b", "c"]
    srates, _, _ = _ucln_rates(sp, sp_labels, rf.CLOCK_SPECIES_BINS, 5, 1.0, True)
    g, tipsp, glabels = genes[0]
    emb = oracle.gene_tree_update(g, tipsp, sp)
    assert emb.compatible
    rates = oracle.starbeast_clock(1.5, srates, emb.occupancy)
    for i, s in enumerate(leaf_sets(g, glabels)):
        assert rates[i] == pytest.approx(rf.CLOCK_EXPECTED[tuple(sorted(s))], abs=rf.ALLOWED_ERROR)


def test_committed_golden_files_are_current():
    """tests/golden/*.json (committed) agree with the oracle built here and with the reference's literals."""
    import json, os
    from starbeast2_b200 import synth
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    known = json.load(open(os.path.join(here, "reference_known_answers.json")))["cases"]
    assert known["ConstantPopulationTest"]["expected"] == rf.EXPECTED["constant"]
    assert abs(known["ConstantIOTest"]["oracle"] - known["ConstantIOTest"]["expected"]) < rf.ALLOWED_ERROR
    cases = json.load(open(os.path.join(here, "oracle_cases.json")))["cases"]
    for name, c in cases.items():
        kw = dict(c["make_problem"]); nm = kw.pop("name")
        r = oracle.posterior(synth.make_problem(nm, **kw))
        np.testing.assert_allclose(r.per_locus_lik, c["per_locus_lik"], rtol=1e-13)
        np.testing.assert_allclose(r.per_locus_coal, c["per_locus_coal"], rtol=1e-13)
        assert r.total == pytest.approx(c["total"], rel=1e-13)
        pb = synth.make_problem(nm, **kw)
        inf = float("inf")
        q = c["operator_queries"]
        for node, want in q["connecting_nodes"].items():
            mask, tip, root, cnt = oracle.connecting_nodes(pb, int(node))
            assert np.flatnonzero(mask).tolist() == want["connecting"] and cnt == want["count"]
            assert tip == (inf if want["tipward"] is None else want["tipward"]) and root == (inf if want["rootward"] is None else want["rootward"])
        cmap = np.zeros(pb.species.n_nodes, np.int32)
        cmap[: pb.species.n_leaves] = np.arange(pb.species.n_leaves)
        assert oracle.max_height(pb, cmap, q["max_height_center"]) == (inf if q["max_height"] is None else q["max_height"])
