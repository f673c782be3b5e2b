"""North star: "identical MCMC traces for a fixed seed over short chains".  No JVM exists here, so the chain is a small
host-side Metropolis-Hastings loop with BEAST's calling protocol (store -> operator -> posterior.calculateLogP ->
accept | restore) over the drop-in plugin classes; at every step the proposed state is also evaluated from scratch by the
CPU oracle.  The two evaluators must agree to 1e-9 AND take the same accept / reject decision at every step, so the
traces are identical by induction.  Operators mirror the reference's schedule in kind: gene-tree node slides (BEAST
Uniform), population-size scaling, an UpDown move on everything (all nodes dirty), and species-node moves that can make a
gene tree incompatible (-infinity, likelihood phase skipped)."""
import math

import numpy as np
import pytest

import oracle
from starbeast2_b200 import synth
from starbeast2_b200.plugins import (HKY, GTR, Alignment, CompoundDistribution, ConstantPopulations, Frequencies, GeneTree,
                                     IntegerParameter, LinearWithConstantRoot, MultispeciesCoalescent, RealParameter, SiteModel,
                                     SpeciesTreeParser, StarBeastClock, StrictClockModel, Taxon, TaxonSet, TreeLikelihood, TreeParser,
                                     UncorrelatedRates)
from starbeast2_b200.problem import CLOCK_SPECIES_RATES, POP_CONSTANT, POP_LWCR, SUBST_HKY

pytestmark = pytest.mark.gpu


def build(pb):
    """Wire a synth Problem into plugin objects the way the StarBEAST2 XML does."""
    n_sp = pb.species.n_leaves
    names = [f"sp{i}" for i in range(n_sp)]
    per = max(int(np.sum(pb.loci[0].tip_species == s)) for s in range(n_sp))
    superset = TaxonSet([TaxonSet(names[i], [Taxon(f"sp{i}_{j}") for j in range(per)]) for i in range(n_sp)])
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("arrays", pb.species, "labels", names, "taxonset", superset)
    params = []
    if pb.pop_kind == POP_CONSTANT:
        popSizes = RealParameter(value=" ".join(repr(float(x)) for x in pb.pop_a), dimension=len(pb.pop_a))
        popModel = ConstantPopulations(populationSizes=popSizes, speciesTree=speciesTree)
        params = [popSizes]
    else:
        tip = RealParameter(value=" ".join(repr(float(x)) for x in pb.pop_a), dimension=len(pb.pop_a))
        top = RealParameter(value=" ".join(repr(float(x)) for x in pb.pop_b), dimension=len(pb.pop_b))
        popModel = LinearWithConstantRoot(tipPopulationSizes=tip, topPopulationSizes=top, speciesTree=speciesTree)
        params = [tip, top]
    trees, geneTrees, likelihoods = [], [], []
    for loc in pb.loci:
        counts = {}
        labels = []
        for s in loc.tip_species:
            labels.append(f"sp{s}_{counts.get(int(s), 0)}")
            counts[int(s)] = counts.get(int(s), 0) + 1
        tree = TreeParser()
        tree.initByName("arrays", loc.tree, "labels", labels)
        trees.append(tree)
        gt = GeneTree()
        gt.initByName("tree", tree, "ploidy", loc.ploidy, "speciesTree", speciesTree, "populationModel", popModel)
        geneTrees.append(gt)
    rates = None
    if pb.loci[0].clock_kind == CLOCK_SPECIES_RATES:
        # species rates are given explicitly by the problem: wrap them in a fixed SpeciesTreeRates
        class FixedRates:
            def __init__(self, r): self.r = np.asarray(r, float)
            def getRatesArray(self): return self.r
        rates = FixedRates(pb.species_rates)
    for l, loc in enumerate(pb.loci):
        data = Alignment()
        data.initByName("patterns", loc.tip_states, "weights", loc.weights, "taxa", trees[l].labels)

        class Fixed:   # substitution model with the problem's exact parameters (HKY or an eigen system)
            def __init__(self, kind, p): self.kind, self.p = kind, p
            def engine_params(self): return self.kind, self.p
        site = SiteModel()
        site.initByName("substModel", Fixed(loc.subst_kind, loc.subst_params), "gammaCategoryCount", loc.n_cat)
        site.category_rates = (lambda loc=loc: (loc.cat_rates, loc.cat_props))
        if rates is not None:
            clock = StarBeastClock()
            clock.initByName("geneTree", geneTrees[l], "speciesTreeRates", rates, "clock.rate", RealParameter(value=repr(float(loc.clock_rate))))
        else:
            clock = StrictClockModel()
            clock.initByName("clock.rate", RealParameter(value=repr(float(loc.clock_rate))))
        tl = TreeLikelihood()
        tl.initByName("data", data, "tree", trees[l], "siteModel", site, "branchRateModel", clock)
        likelihoods.append(tl)
    posterior = CompoundDistribution()
    posterior.initByName("distribution", [MultispeciesCoalescent(distribution=geneTrees), CompoundDistribution(distribution=likelihoods)])
    return speciesTree, trees, params, posterior


@pytest.mark.parametrize("name,kw,steps", [("C2", dict(n_loci=4, n_sites=120), 160), ("C3", dict(n_loci=3, n_sites=90), 120)])
def test_fixed_seed_chain_matches_oracle_step_by_step(name, kw, steps):
    pb = synth.make_problem(name, **kw)
    speciesTree, trees, params, posterior = build(pb)
    state_nodes = [speciesTree] + trees + params
    rng = np.random.Generator(np.random.PCG64(777))
    cur = posterior.calculateLogP()
    assert cur == pytest.approx(oracle.posterior(pb).total, rel=1e-9)
    n_accept = n_inf = 0
    for step in range(steps):
        for sn in state_nodes:
            sn.store()
        posterior.store()
        kind = rng.choice(["slide", "slide", "slide", "pop", "updown", "species"])
        hr = 0.0
        if kind == "slide":
            l = int(rng.integers(0, len(trees)))
            t, node = synth.perturb_gene_tree(rng, trees[l].arrays, pb.loci[l].n_tips)
            trees[l].setHeight(node, t.height[node])
        elif kind == "pop":
            p = params[int(rng.integers(0, len(params)))]
            i = int(rng.integers(0, p.getDimension()))
            s = math.exp(rng.normal(0, 0.3))
            p.setValue(i, p.getValue(i) * s)
            hr = math.log(s)
        elif kind == "updown":
            s = math.exp(rng.normal(0, 0.02))
            speciesTree.scale(s)
            for t in trees:
                t.scale(s)
            for p in params:
                for i in range(p.getDimension()):
                    p.setValue(i, p.getValue(i) * s)
            hr = 0.0
        else:   # move one internal species node: may push it above a coalescence -> incompatible -> -inf
            sp = speciesTree.arrays
            cands = [i for i in range(sp.n_leaves, sp.n_nodes) if sp.parent[i] >= 0]
            i = int(rng.choice(cands))
            lo = max(sp.height[sp.left[i]], sp.height[sp.right[i]])
            hi = sp.height[sp.parent[i]]
            speciesTree.setHeight(i, lo + (hi - lo) * rng.uniform(0.02, 0.98))
        new = posterior.calculateLogP()
        # the same proposed state, from scratch, on the CPU
        pb.species = speciesTree.arrays
        for l, t in enumerate(trees):
            pb.loci[l].tree = t.arrays
        if pb.pop_kind == POP_CONSTANT:
            pb.pop_a = params[0].values
        else:
            pb.pop_a, pb.pop_b = params[0].values, params[1].values
        want = oracle.posterior(pb)
        want_total = want.total if math.isfinite(want.coalescent) else -math.inf
        if math.isfinite(want_total):
            assert new == pytest.approx(want_total, rel=1e-9), (step, kind)
        else:
            assert new == -math.inf, (step, kind)
            n_inf += 1
        logu = math.log(rng.uniform())
        accept_gpu = logu < (new - cur + hr)
        accept_cpu = logu < (want_total - cur + hr)
        assert accept_gpu == accept_cpu, (step, kind)          # identical decisions => identical traces
        if accept_gpu:
            for sn in state_nodes:
                sn.accept()
            posterior.accept()
            cur = new
            n_accept += 1
        else:
            for sn in state_nodes:
                sn.restore()
            posterior.restore()
            assert posterior.calculateLogP() == cur, (step, kind)   # restore is exact
    assert 0 < n_accept < steps      # the chain both accepted and rejected
    # BEAST's periodic from-scratch consistency check, at the end of the chain
    posterior.get("distribution")[0].session.engine.mark_all_dirty()
    assert posterior.calculateLogP() == pytest.approx(cur, rel=1e-12)
