"""The reference's own JUnit tests (/root/reference/src/sb2tests/*.java), restated against the drop-in plugin classes:
same construction by initByName, same inputs, same expected values and the reference's tolerance (10e-6).
Host-only logic runs on CPU; everything that evaluates a logP goes through libsb2gpu and is marked gpu."""
import math

import numpy as np
import pytest

import ref_fixtures as rf
from starbeast2_b200.plugins import (HKY, Alignment, CompoundDistribution, ConstantPopulations, DummyModel, Frequencies, GeneTree,
                                     IntegerParameter, LinearWithConstantRoot, MultispeciesCoalescent, PassthroughModel,
                                     RealParameter, SiteModel, SpeciesTreeParser, StarBeastClock, StrictClockModel, Taxon,
                                     TaxonSet, TreeLikelihood, TreeParser, UncorrelatedRates)

allowedError = rf.ALLOWED_ERROR


def generateSuperset(nSpecies=4, individualsPerSpecies=2, fmt_sp="s%d", fmt_tip="s%d_tip%d", base=0):
    superSetList = []
    for i in range(nSpecies):
        taxonList = [Taxon(fmt_tip % (i + base, j + base)) for j in range(individualsPerSpecies)]
        superSetList.append(TaxonSet(fmt_sp % (i + base), taxonList))
    return TaxonSet(superSetList)


# ---- host-only (CPU) ------------------------------------------------------------------------------------------------

def test_uncorrelated_rates_host():
    # UncorrelatedRatesTest.java:27-52
    testTree = TreeParser()
    testTree.initByName("newick", rf.GENE_CLOCK, "IsLabelledNewick", True)
    meanRateParameter, stdevParameter, branchRatesParameter = RealParameter(), RealParameter(), IntegerParameter()
    meanRateParameter.initByName("value", "1.5")
    stdevParameter.initByName("value", "1.0")
    branchRatesParameter.initByName("value", "1")
    clockModel = UncorrelatedRates()
    clockModel.initByName("tree", testTree, "rates", branchRatesParameter, "stdev", stdevParameter, "estimateRoot", False,
                          "clock.rate", meanRateParameter)
    assert branchRatesParameter.getDimension() == 10 and branchRatesParameter.getUpper() == 9
    for leaves, b in rf.UCLN_BINS.items():
        branchRatesParameter.setValue(testTree.findNode(leaves), b)
    for leaves, want in rf.UCLN_EXPECTED.items():
        assert clockModel.getRateForBranch(testTree.findNode(leaves)) == pytest.approx(want, abs=allowedError)


def test_input_validation_matches_reference_errors():
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("newick", rf.SPECIES_4, "IsLabelledNewick", True, "taxonset", generateSuperset())
    with pytest.raises(ValueError):   # Validate.REQUIRED
        GeneTree().initByName("speciesTree", speciesTree)
    with pytest.raises(ValueError):   # unknown input name
        ConstantPopulations().initByName("speciesTree", speciesTree, "popSizes", RealParameter(value="1.0"))
    alpha = RealParameter(value="1.5")
    with pytest.raises(ValueError):   # MultispeciesCoalescent.java:93-96: shape without mean
        MultispeciesCoalescent().initByName("populationShape", alpha, "distribution", [])
    with pytest.raises(ValueError):   # MultispeciesCoalescent.java:116-118: children must be GeneTrees
        MultispeciesCoalescent().initByName("distribution", [CompoundDistribution()])


def test_pop_model_dimensions_and_init():
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("newick", rf.SPECIES_4, "IsLabelledNewick", True, "taxonset", generateSuperset())
    tip, top = RealParameter(), RealParameter()
    tip.initByName("dimension", "4", "value", 1.0)
    top.initByName("dimension", "6", "value", 1.0)
    pm = LinearWithConstantRoot()
    pm.initByName("tipPopulationSizes", tip, "topPopulationSizes", top, "speciesTree", speciesTree)
    pm.initPopSizes(0.3)   # LinearWithConstantRoot.java:73-94
    assert np.all(tip.values == 0.3) and np.all(top.values == 0.15) and top.getDimension() == 6
    ps = RealParameter(value="0.3")
    cp = ConstantPopulations()
    cp.initByName("populationSizes", ps, "speciesTree", speciesTree)
    assert ps.getDimension() == 7   # ConstantPopulations.java:37
    assert PassthroughModel(childModel=cp).getBaseModel() is cp


def test_stale_alpha_quirk_q1():
    """checkHyperparameters uses the previous alpha for beta (MultispeciesCoalescent.java:137-138)."""
    alpha, mean = RealParameter(value="1.5"), RealParameter(value="5.0")
    msc = MultispeciesCoalescent.__new__(MultispeciesCoalescent)
    MultispeciesCoalescent.__init__(msc)
    msc._inputs.update(populationShape=alpha, populationMean=mean)
    msc.alpha = msc.beta = 0.0
    msc._check_hyperparameters(True)
    assert (msc.alpha, msc.beta) == (1.5, -5.0)            # force at init: beta = mean * (0 - 1)
    msc._check_hyperparameters(False)
    assert (msc.alpha, msc.beta) == (1.5, 2.5)             # first calculateLogP fixes it
    alpha.setValue(2.0)
    msc._check_hyperparameters(False)
    assert (msc.alpha, msc.beta) == (2.0, 2.5)             # beta lags alpha by one evaluation
    msc._check_hyperparameters(False)
    assert (msc.alpha, msc.beta) == (2.0, 5.0)


def test_alpha_beta_are_stored_and_restored_like_the_reference():
    """MultispeciesCoalescent.java:50-88: store() saves (alpha, beta), restore() swaps them back -- so after a REJECTED
    populationShape move beta is recomputed from the restored alpha, not from the rejected one."""
    class _Session:   # stands in for the device session: this test is about the host-side bookkeeping only
        def store(self): pass
        def restore(self): pass
    alpha, mean = RealParameter(value="1.5"), RealParameter(value="5.0")
    msc = MultispeciesCoalescent.__new__(MultispeciesCoalescent)
    MultispeciesCoalescent.__init__(msc)
    msc._inputs.update(populationShape=alpha, populationMean=mean)
    msc.alpha = msc.beta = msc.storedAlpha = msc.storedBeta = 0.0
    msc.dontCalculate, msc.session = False, _Session()
    msc._check_hyperparameters(True)
    msc._check_hyperparameters(False)
    assert (msc.alpha, msc.beta) == (1.5, 2.5)
    msc.store()
    alpha.setValue(0.5)                                    # proposal: alpha < 1
    msc._check_hyperparameters(False)
    assert (msc.alpha, msc.beta) == (0.5, 2.5)
    alpha.setValue(1.5)                                    # rejected: the state node goes back ...
    msc.restore()                                          # ... and so do the cached hyperparameters
    assert (msc.alpha, msc.beta) == (1.5, 2.5)
    msc._check_hyperparameters(False)
    assert (msc.alpha, msc.beta) == (1.5, 2.5)             # without the restore this would be (1.5, 5 * (0.5 - 1)) = (1.5, -2.5)


# ---- through the GPU ------------------------------------------------------------------------------------------------

def _gene_trees(speciesTree, newicks, popModel=None, ploidy=2.0):
    out = []
    for nw in newicks:
        geneTree = TreeParser()
        geneTree.initByName("newick", nw, "IsLabelledNewick", True)
        w = GeneTree()
        args = ["tree", geneTree, "ploidy", ploidy, "speciesTree", speciesTree]
        if popModel is not None:
            args += ["populationModel", popModel]
        w.initByName(*args)
        out.append(w)
    return out


@pytest.mark.gpu
def test_ConstantPopulationTest():
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("newick", rf.SPECIES_4, "IsLabelledNewick", True, "taxonset", generateSuperset())
    popsizeParameter = RealParameter()
    popsizeParameter.initByName("value", "0.3", "dimension", "7")
    popModel = ConstantPopulations()
    popModel.initByName("populationSizes", popsizeParameter, "speciesTree", speciesTree)
    calculatedLogP = 0.0
    for w in _gene_trees(speciesTree, rf.GENES_4, popModel):
        calculatedLogP += w.calculateLogP()
    assert calculatedLogP == pytest.approx(rf.EXPECTED["constant"], abs=allowedError)


@pytest.mark.gpu
def test_LinearWithConstantRootTest():
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("newick", rf.SPECIES_4, "IsLabelledNewick", True, "taxonset", generateSuperset())
    tip, top = RealParameter(), RealParameter()
    tip.initByName("dimension", "4", "value", 1.0)
    top.initByName("dimension", "6", "value", 1.0)
    popModel = LinearWithConstantRoot()
    popModel.initByName("tipPopulationSizes", tip, "topPopulationSizes", top, "speciesTree", speciesTree)
    popModel.initPopSizes(0.3)
    calculatedLogP = sum(w.calculateLogP() for w in _gene_trees(speciesTree, rf.GENES_4, popModel))
    assert calculatedLogP == pytest.approx(rf.EXPECTED["lwcr"], abs=allowedError)


@pytest.mark.gpu
@pytest.mark.parametrize("genes,key", [(rf.GENES_4, "analytic"), (rf.GENES_4_MISSING, "analytic_missing")])
def test_ConstantIOTest(genes, key):
    alpha, beta = 1.5, 2.5
    alphaParameter, meanParameter = RealParameter(), RealParameter()
    alphaParameter.initByName("value", str(alpha))
    meanParameter.initByName("value", str(beta / (alpha - 1.0)))
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("newick", rf.SPECIES_4, "IsLabelledNewick", True, "taxonset", generateSuperset())
    wrappers = _gene_trees(speciesTree, genes)
    msc = MultispeciesCoalescent()
    msc.initByName("populationShape", alphaParameter, "populationMean", meanParameter, "distribution", wrappers)
    for gt in wrappers:
        assert gt.calculateLogP() == 0.0     # MissingDataConstantIOTest.java:97
    assert msc.calculateLogP() == pytest.approx(rf.EXPECTED[key], abs=allowedError)


@pytest.mark.gpu
def test_rejected_population_shape_move_restores_the_analytic_term():
    alphaParameter, meanParameter = RealParameter(value="1.5"), RealParameter(value="5.0")
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("newick", rf.SPECIES_4, "IsLabelledNewick", True, "taxonset", generateSuperset())
    msc = MultispeciesCoalescent()
    msc.initByName("populationShape", alphaParameter, "populationMean", meanParameter, "distribution", _gene_trees(speciesTree, rf.GENES_4))
    first = msc.calculateLogP()
    assert first == pytest.approx(rf.EXPECTED["analytic"], abs=allowedError)
    msc.store()
    alphaParameter.setValue(0.5)                           # proposal with alpha < 1
    msc.calculateLogP()
    alphaParameter.setValue(1.5)                           # rejected
    msc.restore()
    assert msc.hyperparameters() == (1.5, 2.5)
    assert msc.calculateLogP() == first                    # not NaN from beta = mean * (0.5 - 1)


@pytest.mark.gpu
def test_IncompatibleTreeTest():
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("newick", rf.SPECIES_8, "IsLabelledNewick", True, "taxonset",
                           generateSuperset(8, 2, "T%d", "T%d_%d", base=1))
    popsizeParameter = RealParameter()
    popsizeParameter.initByName("value", "0.3", "dimension", "15")
    popModel = ConstantPopulations()
    popModel.initByName("populationSizes", popsizeParameter, "speciesTree", speciesTree)
    calculatedLogP = sum(w.calculateLogP() for w in _gene_trees(speciesTree, rf.GENES_8, popModel))
    assert calculatedLogP == -math.inf


@pytest.mark.gpu
def test_StarbeastClockTest():
    superset = TaxonSet([TaxonSet(s, [Taxon(f"{s}{j + 1}") for j in range(2)]) for s in "abc"])
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("newick", rf.SPECIES_3, "IsLabelledNewick", True, "taxonset", superset)
    geneTree = TreeParser()
    geneTree.initByName("newick", rf.GENE_CLOCK, "IsLabelledNewick", True)
    geneTreeWrapper = GeneTree()
    geneTreeWrapper.initByName("tree", geneTree, "ploidy", 2.0, "speciesTree", speciesTree)
    meanRateParameter, stdevParameter, branchRatesParameter = RealParameter(), RealParameter(), IntegerParameter()
    meanRateParameter.initByName("value", "1.5")
    stdevParameter.initByName("value", "1.0")
    branchRatesParameter.initByName("value", "1")
    speciesTreeClock = UncorrelatedRates()
    speciesTreeClock.initByName("tree", speciesTree, "rates", branchRatesParameter, "stdev", stdevParameter, "estimateRoot", True)
    for leaves, b in rf.CLOCK_SPECIES_BINS.items():
        branchRatesParameter.setValue(speciesTree.findNode(leaves), b)
    geneTreeClock = StarBeastClock()
    geneTreeClock.initByName("geneTree", geneTreeWrapper, "speciesTreeRates", speciesTreeClock, "clock.rate", meanRateParameter)
    for leaves, want in rf.CLOCK_EXPECTED.items():
        assert geneTreeClock.getRateForBranch(geneTree.findNode(leaves)) == pytest.approx(want, abs=allowedError)


@pytest.mark.gpu
def test_posterior_protocol_with_tree_likelihoods():
    """A StarBEAST2-style posterior wired like tutorial/CanisPhylogeny/Canis-Phylogeny.xml:551-700: speciescoalescent first,
    then likelihood; store / propose / evaluate / restore; the short-circuit on an incompatible proposal."""
    import oracle
    from starbeast2_b200 import synth
    pb = synth.make_problem("C2", n_loci=3, n_sites=120)
    n_sp = pb.species.n_leaves
    sp_names = [f"sp{i}" for i in range(n_sp)]
    superset = TaxonSet([TaxonSet(sp_names[i], [Taxon(f"sp{i}_{j}") for j in range(3)]) for i in range(n_sp)])
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("arrays", pb.species, "labels", sp_names, "taxonset", superset)
    popSizes = RealParameter()
    popSizes.initByName("value", " ".join(repr(float(x)) for x in pb.pop_a), "dimension", str(len(pb.pop_a)))
    popModel = PassthroughModel()
    popModel.initByName("childModel", ConstantPopulations(populationSizes=popSizes, speciesTree=speciesTree))
    geneTrees, likelihoods, trees = [], [], []
    for l, loc in enumerate(pb.loci):
        labels = [f"sp{loc.tip_species[t]}_{t % 3}" for t in range(loc.n_tips)]
        tree = TreeParser()
        tree.initByName("arrays", loc.tree, "labels", labels)
        trees.append(tree)
        gt = GeneTree()
        gt.initByName("tree", tree, "ploidy", loc.ploidy, "speciesTree", speciesTree, "populationModel", popModel)
        geneTrees.append(gt)
    for l, loc in enumerate(pb.loci):
        data = Alignment()
        data.initByName("patterns", loc.tip_states, "weights", loc.weights, "taxa", trees[l].labels)
        freqs = Frequencies()
        freqs.initByName("frequencies", RealParameter(value=" ".join(map(str, loc.subst_params[1:5])), dimension=4))
        site = SiteModel()
        site.initByName("substModel", HKY(kappa=RealParameter(value=str(loc.subst_params[0])), frequencies=freqs),
                        "mutationRate", float(loc.cat_rates[0]))
        clock = StrictClockModel()
        clock.initByName("clock.rate", RealParameter(value=repr(float(loc.clock_rate))))
        tl = TreeLikelihood()
        tl.initByName("data", data, "tree", trees[l], "siteModel", site, "branchRateModel", clock)
        likelihoods.append(tl)
    speciescoalescent = MultispeciesCoalescent()
    speciescoalescent.initByName("distribution", geneTrees)
    likelihood = CompoundDistribution()
    likelihood.initByName("distribution", likelihoods)
    posterior = CompoundDistribution()
    posterior.initByName("distribution", [speciescoalescent, likelihood])

    want = oracle.posterior(pb)
    assert posterior.calculateLogP() == pytest.approx(want.total, rel=1e-9)
    assert [gt.logP for gt in geneTrees] == pytest.approx(list(want.per_locus_coal), rel=1e-9)
    assert [tl.logP for tl in likelihoods] == pytest.approx(list(want.per_locus_lik), rel=1e-9)

    # propose: slide a node of locus 1; evaluate; reject; state and value come back
    rng = np.random.default_rng(3)
    before = posterior.logP
    posterior.store()
    trees[1].store()
    new_tree, node = synth.perturb_gene_tree(rng, trees[1].arrays, pb.loci[1].n_tips)
    trees[1].setHeight(node, new_tree.height[node])
    pb.loci[1].tree = trees[1].arrays
    proposed = posterior.calculateLogP()
    w2 = oracle.posterior(pb)
    assert proposed == pytest.approx(w2.total if math.isfinite(w2.coalescent) else -math.inf, rel=1e-9)
    trees[1].restore()          # BEAST: state.restore() first ...
    posterior.restore()         # ... then restoreCalculationNodes()
    assert posterior.calculateLogP() == before

    # an incompatible proposal: the coalescent is -inf and no TreeLikelihood is asked for its value.  Strict protocol
    # (SB2_NO_SPECULATION): the pruning kernel is never launched; default: it ran speculatively in the same device pass and
    # its result is dropped with the restore.
    import os
    for strict in (True, False):
        if strict:
            os.environ["SB2_NO_SPECULATION"] = "1"
        else:
            os.environ.pop("SB2_NO_SPECULATION", None)
        try:
            launches_lik = likelihoods[0].session.engine.launch_count()
            posterior.store(); speciesTree.store()
            root = speciesTree.getRootNr()
            child = int(speciesTree.arrays.left[root]) if speciesTree.arrays.left[int(speciesTree.arrays.left[root])] >= 0 else int(speciesTree.arrays.right[root])
            speciesTree.setHeight(child, speciesTree.arrays.height[root] * 0.9999)
            assert posterior.calculateLogP() == -math.inf
            eng = likelihoods[0].session.engine
            # strict: k_prepare + the reduction, the pruning kernel never launched; default: + the speculative pruning kernel
            # (a species-tree move re-embeds every locus: the three-kernel path; single-locus moves take ONE launch, k_step)
            assert eng.launch_count() - launches_lik == (2 if strict else 3)
            speciesTree.restore(); posterior.restore()
            assert posterior.calculateLogP() == before
        finally:
            os.environ.pop("SB2_NO_SPECULATION", None)


# ---- BASELINE.json configs[0]: the Canis tutorial data (16 loci x 16 individuals / 8 species, HKY, constant pops) ----------

def _canis():
    import json, os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "canis_c1.json")))["loci"]


def test_canis_fixture_matches_reference_fasta():
    """Alignment (calcPatterns) on the reference's FASTA files reproduces the committed fixture (runs where the reference is)."""
    import os
    from starbeast2_b200.io import read_fasta
    data = "/root/reference/tutorial/data"
    if not os.path.isdir(data):
        pytest.skip("reference tree not present on this machine")
    for name, fx in _canis().items():
        a = Alignment()
        a.initByName("sequences", read_fasta(os.path.join(data, name + ".fasta")))
        assert a.taxa == fx["taxa"] and a.weights.tolist() == fx["weights"]
        assert ["".join("ACGT-"[s] for s in row) for row in a.patterns] == fx["patterns"]
        assert int(a.weights.sum()) == fx["sites"]
    assert len(_canis()) == 16 and 8 <= min(len(v["weights"]) for v in _canis().values()) and max(len(v["weights"]) for v in _canis().values()) == 41


@pytest.mark.gpu
def test_canis_c1_posterior_through_plugins():
    """The tutorial's model wiring on the tutorial's alignments (gaps included), random MSC-compatible trees; the CUDA path
    against the oracle, locus by locus."""
    import oracle
    from starbeast2_b200 import synth
    from starbeast2_b200.problem import POP_CONSTANT, Locus, Problem
    fx = _canis()
    species = sorted({t[:-2] for v in fx.values() for t in v["taxa"]})
    assert len(species) == 8
    rng = np.random.Generator(np.random.PCG64(20260101))
    sp_arrays = synth.yule_species_tree(rng, 8)
    superset = TaxonSet([TaxonSet(s, [Taxon(s + "_a"), Taxon(s + "_b")]) for s in species])
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("arrays", sp_arrays, "labels", species, "taxonset", superset)
    pop = rng.gamma(2.0, 0.025, size=15)
    popSizes = RealParameter(value=" ".join(repr(float(x)) for x in pop), dimension=15)
    popModel = PassthroughModel(childModel=ConstantPopulations(populationSizes=popSizes, speciesTree=speciesTree))
    clockRate = RealParameter(value="0.05")
    clock = StrictClockModel()
    clock.initByName("clock.rate", clockRate)       # one clock shared by all loci (Canis-Phylogeny.xml:692-694)
    geneTrees, likelihoods, loci = [], [], []
    code = {c: i for i, c in enumerate("ACGT-")}
    for name, v in fx.items():
        arrays, tip_species = synth.simulate_gene_tree(rng, sp_arrays, pop * 2.0, np.array([2] * 8))
        labels = [species[s] + ("_a" if k % 2 == 0 else "_b") for k, s in enumerate(tip_species)]
        tree = TreeParser()
        tree.initByName("arrays", arrays, "labels", labels)
        gt = GeneTree()
        gt.initByName("tree", tree, "speciesTree", speciesTree, "populationModel", popModel)   # ploidy defaults to 2.0
        geneTrees.append(gt)
        pats = np.array([[code[ch] for ch in row] for row in v["patterns"]], np.uint8)
        data = Alignment()
        data.initByName("patterns", pats, "weights", v["weights"], "taxa", v["taxa"])
        hky = HKY(kappa=RealParameter(value="2.0"), frequencies=Frequencies(data=data))   # empirical frequencies (:689)
        site = SiteModel()
        site.initByName("substModel", hky, "mutationRate", RealParameter(value=str(float(rng.lognormal(0, 0.3)))))
        tl = TreeLikelihood()
        tl.initByName("data", data, "tree", tree, "siteModel", site, "branchRateModel", clock)
        likelihoods.append(tl)
        kind, sparams = hky.engine_params()
        rates, props = site.category_rates()
        st, w = data.patterns_for(labels)
        loci.append(Locus(tree=arrays, tip_states=st, weights=w, tip_species=tip_species, subst_kind=kind, subst_params=sparams,
                          cat_rates=rates, cat_props=props, clock_rate=0.05))
    posterior = CompoundDistribution()
    posterior.initByName("distribution", [MultispeciesCoalescent(distribution=geneTrees), CompoundDistribution(distribution=likelihoods)])
    got = posterior.calculateLogP()
    want = oracle.posterior(Problem(species=sp_arrays, loci=loci, pop_kind=POP_CONSTANT, pop_a=pop))
    assert got == pytest.approx(want.total, rel=1e-9)
    for tl, w in zip(likelihoods, want.per_locus_lik):
        assert tl.logP == pytest.approx(w, rel=1e-9)
    for gt, w in zip(geneTrees, want.per_locus_coal):
        assert gt.logP == pytest.approx(w, rel=1e-9)
    # kappa scaler on one locus (Canis-Phylogeny.xml operators): only that locus's likelihood is recomputed
    eng = likelihoods[0].session.engine
    posterior.store()
    likelihoods[3].get("siteModel").get("substModel").get("kappa").setValue(3.3)
    loci[3].subst_params = likelihoods[3].get("siteModel").get("substModel").engine_params()[1]
    got2 = posterior.calculateLogP()
    assert eng.last_node_updates() == 15
    assert got2 == pytest.approx(oracle.posterior(Problem(species=sp_arrays, loci=loci, pop_kind=POP_CONSTANT, pop_a=pop)).total, rel=1e-9)


# ---------------------------------------------------------------------------------------------------------------------
# UniformPopulations and RandomLocalRates (host-side models that feed the same engine inputs)
# ---------------------------------------------------------------------------------------------------------------------

def test_random_local_rates_mirror_matches_oracle_and_hand_worked_case():
    """No reference test exists for RandomLocalRates (parity unpinned): hand-worked 3-species case + oracle recursion."""
    from starbeast2_b200 import synth
    from starbeast2_b200.plugins import BooleanParameter, RandomLocalRates, RealParameter, SpeciesTreeParser
    from starbeast2_b200.problem import TreeArrays
    import oracle
    # ((A:1,B:1)ab:1,C:2)root -- nodes A0 B1 C2 ab3 root4
    sp = TreeArrays([3, 3, 4, 4, -1], [-1, -1, -1, 0, 3], [-1, -1, -1, 1, 2], [0, 0, 0, 1.0, 2.0])
    tree = SpeciesTreeParser()
    tree.initByName("arrays", sp, "labels", ["A", "B", "C"])
    rates = RealParameter(value="2.0 3.0 4.0 5.0", dimension=4)
    ind = BooleanParameter(value="false true false true", dimension=4)
    rlr = RandomLocalRates(tree=tree, rates=rates, indicators=ind)
    got = rlr.getRatesArray()
    # unscaled: ab switches to 5 (indicator on), A inherits 5, B switches to 3, C inherits the root's 1; lengths A1 B1 C2 ab1
    raw = np.array([5.0, 3.0, 1.0, 5.0])
    scale = 1.0 * (1 + 1 + 2 + 1) / (5 * 1 + 3 * 1 + 1 * 2 + 5 * 1)
    assert np.allclose(got[:4], raw * scale, rtol=1e-15)
    assert got[4] == 1.0                                      # quirk Q9: the root keeps the recursion's start rate
    want = oracle.random_local_rates(sp, ind.values, rates.values, 1.0)
    assert np.array_equal(got, want)
    # with a clock.rate the weighted mean of the non-root rates equals it
    rlr2 = RandomLocalRates(tree=tree, rates=rates, indicators=ind, **{"clock.rate": RealParameter(value="0.01")})
    r2 = rlr2.getRatesArray()
    lengths = np.array([1.0, 1.0, 2.0, 1.0])
    assert np.dot(r2[:4], lengths) / lengths.sum() == pytest.approx(0.01, rel=1e-14)
    assert np.array_equal(r2, oracle.random_local_rates(sp, ind.values, rates.values, 0.01))
    rng = np.random.default_rng(3)
    pb = synth.make_problem("C3", n_loci=1, n_sites=10)
    S = pb.species.n_nodes
    t2 = SpeciesTreeParser()
    t2.initByName("arrays", pb.species, "labels", [f"s{i}" for i in range(pb.species.n_leaves)])
    for _ in range(5):
        rr = RealParameter(value=" ".join(repr(float(x)) for x in rng.gamma(2.0, 0.5, S - 1)), dimension=S - 1)
        ii = BooleanParameter(value=" ".join("true" if x else "false" for x in rng.random(S - 1) < 0.3), dimension=S - 1)
        m = RandomLocalRates(tree=t2, rates=rr, indicators=ii)
        assert np.array_equal(m.getRatesArray(), oracle.random_local_rates(pb.species, ii.values, rr.values, 1.0))


@pytest.mark.gpu
def test_uniform_populations_and_random_local_rates_through_the_engine():
    """UniformPopulations == ConstantPopulations with one size everywhere; RandomLocalRates feeds StarBeastClock."""
    import oracle
    from starbeast2_b200 import synth
    from test_mcmc_trace import build
    from starbeast2_b200.plugins import (BooleanParameter, GeneTree, MultispeciesCoalescent, RandomLocalRates, RealParameter, StarBeastClock,
                                         UniformPopulations)
    pb = synth.make_problem("C3", n_loci=3, n_sites=80)          # species-tree clock, LWCR pops
    speciesTree, trees, params, posterior = build(pb)
    # swap the population model for UniformPopulations(0.03) on fresh GeneTrees of a second session-less check via the oracle
    S = pb.species.n_nodes
    rng = np.random.default_rng(1)
    rr = RealParameter(value=" ".join(repr(float(x)) for x in rng.gamma(2.0, 0.5, S - 1)), dimension=S - 1)
    ii = BooleanParameter(value=" ".join("true" if x else "false" for x in rng.random(S - 1) < 0.3), dimension=S - 1)
    rlr = RandomLocalRates(tree=speciesTree, rates=rr, indicators=ii)
    likelihoods = posterior.get("distribution")[1].get("distribution")
    for tl in likelihoods:
        tl.get("branchRateModel")._inputs["speciesTreeRates"] = rlr
    got = posterior.calculateLogP()
    pb.species_rates = rlr.getRatesArray()
    want = oracle.posterior(pb)
    assert got == pytest.approx(want.total, rel=1e-9)
    # an indicator flip changes the rates of a whole clade: picked up by the session, still equal to the oracle
    ii.setValue(S - 3, not bool(ii.getValue(S - 3)))
    got = posterior.calculateLogP()
    pb.species_rates = rlr.getRatesArray()
    assert got == pytest.approx(oracle.posterior(pb).total, rel=1e-9)

    pb2 = synth.make_problem("C2", n_loci=3, n_sites=80)
    from starbeast2_b200.plugins import SpeciesTreeParser, Taxon, TaxonSet, TreeParser
    n_sp = pb2.species.n_leaves
    names = [f"sp{i}" for i in range(n_sp)]
    superset = TaxonSet([TaxonSet(names[i], [Taxon(f"sp{i}_{j}") for j in range(3)]) for i in range(n_sp)])
    st = SpeciesTreeParser()
    st.initByName("arrays", pb2.species, "labels", names, "taxonset", superset)
    uni = UniformPopulations(speciesTree=st, universalSize=RealParameter(value="0.03"))
    gts = []
    for loc in pb2.loci:
        counts, labels = {}, []
        for sidx in loc.tip_species:
            labels.append(f"sp{sidx}_{counts.get(int(sidx), 0)}")
            counts[int(sidx)] = counts.get(int(sidx), 0) + 1
        t = TreeParser()
        t.initByName("arrays", loc.tree, "labels", labels)
        gts.append(GeneTree(tree=t, speciesTree=st, ploidy=loc.ploidy, populationModel=uni))
    msc = MultispeciesCoalescent(distribution=gts)
    got = msc.calculateLogP()
    pb2.pop_a = np.full(pb2.species.n_nodes, 0.03)
    want = oracle.posterior(pb2)
    assert got == pytest.approx(float(np.sum(want.per_locus_coal)), rel=1e-12)
    uni.get("universalSize").setValue(0, 0.05)
    pb2.pop_a = np.full(pb2.species.n_nodes, 0.05)
    assert msc.calculateLogP() == pytest.approx(float(np.sum(oracle.posterior(pb2).per_locus_coal)), rel=1e-12)


def test_fasta_to_alignment_patterns_and_frequencies(tmp_path):
    """Data format on the way into the path: FASTA (the reference ships tutorial/data/*.fasta) ->
    Alignment -> unique site patterns in lexicographic order, weights, empirical frequencies."""
    from starbeast2_b200.io import read_fasta
    fa = tmp_path / "locus.fasta"
    fa.write_text(">b_2 some description\nACGTAC\nGTNN\n>a_1\nACGTACGT-A\n>c_3\nACGAACGTCA\n")
    seqs = read_fasta(str(fa))
    assert seqs == {"b_2": "ACGTACGTNN", "a_1": "ACGTACGT-A", "c_3": "ACGAACGTCA"}
    aln = Alignment(sequences=seqs)
    assert aln.taxa == ["a_1", "b_2", "c_3"]                       # taxa sorted by name
    full = np.array([[0, 1, 2, 3, 0, 1, 2, 3, 4, 0], [0, 1, 2, 3, 0, 1, 2, 3, 4, 4], [0, 1, 2, 0, 0, 1, 2, 3, 1, 0]], np.uint8)
    assert aln.getPatternCount() == 7 and int(aln.weights.sum()) == 10
    # every site is represented by its pattern with the right multiplicity
    for col in full.T:
        j = [k for k in range(aln.getPatternCount()) if np.array_equal(aln.patterns[:, k], col)]
        assert len(j) == 1 and aln.weights[j[0]] == int(np.sum(np.all(full.T == col, axis=1)))
    assert all(tuple(aln.patterns[:, k]) < tuple(aln.patterns[:, k + 1]) for k in range(aln.getPatternCount() - 1))   # lexicographic
    states, w = aln.patterns_for(["c_3", "a_1", "b_2"])           # rows follow the gene tree's tip order
    assert np.array_equal(states[1], aln.patterns[0]) and w is aln.weights
    f = aln.empirical_frequencies()
    counts = np.array([(full == s).sum() for s in range(4)], float)
    assert np.allclose(f, counts / counts.sum()) and abs(f.sum() - 1) < 1e-15


# ---- morphological TreeLikelihoods on the species tree (tutorial/CanisFBD/Canis-FBD.xml:964-1033) -------------------------------

def _morph_sequences(pb, rng, n_real=(9, 6, 5, 4)):
    """A morphological matrix laid out like the reference's `morphology_canis`: real characters of mixed state counts in
    arbitrary order, then dummy constant columns '0...', '1...', ... '4...' at the end.  Returns (sequences by species label,
    {K: (1-based filter string, excludefrom, excludeto)})."""
    from starbeast2_b200 import synth
    sp = pb.species
    names = [f"sp{i}" for i in range(sp.n_leaves)]
    cols, kinds = [], []
    for K, n in zip((2, 3, 4, 5), n_real):
        pats, w, _ = synth.simulate_morphology(rng, sp, K, n, 0.6 / sp.height[sp.root], missing_frac=0.1)
        real = np.repeat(pats[:, :-K], w[:-K], axis=1)
        for c in range(real.shape[1]):
            cols.append(real[:, c]); kinds.append(K)
    order = rng.permutation(len(cols))
    cols, kinds = [cols[i] for i in order], [kinds[i] for i in order]
    n_chars = len(cols)
    for s in range(5):
        cols.append(np.full(sp.n_leaves, s, np.uint8)); kinds.append(-1)
    seqs = {names[t]: "".join("?" if (k > 0 and c[t] >= k) else str(int(c[t])) for c, k in zip(cols, kinds)) for t in range(sp.n_leaves)}
    filt = {}
    for K in (2, 3, 4, 5):
        mine = [i + 1 for i, k in enumerate(kinds) if k == K]
        filt[K] = (",".join(map(str, mine)) + f",{n_chars + 1}-{n_chars + K}", len(mine), len(mine) + K)
    return seqs, filt, names


def test_filtered_alignment_ascertained_patterns():
    from starbeast2_b200.plugins import FilteredAlignment, StandardData
    seqs = {"a": "01?10" + "01", "b": "011-0" + "01", "c": "00100" + "01"}
    data = Alignment(sequences=seqs)
    fa = FilteredAlignment()
    fa.initByName("data", data, "filter", "1-3,5,6-7", "userDataType", StandardData(nrOfStates=2), "ascertained", True,
                  "excludefrom", 4, "excludeto", 6)
    assert fa.getStateCount() == 2 and fa.taxa == ["a", "b", "c"]
    cols = {tuple(fa.patterns[:, p]): int(fa.weights[p]) for p in range(fa.getPatternCount())}
    # site 5 ("000") coincides with the dummy all-0 column: BEAST zeroes the pattern's weight ("so it does not confuse the tree likelihood")
    assert cols == {(0, 0, 0): 0, (1, 1, 0): 1, (2, 1, 1): 1, (1, 1, 1): 0}
    ex = {tuple(fa.patterns[:, p]) for p in fa.excluded_patterns()}
    assert ex == {(0, 0, 0), (1, 1, 1)}
    with pytest.raises(ValueError):
        FilteredAlignment().initByName("data", data, "filter", "1-3", "ascertained", True, "excludefrom", 2, "excludeto", 9)


@pytest.mark.gpu
def test_posterior_with_morphological_likelihoods_on_the_species_tree():
    """Canis-FBD.xml wiring: gene-tree likelihoods + four morphological TreeLikelihoods (K = 2..5, LewisMK, ascertained
    FilteredAlignment, one shared strict clock) on tree="@Tree.t:Species" with fossil tips; species-tree moves, store / restore."""
    import oracle
    from starbeast2_b200 import synth
    from starbeast2_b200.plugins import FilteredAlignment, LewisMK, StandardData
    from starbeast2_b200.problem import MorphPartition, lewis_mk_eigen
    pb = synth.make_problem("C4", n_loci=3, n_sites=100)
    rng = np.random.default_rng(21)
    seqs, filt, sp_names = _morph_sequences(pb, rng)
    n_sp = pb.species.n_leaves
    counts = np.bincount(np.concatenate([l.tip_species for l in pb.loci]), minlength=n_sp) // len(pb.loci)
    superset = TaxonSet([TaxonSet(sp_names[i], [Taxon(f"sp{i}_{j}") for j in range(max(1, counts[i]))]) for i in range(n_sp)])
    speciesTree = SpeciesTreeParser()
    speciesTree.initByName("arrays", pb.species, "labels", sp_names, "taxonset", superset)
    geneTrees, likelihoods, trees = [], [], []
    for l, loc in enumerate(pb.loci):
        seen = {}
        labels = []
        for t in range(loc.n_tips):
            s = int(loc.tip_species[t]); j = seen.get(s, 0); seen[s] = j + 1
            labels.append(f"sp{s}_{j}")
        tree = TreeParser()
        tree.initByName("arrays", loc.tree, "labels", labels)
        trees.append(tree)
        gt = GeneTree()
        gt.initByName("tree", tree, "ploidy", loc.ploidy, "speciesTree", speciesTree)
        geneTrees.append(gt)
    for l, loc in enumerate(pb.loci):
        data = Alignment()
        data.initByName("patterns", loc.tip_states, "weights", loc.weights, "taxa", trees[l].labels)
        freqs = Frequencies()
        freqs.initByName("frequencies", RealParameter(value=" ".join(map(str, loc.subst_params[1:5])), dimension=4))
        site = SiteModel()
        site.initByName("substModel", HKY(kappa=RealParameter(value=str(loc.subst_params[0])), frequencies=freqs), "mutationRate", float(loc.cat_rates[0]))
        clock = StrictClockModel()
        clock.initByName("clock.rate", RealParameter(value=repr(float(loc.clock_rate))))
        tl = TreeLikelihood()
        tl.initByName("data", data, "tree", trees[l], "siteModel", site, "branchRateModel", clock)
        likelihoods.append(tl)
    morph = Alignment(sequences=seqs)
    morphRate = RealParameter(value="0.7")
    morphClock = StrictClockModel()
    morphClock.initByName("clock.rate", morphRate)
    morphLikelihoods, parts = [], []
    for K in (2, 3, 4, 5):
        f, lo, hi = filt[K]
        dt = StandardData(nrOfStates=K)
        fa = FilteredAlignment()
        fa.initByName("data", morph, "filter", f, "userDataType", dt, "ascertained", True, "excludefrom", lo, "excludeto", hi)
        site = SiteModel()
        site.initByName("substModel", LewisMK(datatype=dt))
        tl = TreeLikelihood()
        tl.initByName("data", fa, "tree", speciesTree, "siteModel", site, "branchRateModel", morphClock)
        morphLikelihoods.append(tl)
        states, w = fa.patterns_for(sp_names)
        parts.append(MorphPartition(tip_states=states, weights=w, n_states=K, subst_params=lewis_mk_eigen(K), clock_rate=0.7,
                                    excluded=fa.excluded_patterns()))
    speciescoalescent = MultispeciesCoalescent()
    speciescoalescent.initByName("distribution", geneTrees, "populationShape", RealParameter(value=str(pb.alpha)),
                                 "populationMean", RealParameter(value=str(pb.beta / (pb.alpha - 1.0))))
    likelihood = CompoundDistribution()
    likelihood.initByName("distribution", likelihoods + morphLikelihoods)
    posterior = CompoundDistribution()
    posterior.initByName("distribution", [speciescoalescent, likelihood])

    def want_total():
        pb.morph = parts
        return oracle.posterior(pb)

    w = want_total()
    assert posterior.calculateLogP() == pytest.approx(w.total, rel=1e-9)
    assert [tl.logP for tl in morphLikelihoods] == pytest.approx([oracle.tree_likelihood_k(m, pb.species) for m in parts], rel=1e-10)
    assert len(morphLikelihoods[0].getPatternLogLikelihoods()) == parts[0].n_pat
    # a species-tree height move: the coalescent AND the morphological likelihoods follow; reject brings everything back
    before = posterior.logP
    posterior.store(); speciesTree.store()
    sp = speciesTree.arrays
    node = next(i for i in range(sp.n_leaves, sp.n_nodes) if sp.parent[i] >= 0 and sp.height[sp.parent[i]] - sp.height[i] > 1e-4)
    speciesTree.setHeight(node, sp.height[node] + 0.3 * (sp.height[sp.parent[node]] - sp.height[node]))
    pb.species = speciesTree.arrays
    proposed = posterior.calculateLogP()
    w2 = want_total()
    if math.isfinite(w2.coalescent):
        assert proposed == pytest.approx(w2.total, rel=1e-9)
        assert [tl.logP for tl in morphLikelihoods] == pytest.approx(list(w2.per_locus_lik[3:]), rel=1e-10)
        assert morphLikelihoods[0].logP != pytest.approx(w.per_locus_lik[3], rel=1e-12)
    else:
        assert proposed == -math.inf
    speciesTree.restore(); posterior.restore()
    pb.species = speciesTree.arrays
    assert posterior.calculateLogP() == before
    # the morphological clock rate moves: only the four partitions change
    posterior.store()
    morphRate.setValue(0, 1.1)
    for m in parts:
        m.clock_rate = 1.1
    w3 = want_total()
    assert posterior.calculateLogP() == pytest.approx(w3.total, rel=1e-9)
    assert [tl.logP for tl in likelihoods] == pytest.approx(list(w.per_locus_lik[:3]), rel=1e-12)
