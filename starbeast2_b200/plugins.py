"""Host-side mirror of the reference's plugin classes for the hot path -- the drop-in boundary, in Python.

Same class names, Input names and calling protocol as the reference's BEASTObjects
(/root/reference/src/starbeast2/{GeneTree, MultispeciesCoalescent, ConstantPopulations, LinearWithConstantRoot,
PassthroughModel, DummyModel, StarBeastClock, UncorrelatedRates}.java) and as the BEAST.base classes the
StarBEAST2 XML wires next to them (TreeLikelihood, SiteModel, HKY, GTR, Frequencies, StrictClockModel,
Alignment; fxtemplates/StarBeast2.xml:144-162).  None of these classes does likelihood arithmetic: they stage
their state into one process-wide GpuSession, which batches ALL registered loci into libsb2gpu evaluations:

  * the first calculateLogP() of the coalescent phase (MultispeciesCoalescent or any GeneTree) triggers one
    sb2_eval_coalescent for every locus; later calls of the phase return cached per-locus values;
  * the first TreeLikelihood.calculateLogP() triggers one sb2_eval_likelihood for every dirty locus;
  * an incompatible gene tree makes the coalescent -inf, the posterior CompoundDistribution returns early and no
    TreeLikelihood is asked for its value (the reference's short-circuit).  By default the engine has
    by then evaluated the likelihood speculatively in the same device pass (one host round trip per step); with
    SB2_NO_SPECULATION=1 the pruning kernel is not even launched.

The Java twins of these classes (java/ in this repo, uncompiled: no JDK in the image) bind the same C ABI through JNI.
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional, Sequence

import numpy as np

from .beast import (BEASTObject, BooleanParameter, CalculationNode, CompoundDistribution, Distribution, IntegerParameter, RealParameter,
                    Taxon, TaxonSet, Tree, TreeParser)
from .engine import Engine
from .problem import (CLOCK_SPECIES_RATES, CLOCK_STRICT, POP_ANALYTIC, POP_CONSTANT, POP_LWCR, SUBST_EIGEN, SUBST_HKY,
                      gtr_eigen_params, hky_params, lewis_mk_eigen)

REQ = BEASTObject.REQUIRED
__all__ = ["SpeciesTree", "SpeciesTreeParser", "GpuSession", "ConstantPopulations", "LinearWithConstantRoot",
           "UniformPopulations", "RandomLocalRates", "BooleanParameter", "PassthroughModel", "DummyModel", "GeneTree", "MultispeciesCoalescent", "StrictClockModel", "UncorrelatedRates",
           "StarBeastClock", "Alignment", "FilteredAlignment", "StandardData", "LewisMK", "Frequencies", "HKY", "GTR", "SiteModel", "TreeLikelihood", "RealParameter",
           "IntegerParameter", "Taxon", "TaxonSet", "Tree", "TreeParser", "CompoundDistribution"]


# --------------------------------------------------------------------------------------------------------------------
# species tree + taxon maps
# --------------------------------------------------------------------------------------------------------------------

class SpeciesTree(Tree):
    """starbeast2.SpeciesTree / SpeciesTreeParser: a Tree whose taxonset is the species superset; makeMaps()
    (SpeciesTreeInterface.java:25-58) gives gene-tip name -> species leaf node number."""

    def initAndValidate(self):
        super().initAndValidate()
        self.tipNumberMap: Dict[str, int] = {}
        taxa = self.get("taxonset")
        if taxa is not None:
            number = {name: i for i, name in enumerate(self.labels)}
            for sp in taxa.taxa:
                for tip in getattr(sp, "taxa", []):
                    self.tipNumberMap[tip.id] = number.get(sp.id, 0)

    def getTipNumberMap(self):
        return self.tipNumberMap


SpeciesTreeParser = SpeciesTree


# --------------------------------------------------------------------------------------------------------------------
# the session: batching shim between per-object BEAST calls and the batched engine
# --------------------------------------------------------------------------------------------------------------------

class GpuSession:
    """One per species tree.  Not thread-safe: one caller, like the reference's single-threaded MCMC."""
    _by_species: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()
    default_device = 0
    default_exact_order = False

    @classmethod
    def of(cls, species_tree: SpeciesTree) -> "GpuSession":
        s = cls._by_species.get(species_tree)
        if s is None:
            s = cls(species_tree)
            cls._by_species[species_tree] = s
        return s

    def __init__(self, species_tree: SpeciesTree):
        self.species_tree = species_tree
        self.gene_trees: List["GeneTree"] = []
        self.likelihoods: Dict[int, "TreeLikelihood"] = {}     # locus -> TreeLikelihood
        self.morph: List["TreeLikelihood"] = []                # TreeLikelihoods on the species tree itself (K-state, likelihood only)
        self.clocks: Dict[int, "StarBeastClock"] = {}          # locus -> StarBeastClock without a TreeLikelihood
        self.msc: Optional["MultispeciesCoalescent"] = None
        self.engine: Optional[Engine] = None
        self._pushed: Dict[object, object] = {}                # what the engine currently holds, per input
        self._coal_valid = self._lik_valid = False
        self._coal = (0.0, None, None)
        self._lik = (0.0, None)

    # -- registration --------------------------------------------------------------------------------------------
    def register_gene_tree(self, gt: "GeneTree") -> int:
        self.gene_trees.append(gt)
        self._drop_engine()
        return len(self.gene_trees) - 1

    def locus_of_tree(self, tree: Tree) -> Optional[int]:
        for i, gt in enumerate(self.gene_trees):
            if gt.get("tree") is tree:
                return i
        return None

    def _drop_engine(self):
        if self.engine is not None:
            self.engine.close()
        self.engine = None
        self._pushed.clear()
        self._coal_valid = self._lik_valid = False

    # -- build ------------------------------------------------------------------------------------------------------
    def register_morph(self, tl: "TreeLikelihood") -> int:
        self.morph.append(tl)
        self._drop_engine()
        return len(self.morph) - 1

    def _n_cat(self) -> int:
        cats = {tl.get("siteModel").getCategoryCount() for tl in list(self.likelihoods.values()) + self.morph}
        if len(cats) > 1:
            raise ValueError("all TreeLikelihoods of one session must share the rate-category count")
        return cats.pop() if cats else 1

    def _build(self):
        sp = self.species_tree
        e = Engine(device=self.default_device, exact_order=self.default_exact_order)
        n_cat = self._n_cat()
        tip_map = sp.getTipNumberMap()
        for i, gt in enumerate(self.gene_trees):
            tree: Tree = gt.get("tree")
            tip_species = np.array([tip_map[name] for name in tree.labels], np.int32)   # GeneTree.java:216-223
            tl = self.likelihoods.get(i)
            if tl is not None:
                states, weights = tl.get("data").patterns_for(tree.labels)
            else:   # a GeneTree with no TreeLikelihood: one all-missing pattern, log-likelihood exactly 0
                states, weights = np.full((len(tree.labels), 1), 4, np.uint8), np.ones(1, np.int32)
            e.add_locus(states, weights, n_cat, tip_species, gt.getPloidy())
        # TreeLikelihoods whose tree IS the species tree (Canis-FBD.xml:964-1033): K-state, likelihood-only loci after the gene loci,
        # tips in the species tree's own leaf numbering
        for tl in self.morph:
            data = tl.get("data")
            states, weights = data.patterns_for(sp.labels)
            e.add_locus_k(states, weights, data.getStateCount(), n_cat, data.excluded_patterns())
        e.finalize(sp.getNodeCount(), sp.getLeafNodeCount())
        self.engine = e
        self._pushed.clear()

    # -- staging ----------------------------------------------------------------------------------------------------
    def _inputs(self):
        """(key, signature, push) for everything the engine takes from the host.  Signatures are cheap value digests
        (byte strings of small arrays): what the Java twins decide from BEAST's dirty flags in requiresRecalculation()."""
        e, sp = self.engine, self.species_tree
        yield "species_tree", sp._version, (lambda: e.set_species_tree(sp.arrays))
        pop = self.gene_trees[0].popModel if self.gene_trees else None
        if pop is not None and not isinstance(pop, DummyModel):
            yield "pop", pop.signature(), (lambda: pop.push(e))
        elif self.msc is not None and not self.msc.dontCalculate:
            a, b = self.msc.hyperparameters()
            yield "pop", (POP_ANALYTIC, a, b), (lambda: e.set_pop_model(POP_ANALYTIC, alpha=a, beta=b))
        else:   # no model and no analytic integration: GeneTree logP is 0 (GeneTree.java:182)
            yield "pop", ("none",), (lambda: e.set_pop_model(POP_ANALYTIC, alpha=1.0, beta=1.0))
        for i in range(len(self.gene_trees)):
            clk = self._clock(i)
            if isinstance(clk, StarBeastClock):
                r = clk.get("speciesTreeRates").getRatesArray()
                yield "species_rates", r.tobytes(), (lambda r=r: e.set_species_rates(r))
                break
        for i, gt in enumerate(self.gene_trees):
            tree: Tree = gt.get("tree")
            yield ("tree", i), tree._version, (lambda i=i, tree=tree: e.set_gene_tree(i, tree.arrays))
            tl = self.likelihoods.get(i)
            if tl is not None:
                sm: SiteModel = tl.get("siteModel")

                def push_site(i=i, sm=sm):
                    kind, params = sm.get("substModel").engine_params()
                    e.set_subst_model(i, kind, params)
                    rates, props = sm.category_rates()
                    e.set_site_model(i, rates, props, sm.getProportionInvariant())
                yield ("site", i), sm.signature(), push_site
            clk = self._clock(i)
            if clk is not None:
                yield ("clock", i), clk.signature(), (lambda i=i, clk=clk: e.set_clock(i, *clk.engine_clock()))

        L = len(self.gene_trees)
        for j, tl in enumerate(self.morph):
            i = L + j
            yield ("tree", i), sp._version, (lambda i=i: e.set_gene_tree(i, sp.arrays))
            sm = tl.get("siteModel")

            def push_msite(i=i, sm=sm):
                kind, params = sm.get("substModel").engine_params()
                e.set_subst_model(i, kind, params)
                rates, props = sm.category_rates()
                e.set_site_model(i, rates, props, sm.getProportionInvariant())
            yield ("site", i), sm.signature(), push_msite
            clk = tl.get("branchRateModel")
            yield ("clock", i), clk.signature(), (lambda i=i, clk=clk: e.set_clock(i, *clk.engine_clock()))

    def _clock(self, i):
        tl = self.likelihoods.get(i)
        return tl.get("branchRateModel") if tl is not None else self.clocks.get(i)

    def _sync(self):
        """Stage every input whose value differs from what the engine holds."""
        if self.engine is None:
            self._build()
        dirty = False
        for key, sig, push in self._inputs():
            if self._pushed.get(key, None) != sig or key not in self._pushed:
                push()
                self._pushed[key] = sig
                dirty = True
        if dirty:
            self._coal_valid = self._lik_valid = False

    # -- evaluation -------------------------------------------------------------------------------------------------
    def coalescent(self):
        self._sync()
        if not self._coal_valid:
            self._coal = self.engine.eval_coalescent()
            self._coal_valid = True
        return self._coal

    def likelihood(self):
        self._sync()
        if not self._lik_valid:
            self._lik = self.engine.eval_likelihood()
            self._lik_valid = True
        return self._lik

    def branch_rates(self, locus: int) -> np.ndarray:
        self.coalescent()   # k_prepare computes the rates in the coalescent phase
        return self.engine.branch_rates(locus)

    # -- MCMC protocol ----------------------------------------------------------------------------------------------
    def store(self):
        if self.engine is not None:
            self._sync()
            self.engine.store()

    def restore(self):
        """Call AFTER the host state nodes have been restored (BEAST: state.restore() precedes
        restoreCalculationNodes()): the engine swaps back to its stored block, which holds exactly those values."""
        if self.engine is not None:
            self.engine.restore()
            for key, sig, _ in self._inputs():
                self._pushed[key] = sig
        self._coal_valid = self._lik_valid = False

    def accept(self):
        if self.engine is not None:
            self.engine.accept()


# --------------------------------------------------------------------------------------------------------------------
# population models
# --------------------------------------------------------------------------------------------------------------------

class PopulationModel(CalculationNode):
    def getBaseModel(self):
        return self


class ConstantPopulations(PopulationModel):
    """starbeast2.ConstantPopulations (ConstantPopulations.java:17-42)."""
    INPUTS = {"speciesTree": REQ, "populationSizes": REQ}

    def initAndValidate(self):
        self.speciesTree = self.get("speciesTree")
        self.get("populationSizes").setDimension(self.speciesTree.getNodeCount())   # :37

    def initPopSizes(self, popInitial: float):   # :52-62
        p = self.get("populationSizes")
        if p.get("estimate") and p.getLower() < popInitial < p.getUpper():
            for i in range(p.getDimension()):
                p.setValue(i, popInitial)

    def signature(self):
        return (POP_CONSTANT, self.get("populationSizes").values.tobytes())

    def push(self, e: Engine):
        e.set_pop_model(POP_CONSTANT, self.get("populationSizes").values)


class UniformPopulations(PopulationModel):
    """starbeast2.UniformPopulations (UniformPopulations.java:16-18: inputs speciesTree, universalSize): one population size
    for every branch.  uniformLogP (:74-99) is the constant-population formula, so the engine runs its constant model with the
    universal size broadcast over the species nodes."""
    INPUTS = {"speciesTree": REQ, "universalSize": REQ}

    def initAndValidate(self):
        self.speciesTree = self.get("speciesTree")

    def initPopSizes(self, popInitial: float):   # :46-55
        p = self.get("universalSize")
        if p.get("estimate") and p.getLower() < popInitial < p.getUpper():
            p.setValue(0, popInitial)

    def _sizes(self) -> np.ndarray:
        return np.full(self.speciesTree.getNodeCount(), float(self.get("universalSize").getValue()))

    def signature(self):
        return (POP_CONSTANT, self._sizes().tobytes())

    def push(self, e: Engine):
        e.set_pop_model(POP_CONSTANT, self._sizes())


class LinearWithConstantRoot(PopulationModel):
    """starbeast2.LinearWithConstantRoot (LinearWithConstantRoot.java:19-48)."""
    INPUTS = {"speciesTree": REQ, "tipPopulationSizes": REQ, "topPopulationSizes": REQ}

    def initAndValidate(self):
        self.speciesTree = self.get("speciesTree")
        self.get("tipPopulationSizes").setDimension(self.speciesTree.getLeafNodeCount())   # :43
        self.get("topPopulationSizes").setDimension(self.speciesTree.getNodeCount() - 1)   # :44

    def initPopSizes(self, tipInitial: float):   # :73-94
        topInitial = tipInitial * 0.5
        tip, top = self.get("tipPopulationSizes"), self.get("topPopulationSizes")
        if tip.get("estimate") and tip.getLower() < tipInitial < tip.getUpper():
            for i in range(tip.getDimension()):
                tip.setValue(i, tipInitial)
        if top.get("estimate") and top.getLower() < topInitial < top.getUpper():
            for i in range(top.getDimension()):
                top.setValue(i, topInitial)

    def signature(self):
        return (POP_LWCR, self.get("tipPopulationSizes").values.tobytes(), self.get("topPopulationSizes").values.tobytes())

    def push(self, e: Engine):
        e.set_pop_model(POP_LWCR, self.get("tipPopulationSizes").values, self.get("topPopulationSizes").values)


class PassthroughModel(PopulationModel):
    """starbeast2.PassthroughModel: BEAUti indirection (PassthroughModel.java:54-56)."""
    INPUTS = {"childModel": REQ}

    def getBaseModel(self):
        return self.get("childModel").getBaseModel()


class DummyModel(PopulationModel):
    """starbeast2.DummyModel: 'use analytic integration' marker (DummyModel.java:19-21, GeneTree.java:182)."""
    INPUTS = {"speciesTree": None}


# --------------------------------------------------------------------------------------------------------------------
# GeneTree / MultispeciesCoalescent
# --------------------------------------------------------------------------------------------------------------------

class GeneTree(Distribution):
    """starbeast2.GeneTree (GeneTree.java:21-25: inputs tree, speciesTree, ploidy = 2.0, populationModel)."""
    INPUTS = {"tree": REQ, "speciesTree": REQ, "ploidy": 2.0, "populationModel": None}

    def initAndValidate(self):
        pm = self.get("populationModel")
        self.popModel = pm.getBaseModel() if pm is not None else None   # GeneTree.java:126
        self.ploidy = float(self.get("ploidy"))
        self.session = GpuSession.of(self.get("speciesTree"))
        self.locus = self.session.register_gene_tree(self)
        self.logP = 0.0

    def requiresRecalculation(self):   # GeneTree.java:73-76
        return True

    def calculateLogP(self) -> float:   # GeneTree.java:171-200
        _, per_locus, _ = self.session.coalescent()
        self.logP = float(per_locus[self.locus])
        return self.logP

    def getPloidy(self):
        return self.ploidy

    def getNodeCount(self):
        return self.get("tree").getNodeCount()

    def getRoot(self):
        return self.get("tree").getRootNr()

    def getTipNumberMap(self):
        m = self.get("speciesTree").getTipNumberMap()
        return [m[name] for name in self.get("tree").labels]

    # pulled from the device only when somebody asks (NodeReheight2 / loggers / StarBeastClock on the Java side)
    def coalescentLineageCounts(self):
        self.session.coalescent()
        return self.session.engine.embedding(self.locus)[0]

    def coalescentCounts(self):
        self.session.coalescent()
        return self.session.engine.embedding(self.locus)[1]


class MultispeciesCoalescent(CompoundDistribution):
    """starbeast2.MultispeciesCoalescent (MultispeciesCoalescent.java:20-22)."""
    INPUTS = {"populationShape": None, "populationMean": None}

    def initAndValidate(self):
        shape, mean = self.get("populationShape"), self.get("populationMean")
        if (shape is None) != (mean is None):   # :93-96
            raise ValueError("Either specify both population size prior parameters for analytical integration,"
                             "or neither for MCMC integration of population sizes.")
        self.alpha = self.beta = 0.0
        self.storedAlpha = self.storedBeta = 0.0
        self.dontCalculate = shape is None
        for d in self.get("distribution"):
            if not isinstance(d, GeneTree):   # :116-118
                raise ValueError("Input distributions must all be of class GeneTree.")
        self.session = self.get("distribution")[0].session if self.get("distribution") else None
        if self.session is not None:
            self.session.msc = self
            self.session._coal_valid = False
        if not self.dontCalculate:
            self._check_hyperparameters(True)

    def _check_hyperparameters(self, force: bool) -> bool:
        """MultispeciesCoalescent.java:134-147 -- including the stale-alpha quirk (Q1): beta uses the PREVIOUS alpha."""
        currentAlpha = float(self.get("populationShape").getValue())
        currentBeta = float(self.get("populationMean").getValue()) * (self.alpha - 1.0)
        if force or currentAlpha != self.alpha or currentBeta != self.beta:
            self.alpha, self.beta = currentAlpha, currentBeta
            return True
        return False

    def hyperparameters(self):
        return self.alpha, self.beta

    def calculateLogP(self) -> float:   # :150-198
        if not self.dontCalculate:
            self._check_hyperparameters(False)
        total, per_locus, _ = self.session.coalescent()
        for d in self.get("distribution"):
            d.logP = float(per_locus[d.locus])
        self.logP = float(total)
        return self.logP

    def store(self):   # MultispeciesCoalescent.java:50-61
        Distribution.store(self)
        if not self.dontCalculate:
            self.storedAlpha, self.storedBeta = self.alpha, self.beta
        self.session.store()

    def restore(self):   # MultispeciesCoalescent.java:63-88: (alpha, beta) swap back BEFORE the session re-derives its signature
        Distribution.restore(self)
        if not self.dontCalculate:
            self.alpha, self.storedAlpha = self.storedAlpha, self.alpha
            self.beta, self.storedBeta = self.storedBeta, self.beta
        self.session.restore()

    def accept(self):
        self.session.accept()


# --------------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------------

class BranchRateModel(CalculationNode):
    INPUTS = {"clock_rate": None}

    def mean_rate(self) -> float:
        r = self.get("clock.rate")
        return 1.0 if r is None else float(r.getValue() if isinstance(r, RealParameter) else r)


class StrictClockModel(BranchRateModel):
    def signature(self):
        return ("strict", self.mean_rate())

    def engine_clock(self):
        return CLOCK_STRICT, self.mean_rate(), None

    def getRateForBranch(self, node):
        return self.mean_rate()


class UncorrelatedRates(BranchRateModel):
    """starbeast2.UncorrelatedRates (UncorrelatedRates.java:16-22): SpeciesTreeRates from discretised log-normal
    (or exponential) bins.  Host work in the reference too; the engine consumes getRatesArray().
    The reference's commons-math 2.x inverse CDF is an iterative solver (quirk Q5); this mirror uses exact quantiles."""
    INPUTS = {"tree": REQ, "nBins": -1, "estimateRoot": False, "noCache": False, "stdev": None, "rates": REQ}

    def initAndValidate(self):   # :81-129
        tree = self.get("tree")
        self.estimateRoot = bool(self.get("estimateRoot"))
        self.rootNodeNumber = tree.getRootNr()
        S = tree.getNodeCount()
        self.nEstimatedRates = S if self.estimateRoot else S - 1
        nb = int(self.get("nBins"))
        self.nBins = self.nEstimatedRates if nb <= 0 else nb
        rates = self.get("rates")
        rates.setDimension(self.nEstimatedRates)
        rates.setLower(0)
        rates.setUpper(self.nBins - 1)
        self.useLogNormal = self.get("stdev") is not None
        self._bins_for = None
        if not self.useLogNormal:
            self.binRates = -np.log(1.0 - (np.arange(self.nBins) + 0.5) / self.nBins)   # Exp(1) quantiles, :113-117

    def _stdev(self):
        s = self.get("stdev")
        return float(s.getValue() if isinstance(s, RealParameter) else s)

    def getRatesArray(self) -> np.ndarray:   # :131-170
        if self.useLogNormal:
            sd = self._stdev()
            if self._bins_for != sd or self.get("noCache"):
                from scipy.special import ndtri
                self.binRates = np.exp(-(0.5 * sd * sd) + sd * ndtri((np.arange(self.nBins) + 0.5) / self.nBins))
                self._bins_for = sd
        mean = self.mean_rate()
        out = np.zeros(self.get("tree").getNodeCount())
        idx = np.asarray(self.get("rates").values[: self.nEstimatedRates], dtype=np.int64)
        out[: self.nEstimatedRates] = self.binRates[idx] * mean
        if not self.estimateRoot:
            out[self.rootNodeNumber] = mean   # :161
        return out

    def getRateForBranch(self, node: int) -> float:
        return float(self.getRatesArray()[node])


class RandomLocalRates(BranchRateModel):
    """starbeast2.RandomLocalRates (RandomLocalRates.java:15-19: inputs tree, noCache, rates, indicators, clock.rate):
    SpeciesTreeRates of a random local clock.  Host work in the reference too; the engine consumes getRatesArray().
    Quirk kept (Q9): the recursion starts at the root with rate 1.0 and stores it (:84, :110), overwriting the mean rate
    written just before (:106), and the normalisation loop stops below the root (:114-116) -- so the root branch carries
    rate 1.0, not clock.rate."""
    INPUTS = {"tree": REQ, "noCache": False, "rates": REQ, "indicators": REQ}

    def initAndValidate(self):   # :49-70
        tree = self.get("tree")
        self.rootNodeNumber = tree.getNodeCount() - 1
        rates, ind = self.get("rates"), self.get("indicators")
        rates.setDimension(self.rootNodeNumber)
        ind.setDimension(self.rootNodeNumber)
        if rates.get("lower") is None or rates.getLower() < 0.0:
            rates.setLower(0.0)
        if rates.get("upper") is None or rates.getUpper() < 0.0:
            rates.setUpper(float(np.finfo(np.float64).max))

    def getRatesArray(self) -> np.ndarray:   # update(), :92-117
        t = self.get("tree").arrays
        ind = np.asarray(self.get("indicators").values, bool)
        br = np.asarray(self.get("rates").values, float)
        mean = self.mean_rate()
        out = np.zeros(t.n_nodes)
        out[self.rootNodeNumber] = mean
        strict = relaxed = 0.0
        stack = [(t.root, float(t.height[t.root]), 1.0)]
        while stack:   # recurseBranchRates, :72-90, pre-order: node, left subtree, right subtree (the sums depend on this order)
            node, parent_h, rate = stack.pop()
            if node < self.rootNodeNumber and ind[node]:
                rate = float(br[node])
            length = parent_h - float(t.height[node])
            strict += length
            relaxed += length * rate
            out[node] = rate
            if t.left[node] >= 0:
                stack.append((int(t.right[node]), float(t.height[node]), rate))
                stack.append((int(t.left[node]), float(t.height[node]), rate))
        scale = mean * strict / relaxed
        out[: self.rootNodeNumber] = out[: self.rootNodeNumber] * scale
        return out

    def getRateForBranch(self, node: int) -> float:
        return float(self.getRatesArray()[node])


class StarBeastClock(BranchRateModel):
    """starbeast2.StarBeastClock (StarBeastClock.java:9-11): gene-branch rate = clock.rate x occupancy-weighted mean of
    the species-branch rates, computed on the device by k_prepare."""
    INPUTS = {"geneTree": REQ, "speciesTreeRates": REQ}

    def initAndValidate(self):
        self.geneTree: GeneTree = self.get("geneTree")
        if not isinstance(self.geneTree, GeneTree):
            raise ValueError("geneTree must be a GeneTree")   # typed Input<GeneTree>, StarBeastClock.java:10
        s = self.geneTree.session
        if self.geneTree.locus not in s.likelihoods:
            s.clocks[self.geneTree.locus] = self
            s._coal_valid = s._lik_valid = False

    def signature(self):
        return ("starbeast", self.mean_rate())

    def engine_clock(self):
        return CLOCK_SPECIES_RATES, self.mean_rate(), None

    def getRateForBranch(self, node: int) -> float:   # StarBeastClock.java:78-84
        return float(self.geneTree.session.branch_rates(self.geneTree.locus)[node])


# --------------------------------------------------------------------------------------------------------------------
# BEAST.base classes the StarBEAST2 XML wires into the likelihood (mirrored from upstream knowledge)
# --------------------------------------------------------------------------------------------------------------------

class Alignment(BEASTObject):
    """beast.base.evolution.alignment.Alignment: sequences -> unique site patterns + weights (calcPatterns)."""
    INPUTS = {"sequences": None, "patterns": None, "weights": None, "taxa": None}
    CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}

    def initAndValidate(self):
        if self.get("patterns") is not None:
            self.taxa = list(self.get("taxa"))
            self.patterns = np.ascontiguousarray(self.get("patterns"), np.uint8)
            self.weights = np.ascontiguousarray(self.get("weights"), np.int32)
            return
        seqs: Dict[str, str] = self.get("sequences")
        self.taxa = sorted(seqs)
        mat = np.array([[self.CODE.get(ch.upper(), 4) for ch in seqs[t]] for t in self.taxa], np.uint8)
        pats, w = np.unique(mat, axis=1, return_counts=True)
        self.patterns, self.weights = np.ascontiguousarray(pats), w.astype(np.int32)

    def patterns_for(self, labels: Sequence[str]):
        idx = [self.taxa.index(l) for l in labels]
        return np.ascontiguousarray(self.patterns[idx]), self.weights

    def getPatternCount(self):
        return self.patterns.shape[1]

    def empirical_frequencies(self):
        counts = np.array([((self.patterns == s) * self.weights[None, :]).sum() for s in range(4)], float)
        return counts / counts.sum()


class StandardData(BEASTObject):
    """beast.base.evolution.datatype.StandardData (userDataType of the morphological alignments, Canis-FBD.xml:973):
    characters '0'..'9' are states 0..9; '?', '-' and anything else are missing (ambiguities="" in the reference's XML)."""
    INPUTS = {"nrOfStates": REQ, "ambiguities": ""}

    def getStateCount(self) -> int:
        return int(self.get("nrOfStates"))

    def encode(self, ch: str) -> int:
        k = self.getStateCount()
        return int(ch) if ch.isdigit() and int(ch) < k else k


def _parse_filter(spec: str, n_sites: int) -> List[int]:
    """FilteredAlignment's filter syntax: 1-based sites, comma separated, 'a-b' ranges, 'a-b\\c' every c-th, '-b', 'a-'."""
    out: List[int] = []
    for part in spec.split(","):
        part = part.strip()
        if not part:
            continue
        step = 1
        if "\\" in part:
            part, st = part.split("\\")
            step = int(st)
        if "-" in part:
            a, b = part.split("-")
            lo = int(a) if a.strip() else 1
            hi = int(b) if b.strip() else n_sites
            out.extend(range(lo - 1, hi, step))
        else:
            out.append(int(part) - 1)
    return out


class FilteredAlignment(BEASTObject):
    """beast.base.evolution.alignment.FilteredAlignment with the ascertainment inputs of Alignment: a subset of the sites of
    `data` (filter), re-typed by userDataType, optionally ascertained -- sites [excludefrom, excludeto) of the FILTERED
    alignment are dummy columns whose patterns get weight 0 and form Alignment.excludedPatterns (Canis-FBD.xml:966-974)."""
    INPUTS = {"data": REQ, "filter": REQ, "userDataType": None, "ascertained": False, "excludefrom": 0, "excludeto": 0, "excludeevery": 1}

    def initAndValidate(self):
        src = self.get("data")
        seqs: Dict[str, str] = src.get("sequences")
        if seqs is None:
            raise ValueError("FilteredAlignment needs an Alignment built from sequences")
        dt = self.get("userDataType")
        self.taxa = sorted(seqs)
        n_sites = len(seqs[self.taxa[0]])
        sites = _parse_filter(str(self.get("filter")), n_sites)
        enc = dt.encode if dt is not None else (lambda ch: Alignment.CODE.get(ch.upper(), 4))
        self.n_states = dt.getStateCount() if dt is not None else 4
        mat = np.array([[enc(seqs[t][s]) for s in sites] for t in self.taxa], np.uint8)
        pats, index, w = np.unique(mat, axis=1, return_inverse=True, return_counts=True)
        index = np.asarray(index).reshape(-1)
        self.patterns, self.weights = np.ascontiguousarray(pats), w.astype(np.int32)
        self.excluded = np.zeros(0, np.int32)
        if self.get("ascertained"):
            lo, hi, every = int(self.get("excludefrom")), int(self.get("excludeto")), int(self.get("excludeevery"))
            if not (0 <= lo <= hi <= len(sites)):
                raise ValueError("excludefrom / excludeto outside the filtered alignment")
            ex = sorted({int(index[i]) for i in range(lo, hi, every)})
            self.weights[ex] = 0                       # "reduce weight, so it does not confuse the tree likelihood"
            self.excluded = np.array(ex, np.int32)

    def patterns_for(self, labels: Sequence[str]):
        idx = [self.taxa.index(l) for l in labels]
        return np.ascontiguousarray(self.patterns[idx]), self.weights

    def getPatternCount(self):
        return self.patterns.shape[1]

    def getStateCount(self) -> int:
        return self.n_states

    def excluded_patterns(self) -> np.ndarray:
        return self.excluded


class LewisMK(BEASTObject):
    """morphmodels.evolution.substitutionmodel.LewisMK (Canis-FBD.xml:975): Lewis's Mk -- equal rates between all K states,
    equal frequencies.  The engine takes its eigen system, as it does for every GeneralSubstitutionModel."""
    INPUTS = {"datatype": None, "stateNumber": None}

    def getStateCount(self) -> int:
        dt = self.get("datatype")
        return dt.getStateCount() if dt is not None else int(self.get("stateNumber"))

    def engine_params(self):
        return SUBST_EIGEN, lewis_mk_eigen(self.getStateCount())


class Frequencies(BEASTObject):
    INPUTS = {"frequencies": None, "data": None, "estimate": True}

    def getFreqs(self) -> np.ndarray:
        f = self.get("frequencies")
        if f is not None:
            return np.asarray(f.values if isinstance(f, RealParameter) else f, float)
        return self.get("data").empirical_frequencies()


def _val(x, default=None):
    if x is None:
        return default
    return float(x.getValue() if isinstance(x, RealParameter) else x)


class HKY(BEASTObject):
    INPUTS = {"kappa": REQ, "frequencies": REQ}

    def engine_params(self):
        return SUBST_HKY, hky_params(_val(self.get("kappa")), self.get("frequencies").getFreqs())


class GTR(BEASTObject):
    INPUTS = {"rateAC": 1.0, "rateAG": 1.0, "rateAT": 1.0, "rateCG": 1.0, "rateCT": 1.0, "rateGT": 1.0, "frequencies": REQ}

    def engine_params(self):
        r = [_val(self.get(k)) for k in ("rateAC", "rateAG", "rateAT", "rateCG", "rateCT", "rateGT")]
        return SUBST_EIGEN, gtr_eigen_params(r, self.get("frequencies").getFreqs())


class SiteModel(CalculationNode):
    """beast.base.evolution.sitemodel.SiteModel: gamma category rates as category medians normalised to mean 1, times
    mutationRate (the Java plugin takes getRateForCategory / getCategoryProportions from the host object, quirk Q5)."""
    INPUTS = {"substModel": REQ, "mutationRate": 1.0, "gammaCategoryCount": 0, "shape": 1.0, "proportionInvariant": 0.0}

    def getCategoryCount(self) -> int:
        return max(1, int(self.get("gammaCategoryCount")))

    def getProportionInvariant(self) -> float:
        return _val(self.get("proportionInvariant"), 0.0)

    def category_rates(self):
        k, mu, pinv = self.getCategoryCount(), _val(self.get("mutationRate"), 1.0), self.getProportionInvariant()
        prop_variable = 1.0 - pinv
        if k == 1:
            rates = np.ones(1)
        else:
            from scipy.stats import gamma
            a = _val(self.get("shape"), 1.0)
            rates = gamma.ppf((2.0 * np.arange(k) + 1.0) / (2.0 * k), a, scale=1.0 / a)
        mean = prop_variable * rates.sum() / k
        return rates / mean * mu, np.full(k, prop_variable / k)

    def signature(self):
        kind, p = self.get("substModel").engine_params()
        r, pr = self.category_rates()
        return (kind, p.tobytes(), r.tobytes(), pr.tobytes(), self.getProportionInvariant())


class TreeLikelihood(Distribution):
    """beast.base.evolution.likelihood.TreeLikelihood drop-in: inputs data, tree, siteModel, branchRateModel
    (fxtemplates/StarBeast2.xml:158-163)."""
    INPUTS = {"data": REQ, "tree": REQ, "siteModel": REQ, "branchRateModel": None}

    def initAndValidate(self):
        tree = self.get("tree")
        clk = self.get("branchRateModel")
        if clk is None:
            clk = StrictClockModel()
            clk.initByName()
            self._inputs["branchRateModel"] = clk
        if isinstance(tree, SpeciesTree):
            # tree="@Tree.t:Species": a likelihood on the species tree itself (the morphological partitions of the FBD analyses,
            # Canis-FBD.xml:964-1033) -- no GeneTree, no coalescent term; K-state, likelihood-only locus of the species tree's session
            if not hasattr(self.get("data"), "excluded_patterns"):
                raise ValueError("a TreeLikelihood on the species tree takes a FilteredAlignment (K-state data)")
            if isinstance(clk, StarBeastClock):
                raise ValueError("StarBeastClock needs a gene tree")
            self.session = GpuSession.of(tree)
            self.locus = None
            self.morph_index = self.session.register_morph(self)
            return
        session = None
        if isinstance(clk, StarBeastClock):
            session = clk.geneTree.session
        else:
            for s in list(GpuSession._by_species.values()):
                if s.locus_of_tree(tree) is not None:
                    session = s
        if session is None:
            raise ValueError("TreeLikelihood's tree is not embedded by any starbeast2.GeneTree (declare the GeneTree first)")
        self.session = session
        self.locus = session.locus_of_tree(tree)
        session.likelihoods[self.locus] = self
        session.clocks.pop(self.locus, None)
        session._drop_engine()   # the locus now has data: rebuild on next evaluation

    def calculateLogP(self) -> float:
        _, per_locus = self.session.likelihood()
        self.logP = float(per_locus[self._index()])
        return self.logP

    def _index(self) -> int:
        return self.locus if self.locus is not None else len(self.session.gene_trees) + self.morph_index

    def getPatternLogLikelihoods(self):
        self.session.likelihood()
        return self.session.engine.pattern_logl(self._index())
