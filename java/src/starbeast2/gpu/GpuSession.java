package starbeast2.gpu;

import java.util.ArrayList;
import java.util.IdentityHashMap;
import java.util.List;
import java.util.Map;

import beast.base.evolution.alignment.Alignment;
import beast.base.evolution.branchratemodel.BranchRateModel;
import beast.base.evolution.branchratemodel.StrictClockModel;
import beast.base.evolution.sitemodel.SiteModel;
import beast.base.evolution.substitutionmodel.EigenDecomposition;
import beast.base.evolution.substitutionmodel.GeneralSubstitutionModel;
import beast.base.evolution.substitutionmodel.HKY;
import beast.base.evolution.substitutionmodel.SubstitutionModel;
import beast.base.evolution.tree.Node;
import beast.base.evolution.tree.TreeInterface;
import beast.base.inference.parameter.RealParameter;
import starbeast2.ConstantPopulations;
import starbeast2.DummyModel;
import starbeast2.GpuGeneTree;
import starbeast2.GpuMultispeciesCoalescent;
import starbeast2.LinearWithConstantRoot;
import starbeast2.UniformPopulations;
import starbeast2.PopulationModel;
import starbeast2.SpeciesTreeInterface;
import starbeast2.SpeciesTreeRates;
import starbeast2.StarBeastClock;

/**
 * Batching shim between BEAST's one-Distribution-at-a-time calls and libsb2gpu's all-loci-at-once evaluations
 * (SURVEY.md 8b).  One session per species tree; every GpuGeneTree / GpuTreeLikelihood registers at initAndValidate.
 * The first calculateLogP() of the coalescent phase evaluates the coalescent of ALL loci, the first TreeLikelihood
 * call evaluates ALL dirty likelihoods; later calls of the same phase return cached per-locus values.  An
 * incompatible gene tree makes the coalescent -infinity, the posterior CompoundDistribution returns early and the
 * likelihood phase never launches.  Single caller, like the reference's MCMC (tutorial/StarBEAST2-tutorial.tex:419).
 *
 * UNCOMPILED in the build image (no JDK); the tested twin is starbeast2_b200/plugins.py:GpuSession.
 */
public final class GpuSession {
    private static final Map<SpeciesTreeInterface, GpuSession> SESSIONS = new IdentityHashMap<>();

    public static synchronized GpuSession of(SpeciesTreeInterface speciesTree) {
        return SESSIONS.computeIfAbsent(speciesTree, GpuSession::new);
    }

    /** The session whose GpuGeneTree embeds this gene tree (strict-clock likelihoods have no other link to it). */
    public static synchronized GpuSession ofGeneTree(TreeInterface tree) {
        for (GpuSession s : SESSIONS.values()) if (s.locusOf(tree) >= 0) return s;
        return null;
    }

    private final SpeciesTreeInterface speciesTree;
    private final List<GpuGeneTree> geneTrees = new ArrayList<>();
    private final Map<Integer, GpuTreeLikelihood> likelihoods = new java.util.HashMap<>();
    private final List<GpuTreeLikelihood> morph = new ArrayList<>();   // TreeLikelihoods on the species tree itself (K-state, no GeneTree)
    private GpuMultispeciesCoalescent msc;
    private long engine = 0;
    private boolean coalValid = false, likValid = false, firstPush = true;
    private double[] perLocusCoal, perLocusLik, perBranch;
    private final double[] total = new double[1];
    private double coalTotal, likTotal;

    private GpuSession(SpeciesTreeInterface speciesTree) { this.speciesTree = speciesTree; }

    public int register(GpuGeneTree gt) { geneTrees.add(gt); close(); return geneTrees.size() - 1; }
    public void register(int locus, GpuTreeLikelihood tl) { likelihoods.put(locus, tl); close(); }
    public int registerMorph(GpuTreeLikelihood tl) { morph.add(tl); close(); return morph.size() - 1; }
    public double morphLikelihood(int index) { evalLikelihood(); return perLocusLik[geneTrees.size() + index]; }
    public void setMsc(GpuMultispeciesCoalescent m) { msc = m; coalValid = false; }
    public int locusOf(TreeInterface tree) {
        for (int i = 0; i < geneTrees.size(); i++) if (geneTrees.get(i).treeInput.get() == tree) return i;
        return -1;
    }

    private void check(int rc) { if (rc < 0) throw new IllegalStateException(Sb2Gpu.lastError(engine)); }
    private void close() { if (engine != 0) Sb2Gpu.destroy(engine); engine = 0; firstPush = true; coalValid = likValid = false; }

    private void build() {
        engine = Sb2Gpu.create(Integer.getInteger("sb2gpu.device", 0), Boolean.getBoolean("sb2gpu.exactOrder"));
        int nCat = 1;
        for (GpuTreeLikelihood tl : likelihoods.values()) nCat = ((SiteModel) tl.siteModelInput.get()).getCategoryCount();
        for (int l = 0; l < geneTrees.size(); l++) {
            final GpuGeneTree gt = geneTrees.get(l);
            final TreeInterface tree = gt.treeInput.get();
            final int nTips = tree.getLeafNodeCount();
            final int[] tipSpecies = gt.getTipNumberMap();            // species LEAF NODE NUMBER per gene tip (GeneTree.java:216-223)
            final GpuTreeLikelihood tl = likelihoods.get(l);
            byte[] states; int[] weights; int nPat;
            if (tl != null) {
                final Alignment data = tl.dataInput.get();
                nPat = data.getPatternCount();
                states = new byte[nTips * nPat];
                weights = data.getWeights();
                for (int t = 0; t < nTips; t++) {
                    final int taxon = data.getTaxonIndex(tree.getNode(t).getID());
                    for (int p = 0; p < nPat; p++) {
                        final int code = data.getPattern(taxon, p);
                        states[t * nPat + p] = (byte) (data.getDataType().isAmbiguousCode(code) || code > 3 ? 4 : code);   // useAmbiguities=false
                    }
                }
            } else {   // a GeneTree without likelihood: one all-missing pattern, log-likelihood exactly 0
                nPat = 1; states = new byte[nTips]; java.util.Arrays.fill(states, (byte) 4); weights = new int[]{1};
            }
            check(Sb2Gpu.addLocus(engine, nTips, nPat, states, weights, nCat, tipSpecies, gt.getPloidy()));
        }
        // likelihoods on the species tree itself: K-state, likelihood-only loci after the gene loci (sb2_add_locus_k, tipToSpecies = null);
        // tips in the species tree's own leaf numbering; the ascertained dummy patterns are Alignment's excludedPatterns
        for (GpuTreeLikelihood tl : morph) {
            final Alignment data = tl.dataInput.get();
            final TreeInterface tree = (TreeInterface) speciesTree;
            final int nTips = tree.getLeafNodeCount(), nPat = data.getPatternCount(), K = data.getDataType().getStateCount();
            final byte[] states = new byte[nTips * nPat];
            for (int t = 0; t < nTips; t++) {
                final int taxon = data.getTaxonIndex(tree.getNode(t).getID());
                for (int p = 0; p < nPat; p++) {
                    final int code = data.getPattern(taxon, p);
                    states[t * nPat + p] = (byte) (data.getDataType().isAmbiguousCode(code) || code >= K ? K : code);   // useAmbiguities=false
                }
            }
            final int[] excluded = data.getExcludedPatternIndices().stream().mapToInt(Integer::intValue).toArray();
            check(Sb2Gpu.addLocusK(engine, nTips, nPat, K, states, data.getWeights(), nCat, null, 2.0, excluded.length == 0 ? null : excluded));
        }
        final int S = speciesTree.getNodeCount();
        check(Sb2Gpu.finalizeEngine(engine, S, speciesTree.getLeafNodeCount()));
        final int nLoci = geneTrees.size() + morph.size();
        perLocusCoal = new double[nLoci]; perLocusLik = new double[nLoci]; perBranch = new double[S];
        firstPush = true;
    }

    private static void pushTree(long e, int locus, TreeInterface tree) {
        final int n = tree.getNodeCount();
        final int[] parent = new int[n], left = new int[n], right = new int[n];
        final double[] height = new double[n];
        for (Node node : tree.getNodesAsArray()) {
            final int i = node.getNr();
            parent[i] = node.isRoot() ? -1 : node.getParent().getNr();
            left[i] = node.isLeaf() ? -1 : node.getLeft().getNr();
            right[i] = node.isLeaf() ? -1 : node.getRight().getNr();
            height[i] = node.getHeight();
        }
        final int rc = locus < 0 ? Sb2Gpu.setSpeciesTree(e, parent, left, right, height) : Sb2Gpu.setGeneTree(e, locus, parent, left, right, height);
        if (rc < 0) throw new IllegalStateException(Sb2Gpu.lastError(e));
    }

    /** Stages whatever BEAST flagged dirty (called from the plugins' requiresRecalculation / calculateLogP). */
    private void sync() {
        if (engine == 0) build();
        boolean dirty = firstPush;
        if (firstPush || speciesTree.getCurrent().somethingIsDirty()) { pushTree(engine, -1, (TreeInterface) speciesTree); dirty = true; }
        // population model (one per analysis; PassthroughModel already unwrapped by GeneTree.initAndValidate, GeneTree.java:126)
        final PopulationModel pop = geneTrees.get(0).baseModel();
        if (pop instanceof ConstantPopulations) {
            final RealParameter n = ((ConstantPopulations) pop).popSizesInput.get();
            if (firstPush || n.somethingIsDirty()) { check(Sb2Gpu.setPopModel(engine, Sb2Gpu.POP_CONSTANT, n.getDoubleValues(), null, 0, 0)); dirty = true; }
        } else if (pop instanceof LinearWithConstantRoot) {
            final RealParameter tip = ((LinearWithConstantRoot) pop).tipPopSizesInput.get(), top = ((LinearWithConstantRoot) pop).topPopSizesInput.get();
            if (firstPush || tip.somethingIsDirty() || top.somethingIsDirty()) {
                check(Sb2Gpu.setPopModel(engine, Sb2Gpu.POP_LWCR, tip.getDoubleValues(), top.getDoubleValues(), 0, 0)); dirty = true;
            }
        } else if (pop instanceof UniformPopulations) {   // one size for every branch: the constant model with the value broadcast
            final RealParameter u = ((UniformPopulations) pop).universalSizeInput.get();
            if (firstPush || u.somethingIsDirty()) {
                final double[] n = new double[speciesTree.getNodeCount()];
                java.util.Arrays.fill(n, u.getValue());
                check(Sb2Gpu.setPopModel(engine, Sb2Gpu.POP_CONSTANT, n, null, 0, 0)); dirty = true;
            }
        } else if (pop == null || pop instanceof DummyModel) {
            final double a = msc != null && msc.analytic() ? msc.alpha() : 1.0, b = msc != null && msc.analytic() ? msc.beta() : 1.0;
            if (firstPush || (msc != null && msc.hyperparametersChanged())) { check(Sb2Gpu.setPopModel(engine, Sb2Gpu.POP_ANALYTIC, null, null, a, b)); dirty = true; }
        }
        SpeciesTreeRates speciesRates = null;
        for (int l = 0; l < geneTrees.size(); l++) {
            final GpuGeneTree gt = geneTrees.get(l);
            final TreeInterface tree = gt.treeInput.get();
            if (firstPush || gt.treeInput.get().somethingIsDirty()) { pushTree(engine, l, tree); dirty = true; }
            final GpuTreeLikelihood tl = likelihoods.get(l);
            if (tl == null) continue;
            final SiteModel site = (SiteModel) tl.siteModelInput.get();
            if (firstPush || site.isDirtyCalculation()) {
                final SubstitutionModel sm = site.substModelInput.get();
                final double[] pi = sm.getFrequencies();
                if (sm instanceof HKY) {
                    check(Sb2Gpu.setSubstModel(engine, l, Sb2Gpu.SUBST_HKY, new double[]{((HKY) sm).kappaInput.get().getValue(), pi[0], pi[1], pi[2], pi[3]}));
                } else {
                    final EigenDecomposition ed = ((GeneralSubstitutionModel) sm).getEigenDecomposition(null);
                    final double[] p = new double[40];
                    System.arraycopy(ed.getEigenValues(), 0, p, 0, 4);
                    System.arraycopy(ed.getEigenVectors(), 0, p, 4, 16);
                    System.arraycopy(ed.getInverseEigenVectors(), 0, p, 20, 16);
                    System.arraycopy(pi, 0, p, 36, 4);
                    check(Sb2Gpu.setSubstModel(engine, l, Sb2Gpu.SUBST_EIGEN, p));
                }
                final int C = site.getCategoryCount();
                final double[] rates = new double[C];
                for (int c = 0; c < C; c++) rates[c] = site.getRateForCategory(c, null);   // BEAST's own category rates (quirk Q5)
                check(Sb2Gpu.setSiteModel(engine, l, rates, site.getCategoryProportions(null), site.getProportionInvariant()));
                dirty = true;
            }
            final BranchRateModel clk = tl.branchRateModelInput.get();
            if (clk instanceof StarBeastClock) {
                speciesRates = ((StarBeastClock) clk).speciesTreeRatesInput.get();
                if (firstPush || ((StarBeastClock) clk).isDirtyCalculation())
                    { check(Sb2Gpu.setClock(engine, l, Sb2Gpu.CLOCK_SPECIES_RATES, ((RealParameter) clk.meanRateInput.get()).getValue(), null)); dirty = true; }
            } else if (clk == null || clk instanceof StrictClockModel) {
                final double r = clk == null ? 1.0 : ((RealParameter) clk.meanRateInput.get()).getValue();
                if (firstPush || (clk != null && ((StrictClockModel) clk).isDirtyCalculation())) { check(Sb2Gpu.setClock(engine, l, Sb2Gpu.CLOCK_STRICT, r, null)); dirty = true; }
            } else {   // any other BranchRateModel: ship its per-node rates
                final double[] r = new double[tree.getNodeCount()];
                for (Node node : tree.getNodesAsArray()) r[node.getNr()] = clk.getRateForBranch(node);
                check(Sb2Gpu.setClock(engine, l, Sb2Gpu.CLOCK_EXPLICIT, 1.0, r)); dirty = true;
            }
        }
        for (int j = 0; j < morph.size(); j++) {   // likelihoods on the species tree: its arrays are their "gene tree"
            final int l = geneTrees.size() + j;
            final GpuTreeLikelihood tl = morph.get(j);
            if (firstPush || speciesTree.getCurrent().somethingIsDirty()) { pushTree(engine, l, (TreeInterface) speciesTree); dirty = true; }
            final SiteModel site = (SiteModel) tl.siteModelInput.get();
            if (firstPush || site.isDirtyCalculation()) {
                final GeneralSubstitutionModel sm = (GeneralSubstitutionModel) site.substModelInput.get();   // LewisMK extends it
                final EigenDecomposition ed = sm.getEigenDecomposition(null);
                final int K = sm.getStateCount();
                final double[] p = new double[2 * K + 2 * K * K];   // {eval[K], evec[K*K], ievc[K*K], pi[K]}
                System.arraycopy(ed.getEigenValues(), 0, p, 0, K);
                System.arraycopy(ed.getEigenVectors(), 0, p, K, K * K);
                System.arraycopy(ed.getInverseEigenVectors(), 0, p, K + K * K, K * K);
                System.arraycopy(sm.getFrequencies(), 0, p, K + 2 * K * K, K);
                check(Sb2Gpu.setSubstModel(engine, l, Sb2Gpu.SUBST_EIGEN, p));
                final int C = site.getCategoryCount();
                final double[] rates = new double[C];
                for (int c = 0; c < C; c++) rates[c] = site.getRateForCategory(c, null);
                check(Sb2Gpu.setSiteModel(engine, l, rates, site.getCategoryProportions(null), 0.0));
                dirty = true;
            }
            final BranchRateModel clk = tl.branchRateModelInput.get();
            if (clk == null || clk instanceof StrictClockModel) {
                final double r = clk == null ? 1.0 : ((RealParameter) clk.meanRateInput.get()).getValue();
                if (firstPush || (clk != null && ((StrictClockModel) clk).isDirtyCalculation())) { check(Sb2Gpu.setClock(engine, l, Sb2Gpu.CLOCK_STRICT, r, null)); dirty = true; }
            } else {   // e.g. the species tree's relaxed clock applied to morphology: ship its per-node rates
                final TreeInterface tree = (TreeInterface) speciesTree;
                final double[] r = new double[tree.getNodeCount()];
                for (Node node : tree.getNodesAsArray()) r[node.getNr()] = clk.getRateForBranch(node);
                check(Sb2Gpu.setClock(engine, l, Sb2Gpu.CLOCK_EXPLICIT, 1.0, r)); dirty = true;
            }
        }
        if (speciesRates != null) { check(Sb2Gpu.setSpeciesRates(engine, speciesRates.getRatesArray())); dirty = true; }   // SpeciesTreeRates.java:4
        firstPush = false;
        if (dirty) coalValid = likValid = false;
    }

    public double coalescent(int locus) { evalCoalescent(); return perLocusCoal[locus]; }
    public double coalescentTotal() { evalCoalescent(); return coalTotal; }

    // ---- operator-side queries (CoordinatedUniform/Exponential.getConnectingNodes, NodeReheight2.recalculateMaxHeight) ----
    // Operators run after accept()/restore() of the previous step: evalCoalescent() is a cache hit and the device's
    // current state is the chain's current state.
    /** mask[sum of gene node counts] in locus order; freedoms = {tipward, rootward}. */
    public void connectingNodes(int speciesNodeNr, byte[] mask, double[] freedoms) {
        evalCoalescent();
        check(Sb2Gpu.connectingNodes(engine, speciesNodeNr, mask, freedoms, null));
    }
    public double maxHeight(byte[] leafIsRight) {
        evalCoalescent();
        final double[] out = new double[1];
        check(Sb2Gpu.maxHeight(engine, leafIsRight, out));
        return out[0];
    }
    public int nodeOffset(int locus) { int o = 0; for (int i = 0; i < locus; i++) o += geneTrees.get(i).getNodeCount(); return o; }
    public double likelihood(int locus) { evalLikelihood(); return perLocusLik[locus]; }

    private void evalCoalescent() {
        sync();
        if (!coalValid) { check(Sb2Gpu.evalCoalescent(engine, perLocusCoal, perBranch, total)); coalTotal = total[0]; coalValid = true; }
    }
    private void evalLikelihood() {
        sync();
        if (!likValid) { check(Sb2Gpu.evalLikelihood(engine, perLocusLik, total)); likTotal = total[0]; likValid = true; }
    }
    public double[] branchRates(int locus) {
        evalCoalescent();
        final double[] r = new double[geneTrees.get(locus).getNodeCount()];
        check(Sb2Gpu.getBranchRates(engine, locus, r));
        return r;
    }
    public void embedding(int locus, int[] n, int[] k, int[] assignment) { evalCoalescent(); check(Sb2Gpu.getEmbedding(engine, locus, n, k, assignment, null)); }

    public void store() { if (engine != 0) check(Sb2Gpu.store(engine)); }
    public void restore() { if (engine != 0) check(Sb2Gpu.restore(engine)); coalValid = likValid = false; }
    public void accept() { if (engine != 0) check(Sb2Gpu.accept(engine)); }
}
