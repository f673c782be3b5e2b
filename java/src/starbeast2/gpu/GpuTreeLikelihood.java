package starbeast2.gpu;

import java.util.List;
import java.util.Random;

import beast.base.core.Description;
import beast.base.evolution.likelihood.GenericTreeLikelihood;
import beast.base.inference.State;
import starbeast2.StarBeastClock;

/**
 * Drop-in for beast.base.evolution.likelihood.TreeLikelihood as StarBEAST2 wires it
 * (fxtemplates/StarBeast2.xml:158-163: data, tree, siteModel, branchRateModel -- all inherited from
 * GenericTreeLikelihood).  The locus is found through the GpuGeneTree that embeds the same Tree.
 * UNCOMPILED in the build image (no JDK); the tested twin is starbeast2_b200/plugins.py:TreeLikelihood.
 */
@Description("GPU (libsb2gpu) drop-in for TreeLikelihood inside a StarBEAST2 analysis")
public class GpuTreeLikelihood extends GenericTreeLikelihood {
    private GpuSession session;
    private int locus = -1;
    private int morphIndex = -1;   // >= 0: a likelihood on the species tree itself (no GeneTree): K-state, likelihood-only locus

    @Override
    public void initAndValidate() {
        if (treeInput.get() instanceof starbeast2.SpeciesTreeInterface) {
            // tree="@Tree.t:Species": the morphological partitions of the FBD analyses (tutorial/CanisFBD/Canis-FBD.xml:964-1033);
            // tested twin: starbeast2_b200/plugins.py:TreeLikelihood with a SpeciesTree (tests/test_plugins.py)
            if (branchRateModelInput.get() instanceof StarBeastClock) throw new IllegalArgumentException("StarBeastClock needs a gene tree");
            session = GpuSession.of((starbeast2.SpeciesTreeInterface) treeInput.get());
            morphIndex = session.registerMorph(this);
            return;
        }
        if (branchRateModelInput.get() instanceof StarBeastClock) {
            final StarBeastClock clk = (StarBeastClock) branchRateModelInput.get();
            session = GpuSession.of(clk.geneTreeInput.get().speciesTreeInput.get());
        }
        if (session == null) session = GpuSession.ofGeneTree(treeInput.get());   // strict clock: found through the embedding GpuGeneTree
        if (session == null) throw new IllegalArgumentException("GpuTreeLikelihood: declare the starbeast2.GpuGeneTree of this tree first "
                + "(XML order: speciescoalescent before likelihood, as in tutorial/CanisPhylogeny/Canis-Phylogeny.xml:551-684)");
        locus = session.locusOf(treeInput.get());
        if (locus < 0) throw new IllegalArgumentException("GpuTreeLikelihood: tree " + treeInput.get().getID() + " is not embedded by any GpuGeneTree");
        session.register(locus, this);
    }
    @Override public double calculateLogP() { logP = morphIndex >= 0 ? session.morphLikelihood(morphIndex) : session.likelihood(locus); return logP; }
    @Override protected boolean requiresRecalculation() { return true; }   // the engine decides on the device what is dirty
    @Override public List<String> getArguments() { return null; }
    @Override public List<String> getConditions() { return null; }
    @Override public void sample(State state, Random random) {}
}
