package starbeast2;

import java.util.Map;

import beast.base.core.Description;
import starbeast2.gpu.GpuSession;

/**
 * Drop-in for starbeast2.GeneTree (GeneTree.java): same inputs (tree, speciesTree, ploidy, populationModel), same
 * protocol, but the embedding and the per-branch coalescent terms come from libsb2gpu.  It SUBCLASSES GeneTree (and
 * lives in its package) because other reference classes hold concrete GeneTree references and read its
 * package-private members: StarBeastClock.geneTreeInput (StarBeastClock.java:10), NodeReheight2.geneTreesInput
 * (NodeReheight2.java:25,68,171), SpeciesTreeLogger (SpeciesTreeLogger.java:19,178-180),
 * MultispeciesCoalescent (MultispeciesCoalescent.java:111-115,172-177).
 * UNCOMPILED in the build image (no JDK); the tested twin is starbeast2_b200/plugins.py:GeneTree.
 */
@Description("GPU (libsb2gpu) drop-in for starbeast2.GeneTree")
public class GpuGeneTree extends GeneTree {
    private GpuSession session;
    private int locus;
    private PopulationModel base;

    @Override
    public void initAndValidate() {
        super.initAndValidate();                         // allocates the host arrays other classes may still look at
        base = popModelInput.get() == null ? null : popModelInput.get().getBaseModel();   // GeneTree.java:126
        final Map<String, Integer> tipNumberMap = speciesTreeInput.get().getTipNumberMap();
        for (int i = 0; i < treeInput.get().getLeafNodeCount(); i++)                       // GeneTree.java:216-223
            leafGeneNodeSpeciesAssignment[i] = tipNumberMap.get(treeInput.get().getNode(i).getID());
        session = GpuSession.of(speciesTreeInput.get());
        locus = session.register(this);
    }

    public PopulationModel baseModel() { return base; }
    public int locus() { return locus; }

    @Override public boolean requiresRecalculation() { return true; }                    // GeneTree.java:73-76
    @Override public double calculateLogP() { logP = session.coalescent(locus); return logP; }   // -infinity when incompatible (:175-178)
    @Override public void store() { storedLogP = logP; if (locus == 0) session.store(); }
    @Override public void restore() { logP = storedLogP; if (locus == 0) session.restore(); }
    @Override protected void accept() { if (locus == 0) session.accept(); super.accept(); }

    /** StarBeastClock on the host (only if a stock StarBeastClock is still wired): rebuild occupancy lazily from the device embedding. */
    @Override public double[] getSpeciesOccupancy() { update(); return super.getSpeciesOccupancy(); }
    /** NodeReheight2 / loggers. */
    @Override int[] getTipNumberMap() { return leafGeneNodeSpeciesAssignment; }
    /** device-resident counts, pulled on demand */
    public void fetchCounts() { session.embedding(locus, coalescentLineageCounts, coalescentCounts, geneNodeSpeciesAssignment); }
}
