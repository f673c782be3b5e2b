package starbeast2;

import beast.base.core.Description;
import beast.base.evolution.tree.Node;
import beast.base.evolution.tree.Tree;
import beast.base.util.Randomizer;
import starbeast2.gpu.GpuSession;

import java.util.List;

/**
 * UNCOMPILED in the build image (no JDK); mirrors starbeast2_b200/operators.py:CoordinatedUniform, which is tested.
 *
 * CoordinatedUniform with getConnectingNodes() (CoordinatedUniform.java:94-177: a recursive walk of every gene tree with
 * HashSet<String> lookups at the leaves) replaced by one device query over the gene trees the engine already holds.
 * Same inputs (speciesTree, geneTree), same proposal distribution, same Hastings ratio.
 */
@Description("Co-ordinated uniform move of a species node and its connected gene nodes; connected components found on the GPU.")
public class GpuCoordinatedUniform extends CoordinatedOperator {
    private SpeciesTreeInterface speciesTree;
    private GpuSession session;
    private byte[] mask;
    private final double[] freedoms = new double[2];

    @Override
    public void initAndValidate() {
        speciesTree = speciesTreeInput.get();
        super.initAndValidate();
        session = GpuSession.of(speciesTree);
        int total = 0;
        for (Tree t : geneTreeInput.get()) total += t.getNodeCount();
        mask = new byte[total];
    }

    @Override
    public double proposal() {
        final int nLeaves = speciesTree.getLeafNodeCount();
        final int nNodes = 2 * nLeaves - 1;
        final Node[] nodes = speciesTree.getNodesAsArray();
        final Node[] candidates = new Node[nNodes];
        int n = 0;
        for (int i = nLeaves; i < nNodes; i++)
            if (!(nodes[i].isFake() || nodes[i].isRoot())) candidates[n++] = nodes[i];
        if (n == 0) return Double.NEGATIVE_INFINITY;
        final Node s = candidates[Randomizer.nextInt(n)];
        final double h = s.getHeight();

        session.connectingNodes(s.getNr(), mask, freedoms);                      // all gene trees, one kernel
        final double twf = Math.min(freedoms[0], Math.min(h - s.getLeft().getHeight(), h - s.getRight().getHeight()));
        final double rwf = Math.min(freedoms[1], s.getParent().getHeight() - h);
        final double shift = (Randomizer.nextDouble() * (twf + rwf)) - twf;

        s.setHeight(h + shift);
        final List<Tree> geneTrees = geneTreeInput.get();
        int o = 0;
        for (int j = 0; j < nGeneTrees; j++) {
            final Tree g = geneTrees.get(j);
            final int G = g.getNodeCount();
            for (int i = 0; i < G; i++)
                if (mask[o + i] != 0) { final Node x = g.getNode(i); x.setHeight(x.getHeight() + shift); }
            o += G;
        }
        return 0.0;
    }
}
