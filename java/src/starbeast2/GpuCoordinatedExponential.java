package starbeast2;

import beast.base.core.Description;
import beast.base.core.Input;
import beast.base.evolution.tree.Node;
import beast.base.evolution.tree.Tree;
import beast.base.util.Randomizer;
import starbeast2.gpu.GpuSession;

import java.util.List;

/**
 * UNCOMPILED in the build image (no JDK); mirrors starbeast2_b200/operators.py:CoordinatedExponential, which is tested.
 *
 * CoordinatedExponential with getConnectingNodes() (CoordinatedExponential.java:92-160) replaced by one device query.
 * Same inputs (speciesTree, geneTree, beta, optimise), same exponential proposal, same Hastings ratio and the same
 * beta auto-tuning rule (CoordinatedExponential.java:184-191).
 */
@Description("Co-ordinated exponential move of the species root and its connected gene nodes; connected components found on the GPU.")
public class GpuCoordinatedExponential extends CoordinatedOperator {
    public final Input<Double> betaInput = new Input<>("beta", "Beta parameter of the exponential proposal distribution", 1.0);
    public final Input<Boolean> optimiseInput = new Input<>("optimise", "Adjust beta parameter during the MCMC run to improve mixing.", true);

    private static final double WAITING_TIME_SCALE = 1.4426950408889634;
    private SpeciesTreeInterface speciesTree;
    private GpuSession session;
    private byte[] mask;
    private final double[] freedoms = new double[2];
    private boolean optimise;
    private double beta, lambda, waitingTime;

    @Override
    public void initAndValidate() {
        beta = betaInput.get();
        lambda = 1.0 / beta;
        optimise = optimiseInput.get();
        speciesTree = speciesTreeInput.get();
        super.initAndValidate();
        session = GpuSession.of(speciesTree);
        int total = 0;
        for (Tree t : geneTreeInput.get()) total += t.getNodeCount();
        mask = new byte[total];
    }

    @Override
    public double proposal() {
        final Node root = speciesTree.getRoot();
        if (root.isFake()) return Double.NEGATIVE_INFINITY;
        final double h = root.getHeight();
        session.connectingNodes(root.getNr(), mask, freedoms);                   // freedoms[1] (rootward) is not used here
        final double twf = Math.min(freedoms[0], Math.min(h - root.getLeft().getHeight(), h - root.getRight().getHeight()));
        waitingTime = twf * WAITING_TIME_SCALE;
        final double shift = Randomizer.nextExponential(lambda) - twf;
        root.setHeight(h + shift);
        final List<Tree> geneTrees = geneTreeInput.get();
        int o = 0;
        for (int j = 0; j < nGeneTrees; j++) {
            final Tree g = geneTrees.get(j);
            final int G = g.getNodeCount();
            for (int i = 0; i < G; i++)
                if (mask[o + i] != 0) { final Node x = g.getNode(i); x.setHeight(x.getHeight() + shift); }
            o += G;
        }
        return lambda * shift;
    }

    @Override
    public double getCoercableParameterValue() { return beta; }

    @Override
    public void setCoercableParameterValue(final double value) { beta = value; lambda = 1.0 / beta; }

    @Override
    public void optimize(final double logAlpha) {
        if (optimise) {
            final double count = (m_nNrRejectedForCorrection + m_nNrAcceptedForCorrection + 1.0);
            setCoercableParameterValue(beta + (waitingTime - beta) / count);
        }
    }
}
