package starbeast2;

import java.util.List;
import java.util.Random;

import beast.base.core.Description;
import beast.base.core.Input;
import beast.base.inference.CompoundDistribution;
import beast.base.inference.Distribution;
import beast.base.inference.State;
import beast.base.inference.parameter.RealParameter;
import starbeast2.gpu.GpuSession;

/**
 * Drop-in for starbeast2.MultispeciesCoalescent (MultispeciesCoalescent.java:20-22): inherited `distribution` list
 * (all children must be GeneTrees, :111-118) + populationShape / populationMean for analytic integration.
 * UNCOMPILED in the build image (no JDK); the tested twin is starbeast2_b200/plugins.py:MultispeciesCoalescent.
 */
@Description("GPU (libsb2gpu) drop-in for starbeast2.MultispeciesCoalescent")
public class GpuMultispeciesCoalescent extends CompoundDistribution {
    final public Input<RealParameter> populationShapeInput = new Input<>("populationShape", "Shape of the inverse gamma prior distribution on population sizes.");
    final public Input<RealParameter> populationMeanInput = new Input<>("populationMean", "Mean of the inverse gamma prior distribution on population sizes.");

    private GpuSession session;
    private boolean dontCalculate, changed;
    private double alpha, beta, storedAlpha, storedBeta;

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        if (populationShapeInput.get() == null ^ populationMeanInput.get() == null)
            throw new IllegalArgumentException("Either specify both population size prior parameters for analytical integration,"
                    + "or neither for MCMC integration of population sizes.");                     // :93-96
        dontCalculate = populationShapeInput.get() == null;
        final List<Distribution> geneTrees = pDistributions.get();
        for (Distribution d : geneTrees)
            if (!(d instanceof GpuGeneTree)) throw new IllegalArgumentException("Input distributions must all be of class GeneTree.");   // :116-118
        if (!geneTrees.isEmpty()) {
            session = GpuSession.of(((GpuGeneTree) geneTrees.get(0)).speciesTreeInput.get());
            session.setMsc(this);
        }
        if (!dontCalculate) checkHyperparameters(true);
    }

    /** MultispeciesCoalescent.java:134-147, stale-alpha quirk included: beta uses the PREVIOUS alpha. */
    private boolean checkHyperparameters(final boolean force) {
        final double currentAlpha = populationShapeInput.get().getValue();
        final double currentBeta = populationMeanInput.get().getValue() * (alpha - 1.0);
        if (force || currentAlpha != alpha || currentBeta != beta) { alpha = currentAlpha; beta = currentBeta; changed = true; return true; }
        return false;
    }
    public boolean analytic() { return !dontCalculate; }
    public double alpha() { return alpha; }
    public double beta() { return beta; }
    public boolean hyperparametersChanged() { final boolean c = changed; changed = false; return c; }

    @Override
    public double calculateLogP() {                                                              // :150-198
        if (!dontCalculate) checkHyperparameters(false);
        logP = session.coalescentTotal();                                                        // -infinity short-circuits the posterior
        return logP;
    }
    @Override public void store() { super.store(); storedAlpha = alpha; storedBeta = beta; }
    @Override public void restore() { super.restore(); alpha = storedAlpha; beta = storedBeta; }
    @Override public List<String> getArguments() { return null; }
    @Override public List<String> getConditions() { return null; }
    @Override public void sample(State state, Random random) {}
}
