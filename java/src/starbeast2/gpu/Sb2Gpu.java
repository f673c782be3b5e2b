package starbeast2.gpu;

/**
 * JNI view of libsb2gpu's C ABI (include/sb2gpu.h), one native method per entry point.
 * UNCOMPILED in the build image (no JDK); see java/README.md.
 * All methods return the C return code (0 = OK, &lt;0 = error; message via lastError).
 */
public final class Sb2Gpu {
    static { System.loadLibrary("sb2gpu_jni"); }
    private Sb2Gpu() {}

    public static final int SUBST_HKY = 0, SUBST_EIGEN = 1;
    public static final int CLOCK_STRICT = 0, CLOCK_SPECIES_RATES = 1, CLOCK_EXPLICIT = 2;
    public static final int POP_CONSTANT = 0, POP_LWCR = 1, POP_ANALYTIC = 2;

    public static native long create(int device, boolean exactOrder);              // sb2_create; 0 on failure
    public static native void destroy(long engine);                                // sb2_destroy
    public static native String lastError(long engine);                            // sb2_last_error
    public static native int addLocus(long engine, int nTips, int nPatterns, byte[] tipStates, int[] patternWeights,
                                      int nCategories, int[] tipToSpecies, double ploidy);   // sb2_add_locus
    // likelihood-only K-state locus (tipToSpecies == null): the morphological TreeLikelihoods on the species tree; sb2_add_locus_k
    public static native int addLocusK(long engine, int nTips, int nPatterns, int nStates, byte[] tipStates, int[] patternWeights,
                                       int nCategories, int[] tipToSpeciesOrNull, double ploidy, int[] excludedPatternsOrNull);
    public static native int finalizeEngine(long engine, int nSpeciesNodes, int nSpeciesLeaves);   // sb2_finalize
    public static native int setSpeciesTree(long engine, int[] parent, int[] left, int[] right, double[] height);
    public static native int setGeneTree(long engine, int locus, int[] parent, int[] left, int[] right, double[] height);
    public static native int setSubstModel(long engine, int locus, int kind, double[] params);
    public static native int setSiteModel(long engine, int locus, double[] catRates, double[] catProps, double pInv);
    public static native int setClock(long engine, int locus, int kind, double meanRate, double[] ratesOrNull);
    public static native int setSpeciesRates(long engine, double[] rates);
    public static native int setPopModel(long engine, int kind, double[] a, double[] b, double alpha, double beta);
    public static native int evalCoalescent(long engine, double[] perLocus, double[] perBranch, double[] total1);
    public static native int evalLikelihood(long engine, double[] perLocus, double[] total1);
    public static native int evalPosterior(long engine, double[] perLocusCoal, double[] perLocusLik, double[] perBranch, double[] total3);
    public static native int store(long engine);
    public static native int restore(long engine);
    public static native int accept(long engine);
    public static native int markAllDirty(long engine);
    public static native int getBranchRates(long engine, int locus, double[] rates);
    public static native int getEmbedding(long engine, int locus, int[] lineageCounts, int[] coalCounts, int[] assignment, double[] perBranchLogP);
    // operator-side walks over the resident gene trees (sb2gpu.h: sb2_connecting_nodes, sb2_max_height)
    public static native int connectingNodes(long engine, int speciesNode, byte[] maskOrNull, double[] freedoms2, long[] count1OrNull);
    public static native int maxHeight(long engine, byte[] leafIsRight, double[] maxHeight1);
    // several engines (one per GPU) driven by this JVM thread: begin on all, sum the reduce vectors on the host, end on all
    public static native int evalBegin(long engine);
    public static native int evalEnd(long engine, double[] perLocusCoal, double[] perLocusLik, double[] perBranch, double[] total3);
    public static native int reduceVectorLength(long engine);
    public static native int readReduceVector(long engine, double[] host);
    public static native int writeReduceVector(long engine, double[] host);
    public static native int setGeneTrees(long engine, int firstLocus, int nLoci, int[] parent, int[] left, int[] right, double[] height);
    public static native int numLoci(long engine);
    public static native long launchCount(long engine);
}
