package starbeast2.gpu;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.foreign.ValueLayout;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM (java.lang.foreign, final in JDK 22) twin of {@link Sb2Gpu}: the same static methods with the same
 * signatures, bound straight to libsb2gpu.so's C ABI (include/sb2gpu.h) -- no JNI shim, no C compiler on the user's side.
 * UNCOMPILED in the build image (no JDK); see java/README.md.
 *
 * Java arrays are passed as heap segments through downcall handles linked with {@code Linker.Option.critical(true)}
 * (the FFM counterpart of GetPrimitiveArrayCritical): the engine copies inputs before it returns and writes outputs into
 * the caller's arrays, and no call blocks on anything but the GPU, so the critical window is one evaluation at most.
 * GpuSession selects this class instead of Sb2Gpu when -Dsb2gpu.ffm=true and the runtime is JDK 22 or newer.
 */
public final class Sb2GpuFfm {
    private Sb2GpuFfm() {}

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("sb2gpu.lib", System.mapLibraryName("sb2gpu")), Arena.global());
    private static final Linker.Option HEAP = Linker.Option.critical(true);

    private static MethodHandle bind(String name, MemoryLayout ret, MemoryLayout... args)
    {
        final MemorySegment sym = LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name + " not found in libsb2gpu"));
        final FunctionDescriptor fd = ret == null ? FunctionDescriptor.ofVoid(args) : FunctionDescriptor.of(ret, args);
        return LINKER.downcallHandle(sym, fd, HEAP);
    }

    private static MemorySegment seg(double[] a) { return a == null ? MemorySegment.NULL : MemorySegment.ofArray(a); }
    private static MemorySegment seg(int[] a) { return a == null ? MemorySegment.NULL : MemorySegment.ofArray(a); }
    private static MemorySegment seg(byte[] a) { return a == null ? MemorySegment.NULL : MemorySegment.ofArray(a); }
    private static MemorySegment seg(long[] a) { return a == null ? MemorySegment.NULL : MemorySegment.ofArray(a); }
    private static MemorySegment eng(long h) { return MemorySegment.ofAddress(h); }

    private static int rc(MethodHandle mh, Object... args)
    {
        try {
            return (int) mh.invokeWithArguments(args);
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    // sb2_config {int32 device, exact_order, profile, reserved[5]} = 32 bytes
    private static final MethodHandle CREATE = bind("sb2_create", JAVA_INT, ADDRESS, ADDRESS);
    private static final MethodHandle DESTROY = bind("sb2_destroy", null, ADDRESS);
    private static final MethodHandle LAST_ERROR = bind("sb2_last_error", ADDRESS, ADDRESS);
    private static final MethodHandle ADD_LOCUS = bind("sb2_add_locus", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_DOUBLE);
    private static final MethodHandle ADD_LOCUS_K = bind("sb2_add_locus_k", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_DOUBLE, JAVA_INT, ADDRESS);
    private static final MethodHandle FINALIZE = bind("sb2_finalize", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT);
    private static final MethodHandle SET_SPECIES_TREE = bind("sb2_set_species_tree", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle SET_GENE_TREE = bind("sb2_set_gene_tree", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle SET_GENE_TREES = bind("sb2_set_gene_trees", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle SET_SUBST = bind("sb2_set_subst_model", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS);
    private static final MethodHandle SET_SITE = bind("sb2_set_site_model", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_DOUBLE);
    private static final MethodHandle SET_CLOCK = bind("sb2_set_clock", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_DOUBLE, ADDRESS);
    private static final MethodHandle SET_SPECIES_RATES = bind("sb2_set_species_rates", JAVA_INT, ADDRESS, ADDRESS);
    private static final MethodHandle SET_POP = bind("sb2_set_pop_model", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE);
    private static final MethodHandle EVAL_COAL = bind("sb2_eval_coalescent", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle EVAL_LIK = bind("sb2_eval_likelihood", JAVA_INT, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle EVAL_POST = bind("sb2_eval_posterior", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle EVAL_BEGIN = bind("sb2_eval_begin", JAVA_INT, ADDRESS);
    private static final MethodHandle EVAL_END = bind("sb2_eval_end", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle STORE = bind("sb2_store", JAVA_INT, ADDRESS);
    private static final MethodHandle RESTORE = bind("sb2_restore", JAVA_INT, ADDRESS);
    private static final MethodHandle ACCEPT = bind("sb2_accept", JAVA_INT, ADDRESS);
    private static final MethodHandle MARK_ALL = bind("sb2_mark_all_dirty", JAVA_INT, ADDRESS);
    private static final MethodHandle GET_RATES = bind("sb2_get_branch_rates", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS);
    private static final MethodHandle GET_EMBED = bind("sb2_get_embedding", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle CONNECTING = bind("sb2_connecting_nodes", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle MAX_HEIGHT = bind("sb2_max_height", JAVA_INT, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle REDUCE_VECTOR = bind("sb2_reduce_vector", JAVA_INT, ADDRESS, ADDRESS, ADDRESS);
    private static final MethodHandle READ_REDUCE = bind("sb2_read_reduce_vector", JAVA_INT, ADDRESS, ADDRESS);
    private static final MethodHandle WRITE_REDUCE = bind("sb2_write_reduce_vector", JAVA_INT, ADDRESS, ADDRESS);
    private static final MethodHandle NUM_LOCI = bind("sb2_num_loci", JAVA_INT, ADDRESS);
    private static final MethodHandle LAUNCH_COUNT = bind("sb2_launch_count", JAVA_LONG, ADDRESS);

    /** sb2_create; throws (no CPU fallback) when no sm_100 device is usable, like the JNI twin. */
    public static long create(int device, boolean exactOrder)
    {
        try (Arena a = Arena.ofConfined()) {
            final MemorySegment cfg = a.allocate(32, 4);
            cfg.fill((byte) 0);
            cfg.set(JAVA_INT, 0, device);
            cfg.set(JAVA_INT, 4, exactOrder ? 1 : 0);
            final MemorySegment out = a.allocate(ADDRESS);
            final int rc = (int) CREATE.invokeExact(cfg, out);
            final MemorySegment h = out.get(ADDRESS, 0);
            if (rc != 0) {
                final String msg = h.equals(MemorySegment.NULL) ? "sb2_create failed" : lastError(h.address());
                if (!h.equals(MemorySegment.NULL)) DESTROY.invokeExact(h);
                throw new IllegalStateException(msg);
            }
            return h.address();
        } catch (IllegalStateException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    public static void destroy(long engine)
    {
        try { DESTROY.invokeExact(eng(engine)); } catch (Throwable t) { throw new IllegalStateException(t); }
    }

    public static String lastError(long engine)
    {
        try {
            final MemorySegment p = (MemorySegment) LAST_ERROR.invokeExact(eng(engine));
            return p.reinterpret(512).getString(0);
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    public static int addLocus(long engine, int nTips, int nPatterns, byte[] tipStates, int[] patternWeights, int nCategories, int[] tipToSpecies, double ploidy)
    { return rc(ADD_LOCUS, eng(engine), nTips, nPatterns, seg(tipStates), seg(patternWeights), nCategories, seg(tipToSpecies), ploidy); }
    public static int addLocusK(long engine, int nTips, int nPatterns, int nStates, byte[] tipStates, int[] patternWeights, int nCategories,
                                int[] tipToSpeciesOrNull, double ploidy, int[] excludedPatternsOrNull)
    {
        return rc(ADD_LOCUS_K, eng(engine), nTips, nPatterns, nStates, seg(tipStates), seg(patternWeights), nCategories, seg(tipToSpeciesOrNull), ploidy,
                  excludedPatternsOrNull == null ? 0 : excludedPatternsOrNull.length, seg(excludedPatternsOrNull));
    }
    public static int finalizeEngine(long engine, int nSpeciesNodes, int nSpeciesLeaves) { return rc(FINALIZE, eng(engine), nSpeciesNodes, nSpeciesLeaves); }
    public static int setSpeciesTree(long engine, int[] parent, int[] left, int[] right, double[] height)
    { return rc(SET_SPECIES_TREE, eng(engine), seg(parent), seg(left), seg(right), seg(height)); }
    public static int setGeneTree(long engine, int locus, int[] parent, int[] left, int[] right, double[] height)
    { return rc(SET_GENE_TREE, eng(engine), locus, seg(parent), seg(left), seg(right), seg(height)); }
    public static int setGeneTrees(long engine, int firstLocus, int nLoci, int[] parent, int[] left, int[] right, double[] height)
    { return rc(SET_GENE_TREES, eng(engine), firstLocus, nLoci, seg(parent), seg(left), seg(right), seg(height)); }
    public static int setSubstModel(long engine, int locus, int kind, double[] params) { return rc(SET_SUBST, eng(engine), locus, kind, seg(params)); }
    public static int setSiteModel(long engine, int locus, double[] catRates, double[] catProps, double pInv)
    { return rc(SET_SITE, eng(engine), locus, seg(catRates), seg(catProps), pInv); }
    public static int setClock(long engine, int locus, int kind, double meanRate, double[] ratesOrNull)
    { return rc(SET_CLOCK, eng(engine), locus, kind, meanRate, seg(ratesOrNull)); }
    public static int setSpeciesRates(long engine, double[] rates) { return rc(SET_SPECIES_RATES, eng(engine), seg(rates)); }
    public static int setPopModel(long engine, int kind, double[] a, double[] b, double alpha, double beta)
    { return rc(SET_POP, eng(engine), kind, seg(a), seg(b), alpha, beta); }
    public static int evalCoalescent(long engine, double[] perLocus, double[] perBranch, double[] total1)
    { return rc(EVAL_COAL, eng(engine), seg(perLocus), seg(perBranch), seg(total1)); }
    public static int evalLikelihood(long engine, double[] perLocus, double[] total1) { return rc(EVAL_LIK, eng(engine), seg(perLocus), seg(total1)); }
    public static int evalPosterior(long engine, double[] perLocusCoal, double[] perLocusLik, double[] perBranch, double[] total3)
    { return rc(EVAL_POST, eng(engine), seg(perLocusCoal), seg(perLocusLik), seg(perBranch), seg(total3)); }
    public static int evalBegin(long engine) { return rc(EVAL_BEGIN, eng(engine)); }
    public static int evalEnd(long engine, double[] perLocusCoal, double[] perLocusLik, double[] perBranch, double[] total3)
    { return rc(EVAL_END, eng(engine), seg(perLocusCoal), seg(perLocusLik), seg(perBranch), seg(total3)); }
    public static int store(long engine) { return rc(STORE, eng(engine)); }
    public static int restore(long engine) { return rc(RESTORE, eng(engine)); }
    public static int accept(long engine) { return rc(ACCEPT, eng(engine)); }
    public static int markAllDirty(long engine) { return rc(MARK_ALL, eng(engine)); }
    public static int getBranchRates(long engine, int locus, double[] rates) { return rc(GET_RATES, eng(engine), locus, seg(rates)); }
    public static int getEmbedding(long engine, int locus, int[] lineageCounts, int[] coalCounts, int[] assignment, double[] perBranchLogP)
    { return rc(GET_EMBED, eng(engine), locus, seg(lineageCounts), seg(coalCounts), seg(assignment), seg(perBranchLogP)); }
    public static int connectingNodes(long engine, int speciesNode, byte[] maskOrNull, double[] freedoms2, long[] count1OrNull)
    { return rc(CONNECTING, eng(engine), speciesNode, seg(maskOrNull), seg(freedoms2), seg(count1OrNull)); }
    public static int maxHeight(long engine, byte[] leafIsRight, double[] maxHeight1) { return rc(MAX_HEIGHT, eng(engine), seg(leafIsRight), seg(maxHeight1)); }
    public static int reduceVectorLength(long engine)
    {
        try (Arena a = Arena.ofConfined()) {
            final MemorySegment dev = a.allocate(ADDRESS), n = a.allocate(ValueLayout.JAVA_INT);
            // off-heap outputs: this handle is not called with heap segments
            return rc(REDUCE_VECTOR, eng(engine), dev, n) == 0 ? n.get(JAVA_INT, 0) : -1;
        }
    }
    public static int readReduceVector(long engine, double[] host) { return rc(READ_REDUCE, eng(engine), seg(host)); }
    public static int writeReduceVector(long engine, double[] host) { return rc(WRITE_REDUCE, eng(engine), seg(host)); }
    public static int numLoci(long engine) { return rc(NUM_LOCI, eng(engine)); }
    public static long launchCount(long engine)
    {
        try { return (long) LAUNCH_COUNT.invokeExact(eng(engine)); } catch (Throwable t) { throw new IllegalStateException(t); }
    }
}
