/*
 * sb2gpu.h -- C ABI of libsb2gpu.so, the B200 (sm_100a) multi-locus likelihood engine that sits
 * behind StarBEAST2's own plugin classes.
 *
 * The entry points are what a JNI (or Panama FFM) shim for the reference's hot path binds; each one
 * cites the reference interface it replaces (paths relative to /root/reference).  Conventions:
 *   - plain pointers and sizes only; the caller owns every pointer; the engine copies inputs before
 *     returning and writes outputs into caller buffers; no callbacks; no global state besides the handle;
 *   - return 0 = OK, <0 = error code, message via sb2_last_error();
 *   - NOT thread-safe: one caller (the reference runs its MCMC single-threaded,
 *     tutorial/StarBEAST2-tutorial.tex:419 "threads: 1");
 *   - every array is indexed by the *current* BEAST node number: leaves 0..tips-1, root = last node
 *     (StarBeastClock.java:60,72; LinearWithConstantRoot.java:42-44);
 *   - numerical failure is reported the way the reference does it: -INFINITY in the output
 *     (GeneTree.java:175-178), never an error code.
 *   - there is no CPU fallback: every call fails with SB2_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SB2GPU_H
#define SB2GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB2_API __attribute__((visibility("default")))

typedef struct sb2_engine sb2_engine;

enum { SB2_OK = 0, SB2_ERR_ARG = -1, SB2_ERR_STATE = -2, SB2_ERR_CUDA = -3, SB2_ERR_NOMEM = -4 };

/* substitution model kinds: HKY.getTransitionProbabilities / GeneralSubstitutionModel (BEAST.base; named at
 * tutorial/CanisPhylogeny/Canis-Phylogeny.xml:688, fxtemplates/sbSubstModels.xml:75-215) */
enum { SB2_SUBST_HKY = 0, SB2_SUBST_EIGEN = 1 };
/* branch-rate model kinds: StrictClockModel (Canis-Phylogeny.xml:692), StarBeastClock (StarBeastClock.java:9),
 * or explicit per-node rates from any other BranchRateModel.getRateForBranch */
enum { SB2_CLOCK_STRICT = 0, SB2_CLOCK_SPECIES_RATES = 1, SB2_CLOCK_EXPLICIT = 2 };
/* population models: ConstantPopulations.java, LinearWithConstantRoot.java, and analytic integration
 * (MultispeciesCoalescent.java:200-235 with DummyModel / no populationModel, GeneTree.java:182) */
enum { SB2_POP_CONSTANT = 0, SB2_POP_LWCR = 1, SB2_POP_ANALYTIC = 2 };

typedef struct {
    int32_t device;       /* CUDA device ordinal */
    int32_t exact_order;  /* 1: no FMA in the pruning recursion and strictly sequential pattern / locus sums,
                                i.e. the reference's (Java) evaluation order; 0: fused multiply-add + fixed-shape
                                tree reductions (deterministic, bit-stable run to run) */
    int32_t profile;      /* 1: bracket the pruning kernel with CUDA events (sb2_kernel_time) */
    int32_t reserved[5];
} sb2_config;

/* ---- life cycle ------------------------------------------------------------------------------------ */

SB2_API int sb2_create(const sb2_config *cfg, sb2_engine **out);
SB2_API void sb2_destroy(sb2_engine *e);
SB2_API const char *sb2_last_error(const sb2_engine *e);

/* Registers one locus = one (GeneTree, TreeLikelihood) pair.  Replaces TreeLikelihood.initAndValidate's
 * tip-state upload (setStates per leaf) and GeneTree.update's leaf->species map (GeneTree.java:211-226).
 * tipStates[tip*nPatterns + pattern] in 0..3 (A,C,G,T), >=4 = gap/ambiguous/missing (useAmbiguities=false);
 * patternWeights = Alignment.getPatternWeight; tipToSpecies[tip] = species-tree LEAF NODE NUMBER
 * (SpeciesTreeInterface.getTipNumberMap).  Returns the locus index (>=0) or an error (<0). */
SB2_API int sb2_add_locus(sb2_engine *e, int32_t nTips, int32_t nPatterns, const uint8_t *tipStates,
                          const int32_t *patternWeights, int32_t nCategories, const int32_t *tipToSpecies,
                          double ploidy);

/* General form.  tipToSpecies != NULL: as sb2_add_locus (nStates must be 4, nExcluded 0).
 * tipToSpecies == NULL: a LIKELIHOOD-ONLY locus -- a TreeLikelihood whose tree is not a GeneTree, i.e. the morphological
 * partitions that the FBD configurations score on the species tree (tutorial/CanisFBD/Canis-FBD.xml:964-1033:
 * spec="TreeLikelihood" tree="@Tree.t:Species", StandardData nrOfStates 2..5, LewisMK, ascertained FilteredAlignment).
 * It has no embedding and no coalescent term (its perLocus coalescent entry is 0); nStates in 2..5, states >= nStates are
 * gaps / '?'; its tree is staged with sb2_set_gene_tree like any other (for these partitions: the species tree's arrays, tips =
 * the species tree's leaves), its model with sb2_set_subst_model(SB2_SUBST_EIGEN, {eval[K], evec[K*K], ievc[K*K], pi[K]}),
 * its clock is STRICT or EXPLICIT, proportionInvariant must be 0.  excludedPatterns[nExcluded] = Alignment.excludedPatterns of
 * an ascertained alignment (their weight is 0): logP = sum_p (patternLogL[p] - log(1 - sum_excluded exp(patternLogL))) * weight[p]
 * (Alignment.getAscertainmentCorrection, TreeLikelihood.calcLogP; BEAST.base).  Returns the locus index or an error. */
SB2_API int sb2_add_locus_k(sb2_engine *e, int32_t nTips, int32_t nPatterns, int32_t nStates, const uint8_t *tipStates,
                            const int32_t *patternWeights, int32_t nCategories, const int32_t *tipToSpecies, double ploidy,
                            int32_t nExcluded, const int32_t *excludedPatterns);

/* Shards nothing (one engine = one GPU; the host shards loci over engines), allocates device state. */
SB2_API int sb2_finalize(sb2_engine *e, int32_t nSpeciesNodes, int32_t nSpeciesLeaves);

/* ---- state input (all staged on the host, uploaded by the next sb2_eval_*) --------------------------- */

/* SpeciesTree state node (SpeciesTree.java; read at GeneTree.java:257,337-343).  left/right = -1 for leaves. */
SB2_API int sb2_set_species_tree(sb2_engine *e, const int32_t *parent, const int32_t *left, const int32_t *right,
                                 const double *height);
/* Gene Tree state node of one locus (GeneTree.java:22 `tree`, TreeLikelihood `tree`). */
SB2_API int sb2_set_gene_tree(sb2_engine *e, int32_t locus, const int32_t *parent, const int32_t *left,
                              const int32_t *right, const double *height);
/* Bulk form: nLoci consecutive loci, arrays concatenated in locus order. */
SB2_API int sb2_set_gene_trees(sb2_engine *e, int32_t firstLocus, int32_t nLoci, const int32_t *parent,
                               const int32_t *left, const int32_t *right, const double *height);
/* HKY: params = {kappa, piA, piC, piG, piT}; EIGEN: {eval[4], evec[16], ievc[16], pi[4]} (row-major, the arrays
 * EigenDecomposition.getEigenValues/getEigenVectors/getInverseEigenVectors return). */
SB2_API int sb2_set_subst_model(sb2_engine *e, int32_t locus, int32_t kind, const double *params);
/* SiteModel: catRates[c] = getRateForCategory(c) (already x mutationRate), catProps = getCategoryProportions,
 * pInv = getProportionInvariant (added on constant patterns at the root, as TreeLikelihood does). */
SB2_API int sb2_set_site_model(sb2_engine *e, int32_t locus, const double *catRates, const double *catProps,
                               double pInv);
/* branchRateModel of the locus.  STRICT: rate = meanRate; SPECIES_RATES: StarBeastClock.update
 * (StarBeastClock.java:54-75) with meanRate = clock.rate; EXPLICIT: rates[node] given (2*nTips-1 values). */
SB2_API int sb2_set_clock(sb2_engine *e, int32_t locus, int32_t kind, double meanRate, const double *ratesOrNull);
/* SpeciesTreeRates.getRatesArray() (SpeciesTreeRates.java:4; UncorrelatedRates.java:173-182), length nSpeciesNodes. */
SB2_API int sb2_set_species_rates(sb2_engine *e, const double *rates);
/* CONSTANT: a = populationSizes[nSpeciesNodes] (ConstantPopulations.java:19);
 * LWCR: a = tipPopulationSizes[nSpeciesLeaves], b = topPopulationSizes[nSpeciesNodes-1] (LinearWithConstantRoot.java:21-22);
 * ANALYTIC: alpha = populationShape, beta = as MultispeciesCoalescent.checkHyperparameters computes it
 * (MultispeciesCoalescent.java:134-147, stale-alpha quirk included by the caller). */
SB2_API int sb2_set_pop_model(sb2_engine *e, int32_t kind, const double *a, const double *b, double alpha, double beta);

/* ---- evaluation ---------------------------------------------------------------------------------------- */

/* MultispeciesCoalescent.calculateLogP (MultispeciesCoalescent.java:150-198): embeds every dirty gene tree
 * (GeneTree.update, GeneTree.java:202-318), evaluates the per-branch terms.  perLocus[nLoci] = GeneTree.calculateLogP
 * (-INFINITY when incompatible); perBranch[nSpeciesNodes] = analytic per-branch terms (zeros otherwise);
 * *total = the distribution's logP.  Output pointers may be NULL. */
SB2_API int sb2_eval_coalescent(sb2_engine *e, double *perLocus, double *perBranch, double *total);
/* Sum of TreeLikelihood.calculateLogP over loci (the `likelihood` CompoundDistribution, Canis-Phylogeny.xml:683),
 * recomputing only dirty branches / nodes.  perLocus[nLoci], *total. */
SB2_API int sb2_eval_likelihood(sb2_engine *e, double *perLocus, double *total);
/* Both phases with one device round trip.  total3 = {coalescent, likelihood, coalescent + likelihood}
 * (likelihood and sum are still produced when the coalescent is -INFINITY; the caller short-circuits). */
SB2_API int sb2_eval_posterior(sb2_engine *e, double *perLocusCoal, double *perLocusLik, double *perBranch,
                               double *total3);

/* Split form for multi-GPU hosts: begin() uploads staged state and launches everything up to the rank-local
 * partial sums, which land in the reduce vector {coalescent, likelihood, nIncompatible, (q,gamma,logR) x nSpeciesNodes}
 * (MultispeciesCoalescent.java:203-232 needs the per-branch sums over ALL loci before its closed form);
 * the host all-reduces (sum) that device vector across ranks; end() finishes and downloads. */
SB2_API int sb2_eval_begin(sb2_engine *e);
SB2_API int sb2_eval_end(sb2_engine *e, double *perLocusCoal, double *perLocusLik, double *perBranch, double *total3);
/* Device address and length (in doubles) of the reduce vector; bind an external device buffer (e.g. a
 * torch tensor handed to torch.distributed.all_reduce) with sb2_bind_reduce_vector before the first eval. */
SB2_API int sb2_reduce_vector(sb2_engine *e, void **devPtr, int32_t *nDoubles);
SB2_API int sb2_bind_reduce_vector(sb2_engine *e, void *devPtr);
/* One host thread driving several engines (one per GPU, loci sharded): after sb2_eval_begin on
 * every engine, read each engine's reduce vector into host memory (synchronises that engine's stream), add them up, write
 * the sum back into every engine, then sb2_eval_end.  host has sb2_reduce_vector's nDoubles entries. */
SB2_API int sb2_read_reduce_vector(sb2_engine *e, double *host);
SB2_API int sb2_write_reduce_vector(sb2_engine *e, const double *host);
/* Peer-memory all-reduce for one-process-per-GPU hosts on an NVLink / NVSwitch box: after sb2_comm_connect the
 * single-call evaluations (sb2_eval_posterior / _coalescent / _likelihood) exchange the reduce vector inside the
 * reduction kernel (stores into every peer's buffer + sequence flags, summed in rank order: all ranks get
 * bit-identical totals) instead of needing a host-driven all-reduce between begin() and end().
 * sb2_comm_handle writes the 64-byte CUDA IPC handle of this engine's communication buffer; the host all-gathers the
 * handles and passes them, ordered by rank, to sb2_comm_connect on every rank (collective: all ranks must call it,
 * and afterwards every rank must take part in every evaluation). */
SB2_API int sb2_comm_handle(sb2_engine *e, void *handle64);
SB2_API int sb2_comm_connect(sb2_engine *e, int32_t rank, int32_t nRanks, const void *handles);
/* Run on this CUDA stream (cudaStream_t) instead of the engine's own. */
SB2_API int sb2_set_stream(sb2_engine *e, void *cudaStream);

/* CalculationNode.store / restore / accept (GeneTree.java:79-123, MultispeciesCoalescent.java:50-88,
 * StarBeastClock.java:41-52; BeerLikelihoodCore store/restore): double-buffered device state, index flip, no copies
 * of partials. */
SB2_API int sb2_store(sb2_engine *e);
SB2_API int sb2_restore(sb2_engine *e);
SB2_API int sb2_accept(sb2_engine *e);
/* Marks every locus and model dirty so that the next eval recomputes everything from the device-resident state
 * (what BEAST's periodic from-scratch consistency check does; also the bench's "all nodes dirty" regime).  Nothing cached
 * counts: the memo of the analytic per-branch closed forms is bypassed (and refilled) by that evaluation too. */
SB2_API int sb2_mark_all_dirty(sb2_engine *e);

/* ---- read-back for the plugin classes' getters and for parity tests --------------------------------------- */

/* GeneTree.getSpeciesOccupancy-derived StarBeastClock.getRateForBranch values, [2*nTips-1]. */
SB2_API int sb2_get_branch_rates(sb2_engine *e, int32_t locus, double *rates);
/* GeneTree.coalescentLineageCounts / coalescentCounts [nSpeciesNodes], geneNodeSpeciesAssignment [2*nTips-1],
 * and per-branch GeneTree perBranchLogP [nSpeciesNodes]; any pointer may be NULL. */
SB2_API int sb2_get_embedding(sb2_engine *e, int32_t locus, int32_t *lineageCounts, int32_t *coalCounts,
                              int32_t *assignment, double *perBranchLogP);
/* Current partials of an internal node, [nCategories][nPatterns][4], scaled by 2^-exponent
 * (exponents[nCategories][nPatterns], cumulative over the subtree). */
SB2_API int sb2_get_partials(sb2_engine *e, int32_t locus, int32_t node, double *partials, int32_t *exponents);
/* Current transition matrices of the branch above `node`, [nCategories][16]. */
SB2_API int sb2_get_matrices(sb2_engine *e, int32_t locus, int32_t node, double *matrices);
/* patternLogLikelihoods of the last evaluation of this locus, [nPatterns]. */
SB2_API int sb2_get_pattern_logl(sb2_engine *e, int32_t locus, double *patternLogL);

/* ---- operator-side queries over the resident gene trees ---------------------------------- */
/* Both read the LAST EVALUATED (or restored) state, i.e. the chain's current state when an operator runs;
 * SB2_ERR_STATE if a gene tree is staged but not yet evaluated.  With loci sharded over several engines the caller
 * combines the per-engine answers with min().
 *
 * CoordinatedUniform.getConnectingNodes + findConnectingNodes (CoordinatedUniform.java:94-177) and the
 * CoordinatedExponential twins (CoordinatedExponential.java:92-160), for every gene tree at once:
 *   mask[sum of (2*nTips-1)]  1 where the gene node belongs to a connected component of `speciesNode` (loci concatenated in
 *                             engine order, node-number order inside a locus; likelihood-only loci added with
 *                             sb2_add_locus_k have no gene tree: their slots stay 0), may be NULL;
 *   freedoms[0] tipwardFreedom, freedoms[1] rootwardFreedom as accumulated by the gene-tree walk (+inf if unconstrained;
 *   the caller adds the species-branch constraints of :76-81; CoordinatedExponential ignores freedoms[1]);
 *   count       number of connecting nodes, may be NULL. */
SB2_API int sb2_connecting_nodes(sb2_engine *e, int32_t speciesNode, uint8_t *mask, double *freedoms, int64_t *count);
/* NodeReheight2.recalculateMaxHeight (NodeReheight2.java:150-203) without the origin: leafIsRight[species leaf] =
 * (canonicalMap[leaf] >= centerIndex); *maxHeight = height of the lowest gene node joining a pure-left with a pure-right
 * subtree over all gene trees (+inf if none). */
SB2_API int sb2_max_height(sb2_engine *e, const uint8_t *leafIsRight, double *maxHeight);

/* ---- alignment -> site patterns (engine-independent) ---------------------------------------- */
/* BEAST.base Alignment.calcPatterns (the <data> blocks of the reference's XML, e.g. tutorial/CanisPhylogeny/
 * Canis-Phylogeny.xml:3-23): sequences[tip][site] (0-3, >= 4 missing / ambiguous) -> the unique site columns in
 * lexicographic order (taxon 0 most significant) as patterns[tip][*nPatterns] (compact rows), their multiplicities
 * weights[*nPatterns] and, optionally, siteToPattern[nSites].  Buffers must hold nSites entries per row / array.
 * Feed the result to sb2_add_locus.  Alignments longer than 16384 sites are sorted in blocks of 16384 columns on the device
 * and the blocks' pattern lists merged on the host (same result).  Errors: sb2_compress_last_error(). */
SB2_API int sb2_compress_alignment(int32_t device, int32_t nTips, int32_t nSites, const uint8_t *sequences, uint8_t *patterns,
                                   int32_t *weights, int32_t *siteToPattern, int32_t *nPatterns);
SB2_API const char *sb2_compress_last_error(void);

/* ---- introspection ------------------------------------------------------------------------------------------ */

SB2_API int32_t sb2_num_loci(const sb2_engine *e);
/* kernels launched by this engine since creation */
SB2_API int64_t sb2_launch_count(const sb2_engine *e);
/* number of (node) pruning operations scheduled by the last evaluation, summed over loci */
SB2_API int64_t sb2_last_node_updates(sb2_engine *e);
/* profile=1: accumulated device time (ms) and launch count of the pruning kernel since the last call */
SB2_API int sb2_kernel_time(sb2_engine *e, double *ms, int64_t *launches);
/* switch the CUDA-event bracketing of the pruning kernel on/off (events between launches defeat programmatic
 * dependent launch, so latency measurements run with profiling off) */
SB2_API int sb2_set_profile(sb2_engine *e, int32_t on);
/* profile=1, after an evaluation that went through the fused one-launch step: microseconds since the start of that launch at
 * its phase boundaries, read on the device (%globaltimer) by the leader CTA of the first cluster; n <= 16:
 * [0] start (0), [1] tree-shaped work done, [2] walk done, [3] chunk sums gathered, [4] reductions entered, [5] result published,
 * [8] state loaded, [9] embedding done, [10] coalescent terms done, [11] dirtiness known, [12] schedule built, [13] matrices done */
SB2_API int sb2_step_phase_times(sb2_engine *e, double *us, int32_t n);
/* device bytes held by the engine */
SB2_API int64_t sb2_device_bytes(const sb2_engine *e);
/* after sb2_comm_connect: microseconds the last evaluation's reduction kernel spent waiting for the peers' words
 * (rank skew + one NVLink trip), measured on the device with %globaltimer; 0 without peers */
SB2_API double sb2_exchange_wait_us(const sb2_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* SB2GPU_H */
