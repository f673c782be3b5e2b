// Shared device-side types of libsb2gpu (B200 / sm_100a).
//
// HBM layout (one engine = one GPU = one shard of loci), all arrays locus-major:
//   tipStates   u8  [locus][tip][stateStride]                  static; rows padded to 128 B with 'missing'
//   partials    f64 [locus][buf 0/1][internal node][cat][patStride][4] (BEAST layout per node, rows padded to 64 patterns, two buffers
//                                                               per node for store/restore, index flip only)
//   cumExp      i16 [locus][buf][internal node][cat][patStride] power-of-two scale exponent, cumulative over subtree
//   mats        f64 [locus][buf][node][cat][16]                P = exp(Q d) per branch x category
//   StateBlock (x2: current / stored)  small per-node and per-locus state, see below
//   ops         Op  [locus][internal node]                      post-order pruning schedule built on device
//   opMats      f64 [locus][op][cat][side][16]                  the scheduled ops' matrix pairs, in schedule order
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb2 {

constexpr int SUBST_STRIDE = 64;     // 4 states: HKY {kappa, pi[4]} or EIGEN {eval[4], evec[16], ievc[16], pi[4]} (40 doubles);
                                     // K states (likelihood-only loci, kstate.cu): EIGEN {eval[K], evec[K*K], ievc[K*K], pi[K]}, K <= 5
constexpr int KSTATE_MAX = 5;

struct LocusDesc {
    int32_t nTips, nNodes, nInt, nPat, nCat;
    int32_t nodeBase;   // into per-node arrays
    int32_t tipBase;    // into tipSpecies
    int32_t patBase;    // into weights / patternLogL / constMask
    int32_t catBase;    // into catRates / catProps
    int32_t tileBase;   // first pattern tile
    int32_t nTiles;
    int32_t opBase;     // into ops (sum of nInt)
    int32_t stackDepth; // floor(log2(nInt)) + 1
    int32_t stateStride;// bytes per tip row of tipStates: nPat rounded up to 128, padding = 4 (missing)
    int32_t chunkBase;  // into chunkSum: first 32-pattern chunk of this locus
    int32_t nChunks;    // ceil(nPat / 32)
    // K-state likelihood-only locus (kStates != 0): a TreeLikelihood that is not paired with a GeneTree -- the morphological
    // partitions on the species tree, tutorial/CanisFBD/Canis-FBD.xml:964-1033.  No embedding, no coalescent term, no
    // pruning schedule: k_kstate evaluates it from scratch whenever it is flagged.
    int32_t kStates;    // 0, or the state count 2..KSTATE_MAX
    int32_t nExcl;      // ascertained (excluded) patterns
    int32_t exclBase;   // into the excluded-pattern index array
    int32_t patStride;  // patterns per (node, category) row of partials / cumExp: nPat rounded up to 64, so that every row starts on a
                        // 128-byte line in both arrays (a misaligned scattered row costs 14-50 % of HBM write bandwidth, measured)
    int64_t kBase;      // into the K-state scratch: [internal node][category][pattern][K] doubles (exponents: the same / K)
    int64_t stateBase;  // into tipStates (bytes)
    int64_t partBase;   // into partials (doubles)
    int64_t expBase;    // into cumExp (int16)
    int64_t matBase;    // into mats (doubles)
    int64_t treeOff;    // into the incoming tree staging buffer (bytes)
    double ploidy;
};

// per-locus model parameters (staged on host, uploaded when dirty)
struct LocusParams {
    double subst[SUBST_STRIDE];
    double clockRate;
    double pInv;
    int32_t substKind;
    int32_t clockKind;
};

// per-eval per-locus flags
enum : uint8_t {
    LF_RUN = 1,        // something relevant changed: rebuild embedding, rates and schedule
    LF_TREE_IN = 2,    // a new gene tree sits in the incoming buffer
    LF_FORCE_MATS = 4, // substitution / site model changed: every branch matrix is dirty
    LF_FORCE_ALL = 8,  // first evaluation / mark_all_dirty: everything dirty
};

// operand kinds of a pruning op
// TIP: tip states; PREV: the result of the previous op of the walk, still in registers (every result is consumed
// exactly once, by its parent, and in post-order the parent follows its last-evaluated child immediately);
// STACK: a result parked in the shared-memory stack (the first-evaluated child of a node with two dirty children);
// GLOBAL: clean node, partials resident in HBM from an earlier evaluation; SPILL: parked result whose slot is deeper
// than the shared-memory stack -> re-read from the global buffer this launch has just written
// REMOTE: (latency-bound shapes only, see WALK_MAX_WORKERS) the result of a side subtree that another worker warp of the
// same CTA evaluates concurrently; it arrives in that warp's mailbox slot (slot field) and the operand's offset field holds
// the worker number
enum : uint8_t { OPK_TIP = 0, OPK_STACK = 1, OPK_GLOBAL = 2, OPK_SPILL = 3, OPK_PREV = 4, OPK_REMOTE = 5 };
constexpr unsigned OP_NO_SLOT = 255;   // outSlot: result is not parked (forwarded in registers, or spilled)
constexpr unsigned OP_SIGNAL = 1u << 22;   // ctl: the parked result is awaited by another worker: publish the mailbox flag
// Tree-level parallelism for latency-bound shapes: the walk of one pattern chunk is split over up to WALK_MAX_WORKERS warps.
// Worker 0 walks the main (heavy-child) path; side subtrees hanging off it go to the helpers, whose results come back through
// WALK_MAILBOX dedicated stack slots per helper (slots stackDepth .. stackDepth + WALK_MAILBOX - 1).
constexpr int WALK_MAX_WORKERS = 4;
constexpr int WALK_MAILBOX = 2;
constexpr int WALK_MAX_WORKER_OPS = 64;    // multi-worker walks keep the whole op list in shared memory

struct __align__(16) Op {
    // offsets are pre-multiplied by k_prepare so that the pruning kernel needs one IMAD.WIDE per address:
    uint32_t outOff;  // (buf*nInt + internal index) * nCat*patStride -> (pattern, category) element of the output
    uint32_t aOff;    // TIP: tip index; GLOBAL / SPILL: (buf*nInt + internal index) * nCat*patStride
    uint32_t bOff;
    uint32_t ctl;     // bits 0-2 aKind, 3-5 bKind, 6-9 aSlot, 10-13 bSlot, 14-21 outSlot (>= slot count: not parked), 22 OP_SIGNAL
};
static_assert(sizeof(Op) == 16, "Op must stay 16 bytes (one LDS.128)");
__host__ __device__ inline uint32_t op_ctl(unsigned aKind, unsigned bKind, unsigned aSlot, unsigned bSlot, unsigned outSlot)
{
    return aKind | (bKind << 3) | (aSlot << 6) | (bSlot << 10) | (outSlot << 14);
}

// Double-buffered small state (current / stored).  restore() swaps the two blocks, store() copies cur->stored.
struct StateBlock {
    // by node (sum over loci of nNodes)
    int32_t *gParent, *gLeft, *gRight;
    double *gHeight;
    double *bRate;      // branch rate per node (branchRateModel.getRateForBranch)
    double *bTime;      // length*rate the current matrices were computed for (TreeLikelihood.m_branchLengths)
    uint8_t *curP;      // which of the two partials buffers is current, per node
    uint8_t *curM;      // which of the two matrix buffers is current, per node
    // by locus
    double *logL;       // TreeLikelihood logP
    double *coal;       // GeneTree logP
    // by (locus, species node)
    double *brLogP;     // GeneTree.perBranchLogP
    double *anG;        // analytic: gamma_j / ploidy_j
    double *anR;        // analytic: k_j * log(ploidy_j)
    int32_t *brN, *brK; // coalescentLineageCounts / coalescentCounts
    // by node
    int32_t *assign;    // geneNodeSpeciesAssignment
    // species tree + shared parameters
    int32_t *sParent, *sLeft, *sRight;
    double *sHeight, *sRates, *popA, *popB;
    double *hyper;      // {alpha, beta}
};

struct PrepareArgs {
    const LocusDesc *loci;
    const LocusParams *params;
    const uint8_t *flags;        // [nLoci]
    const uint8_t *inTree;       // incoming gene trees: per locus {height f64[G], parent, left, right i32[G]}
    const int32_t *tipSpecies;
    const double *catRates;
    const double *explicitRates; // by node
    StateBlock cur;
    StateBlock stored;
    Op *ops;
    int32_t *nOps;               // [nLoci]
    double *mats;
    double *opMats;              // [op][category][side][16]: each scheduled op's two matrices, in schedule order
    int32_t S, nSpeciesLeaves, popKind, nLoci;
    int32_t stackDepth;          // shared-memory stack slots of k_treewalk
    int32_t uniformFlags;        // >= 0: every locus uses these flags (no per-step flags upload); < 0: read `flags`
    // lazy store(): loci whose slice of the `stored` block must first be brought up to `cur` (they changed in the previous
    // step); each such locus's CTA copies its own slices before anything else.  nSync == 0: nothing owed (or the host did a
    // whole-block copy)
    int32_t nSync;
    int32_t syncLoci[8];
    int32_t singleLocus;         // >= 0 (and uniformFlags < 0): only this locus has non-zero flags, `singleFlags` (the common MCMC step)
    int32_t singleFlags;
    // static inputs of k_treewalk, only prefetched into L2 here (that kernel runs next and would otherwise chase them
    // through DRAM one dependent miss at a time whenever L2 is cold)
    const uint8_t *pfTipStates;
    const int32_t *pfWeights, *pfTileLocus;
    const uint8_t *pfConstMask;
    const double *pfCatProps;
    int32_t workers;             // worker warps per pattern chunk of k_treewalk (1 = single walk)
    int32_t *laneOps;            // [nLoci][WALK_MAX_WORKERS + 1] first op of each worker's list (workers > 1)
};

struct ReduceArgs {
    const LocusDesc *loci;
    const int32_t *nOps;
    const double *chunkSum;   // [sum of nChunks] weighted log-likelihood of every 32-pattern chunk
    const double *patLogL;
    const int32_t *weights;
    StateBlock cur;
    double *red;        // reduce vector {coal, lik, nIncompat, (q, gamma, logR) x S}
    double *out;        // {total3, perLocusCoal[L], perLocusLik[L], perBranch[S]}
    int32_t nLoci, S, popKind, exact;
    int32_t useTiles;   // 0: pruning results are stale or pending -> use the cached per-locus logL
    double *hostOut;    // host-mapped copy of `out` (zero-copy download), or nullptr
    // Polled download (latency path): hostTag != 0 -> every double of the result goes to hostWords[2 i], [2 i + 1] as two
    // 8-byte words {tag | 32 data bits} instead of into hostOut.  An aligned 8-byte store lands atomically, so a host that sees
    // this evaluation's tag in a word also sees its data: no system-scope fence in the kernel, no wait for the kernel to
    // drain on the host (the flag-in-data idea of the peer-memory all-reduce, pointed at the host).
    unsigned long long *hostWords;
    unsigned long long hostTag;
    // Analytic mode: a two-entry memo per species branch of the closed form MultispeciesCoalescent.java:227-232, keyed by the VALUES
    // it is a function of -- {q, gamma, logR, alpha, beta} -> term.  The reference caches a branch's term while nothing it
    // depends on changes (MultispeciesCoalescent.java:165-195); a memo keyed by value gives the same saving and, being a pure
    // function's cache, needs no store / restore bookkeeping: after a rejected proposal the restored state's key is simply the
    // other entry.  16 doubles per branch: {key[5], term} x 2, next victim; nullptr = off.
    double *memo;
    int32_t memoFlush;   // 1: sb2_mark_all_dirty -- "recompute everything": no memo hit counts this time (entries are refilled)
    int32_t pad2;
};
constexpr int MEMO_STRIDE = 16;

struct WalkArgs {
    const LocusDesc *loci;
    const LocusParams *params;
    const int32_t *tileLocus;    // [nTiles]
    const Op *ops;
    const int32_t *nOps;
    const uint8_t *tipStates;
    const int32_t *weights;
    const uint8_t *constMask;
    const double *catProps;
    const double *opMats;
    double *partials;
    int16_t *cumExp;
    double *patLogL;             // by pattern
    double *chunkSum;            // by 32-pattern chunk (LocusDesc::chunkBase)
    int32_t nCat;                // categories of every locus in this launch
    int32_t chunksPerTile;       // pattern chunks per tile (block = 32 * nCat * chunksPerTile threads)
    int32_t patPerThread;        // R: patterns per thread (a chunk is 32*R patterns)
    int32_t stackDepth;          // shared-memory stack slots; deeper results spill to their global buffer
    int32_t maxTips;             // largest tip count over loci (sizes the shared-memory tip-state tile)
    int32_t tipsInSmem;          // 1: tip states of the tile are staged in shared memory
    int32_t workers;             // worker warps per pattern chunk (tree-level parallelism), 1 = off
    int32_t debugMaxOps;         // timing experiments only (SB2_DEBUG_MAX_OPS): walk at most this many ops per worker; < 0 = all
    const int32_t *laneOps;      // [nLoci][WALK_MAX_WORKERS + 1]
    int32_t ctasPerSM;           // > 0: cap on resident CTAs per SM (see engine.cu: walk_ctas_per_sm)
    unsigned long long *dbgTimes; // profiling only (SB2_DEBUG_WALK_TIMES): [tile][8] %globaltimer stamps of warp 0 (start, staged, dependency met,
                                 // walk begins, walk ends, CTA ends), else nullptr
};


// The fused one-launch step (step.cu): one thread-block cluster per locus, one CTA of STEP_WARPS warps per pattern tile
constexpr int STEP_THREADS = 256;       // == PREP_THREADS: the tree-shaped part is prepare_body.cuh's, warp-specialised in two halves
constexpr int STEP_WARPS = STEP_THREADS / 32;
constexpr int STEP_MAX_CLUSTER = 16;    // non-portable cluster size limit of sm_100
constexpr size_t STEP_DYNAMIC_SMEM_MAX = (227 - 24) * 1024;   // k_step also holds ~20 KB of STATIC shared memory (reduction scratch, memo lists)
constexpr int STEP_INLINE_TREE_MAX = 3072;   // bytes of gene tree (20 per node) that fit the kernel arguments next to the rest

// Peer-memory all-reduce of the reduce vector (one process per GPU, NVLink / NVSwitch; see reduce.cu:k_reduce_xgpu)
constexpr int COMM_MAX_RANKS = 16;
struct CommArgs {
    unsigned long long *peerSlots[COMM_MAX_RANKS];  // every rank's communication buffer: slots [2][COMM_MAX_RANKS][nRed][2] words,
                                                    // each word = {32 bits of the double, 32-bit sequence flag} (one 8-byte store)
    unsigned long long *peerFlags[COMM_MAX_RANKS];  // (unused since the flags travel inside the data words)
    int32_t rank, nRanks, nRed;
    uint32_t pad;
    unsigned long long seq;                         // evaluation sequence number (parity selects the buffer half)
};
cudaError_t launch_reduce_xgpu(const ReduceArgs &a, const CommArgs &c, cudaStream_t st, int64_t *launches);

struct StepArgs {
    PrepareArgs prep;            // ops / opMats / laneOps / pf* unused: schedule and matrices stay in shared memory
    ReduceArgs red;              // useTiles = 0: the leaders write the per-locus values into the state block
    CommArgs comm;
    const uint8_t *tipStates;
    const int32_t *weights;
    const uint8_t *constMask;
    const double *catProps;
    double *partials;
    int16_t *cumExp;
    double *patLogL;
    uint32_t *doneCounter;       // leaders finished so far (self-resetting), all-loci mode
    int32_t nRun;                // > 0: only clusters runLoci[0..nRun) exist; [0] runs with runFlags, the others only owe the lazy
                                 // store() copy.  0: one cluster per locus, flags as in PrepareArgs
    int32_t runLoci[9];
    int32_t runFlags;
    int32_t maxNodes, nCat, stackDepth, maxTips, useComm;
    unsigned long long *dbg;     // profile mode: %globaltimer stamps of cluster 0's leader at the phase boundaries, else nullptr
    // single-locus steps: the staged gene tree of runLoci[0] rides in the kernel arguments (no PCIe round trip to the mapped
    // host buffer, no copy-engine hop on the latency path); inlineTreeBytes == 0: read it from prep.inTree as usual
    int32_t inlineTreeBytes;
    int32_t pad0;
    unsigned char inlineTree[STEP_INLINE_TREE_MAX];
};
constexpr int STEP_STAMPS = 16;
size_t step_smem_bytes(int maxNodes, int S, int nCat, int stackDepth, int maxTips);
cudaError_t launch_step(const StepArgs &a, int nClusters, int clusterSize, cudaStream_t st);

// K-state likelihood-only loci (kstate.cu)
struct KStateArgs {
    const LocusDesc *loci;
    const LocusParams *params;
    const int32_t *kLoci;        // [grid] locus index of each K-state locus
    const uint8_t *flags;
    int32_t uniformFlags, singleLocus, singleFlags;
    const uint8_t *tipStates;
    const int32_t *weights, *excluded;
    const double *catRates, *catProps, *explicitRates;
    StateBlock cur;
    double *patLogL;
    double *scratch;             // partials, see LocusDesc::kBase
    int32_t *scratchExp;         // power-of-two exponents per (node, category, pattern)
    int32_t maxNodes, maxCat;
    size_t smemScratch;          // bytes of shared memory set aside for a locus's partials (loci that do not fit use `scratch`)
};
size_t kstate_smem_bytes(int maxNodes, int maxCat, int K, size_t scratchBytes);
size_t kstate_scratch_bytes(int nInt, int nCat, int nPat, int K);
cudaError_t launch_kstate(const KStateArgs &a, int nK, int maxK, cudaStream_t st);

// Operator-side queries over the resident gene trees (query.cu)
constexpr int QUERY_THREADS = 128;
enum : int32_t { QUERY_CONNECTING = 0, QUERY_MAX_HEIGHT = 1 };
struct QueryArgs {
    const LocusDesc *loci;
    const int32_t *tipSpecies;
    const uint8_t *leafClass;        // [nSpeciesLeaves] class bits of each species leaf: 1, 2 or 4
    const int32_t *gParent, *gLeft, *gRight;
    const double *gHeight;
    uint8_t *mask;                   // [total nodes] connecting-node flags, or nullptr
    unsigned long long *result;      // {tipward, rootward, maxHeight} order-preserving encodings (pre-set to enc(+inf)), count
    int32_t mode, maxNodes;
};
cudaError_t launch_tip_class_query(const QueryArgs &a, int nLoci, cudaStream_t st);

// Every launcher returns the launch's cudaError_t (cudaErrorInvalidValue for a geometry no instantiation covers) -- the
// library lives inside a JVM, so nothing in here may abort() or drop an error.
cudaError_t launch_prepare(const PrepareArgs &a, int maxNodes, cudaStream_t st);
cudaError_t launch_treewalk(const WalkArgs &a, int nTiles, bool exact, cudaStream_t st);
cudaError_t launch_reduce_local(const ReduceArgs &a, cudaStream_t st, int64_t *launches);
cudaError_t launch_finish(const ReduceArgs &a, cudaStream_t st, int64_t *launches);
cudaError_t launch_reduce_finish(const ReduceArgs &a, cudaStream_t st, int64_t *launches);
// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: one host thread may drive one engine per GPU
// (INTEGRATION.md section 4), so the opt-in above 48 KB is cached per (kernel instantiation, device), never per process.
constexpr int MAX_DEVICES = 64;
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kernel, size_t smem, size_t (&configured)[MAX_DEVICES + 1])
{
    // the 48 KB a launch gets without opting in cover the kernel's STATIC shared memory too
    // (cached behind the per-device entries of the call site's array, as bytes + 1: the template is instantiated per kernel
    // TYPE, and kernels of one signature share it)
    if (configured[MAX_DEVICES] == 0) {
        cudaFuncAttributes fa;
        const cudaError_t fst = cudaFuncGetAttributes(&fa, kernel);
        if (fst != cudaSuccess) return fst;
        configured[MAX_DEVICES] = fa.sharedSizeBytes + 1;
    }
    if (smem + (configured[MAX_DEVICES] - 1) <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaError_t st = cudaGetDevice(&dev);
    if (st != cudaSuccess) return st;
    if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
    if (smem <= configured[dev]) return cudaSuccess;
    st = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st == cudaSuccess) configured[dev] = smem;
    return st;
}
size_t prepare_smem_bytes(int maxNodes, int S);
size_t treewalk_smem_bytes(int nCat, int chunksPerTile, int R, int stackDepth, int maxTips, int workers);
bool treewalk_tips_in_smem(int maxTips, int tilePat);

}  // namespace sb2
