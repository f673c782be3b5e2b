// Host side of libsb2gpu.so: the engine object behind include/sb2gpu.h.
//
// The engine is thin on purpose: the host stages changed inputs in pinned memory and marks loci dirty; every
// decision that depends on tree state (what is dirty, in which order to prune) is taken on the device by
// k_prepare, so one evaluation is  H2D(staged ranges) -> k_prepare -> k_treewalk -> k_reduce_local -> [all-reduce]
// -> k_finish -> D2H(one small vector).  store/restore/accept are a D2D copy / pointer swap of the small state block.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/sb2gpu.h"
#include "sb2_device.cuh"

using namespace sb2;

namespace {

struct HostLocus {
    LocusDesc d;
    std::vector<uint8_t> states;
    std::vector<int32_t> weights;
    std::vector<int32_t> tipSpecies;
    std::vector<uint8_t> constMask;
    std::vector<int32_t> excluded;   // K-state loci: ascertained pattern indices
};

}  // namespace

struct sb2_engine {
    sb2_config cfg{};
    std::string err;
    bool finalized = false;
    cudaStream_t stream = nullptr;
    // K-state likelihood-only loci run on a side stream, concurrently with the 4-state pruning kernel (they only depend on
    // k_prepare); every reduction -- and anything else that touches their inputs or outputs -- joins it first (join_kstate)
    cudaStream_t kStream = nullptr;
    cudaEvent_t evPrepared = nullptr, evKDone = nullptr;
    bool kPending = false;
    bool ownStream = false;
    std::vector<HostLocus> loci;
    int S = 0, nSpeciesLeaves = 0, L = 0;
    int64_t totalNodes = 0, totalTips = 0, totalPat = 0, totalCat = 0, totalInt = 0, totalTiles = 0;
    int maxNodes = 0, nCat = 0, chunksPerTile = 1, patPerThread = 1, stackDepth = 1, workers = 1;
    int32_t *dLaneOps = nullptr;
    // K-state likelihood-only loci (kstate.cu)
    int nK = 0, maxK = 0;
    size_t kSmemScratch = 0;         // shared-memory partials scratch of k_kstate: the largest K-state locus that fits 96 KB
    std::vector<int32_t> kLoci;
    int32_t *dKLoci = nullptr, *dExcluded = nullptr, *dKExp = nullptr;
    double *dKScratch = nullptr;
    // fused one-launch step (step.cu)
    bool fusedOk = false;            // shapes fit: cluster <= 16 CTAs per locus, shared memory, category count
    int fusedMode = -1;              // SB2_FUSED: 0 never, 1 whenever eligible, -1 auto (single-locus steps; all-loci steps of small problems)
    int64_t fusedMaxLanes = 0;       // auto: all-loci evaluations go through k_step up to this many (pattern, category) lanes; 0 = never:
                                     // measured on B200, the three PDL-chained kernels win when every locus runs (C2: 37.7 vs 43.4 us per step)
    int64_t totalLanes = 0;
    int fusedCluster = 1;            // CTAs per cluster when every locus runs (widest locus)
    unsigned long long *hStamps = nullptr, *hStampsDev = nullptr;   // profile mode: phase stamps of k_step (mapped host memory)
    int lastFusedSingle = -1;        // >= 0: the last evaluation ran this one locus through k_step (only its nOps entry is fresh)
    int64_t devBytes = 0;
    int64_t launches = 0;
    std::vector<void *> devAllocs;

    // static device data
    LocusDesc *dLoci = nullptr;
    uint8_t *dStates = nullptr, *dConstMask = nullptr;
    int32_t *dWeights = nullptr, *dTipSpecies = nullptr, *dTileLocus = nullptr;
    // big dynamic device data
    double *dPartials = nullptr, *dMats = nullptr, *dOpMats = nullptr, *dPatLogL = nullptr, *dChunkSum = nullptr;
    int16_t *dCumExp = nullptr;
    Op *dOps = nullptr;
    unsigned long long *dDbgTimes = nullptr;   // SB2_DEBUG_WALK_TIMES: per-tile stamps of the pruning kernel (profiling only)
    double *dMemo = nullptr;         // analytic closed forms, two-entry memo per species branch (ReduceArgs::memo)
    int32_t *dNOps = nullptr;
    // staged inputs: pinned host mirror + device copy
    LocusParams *hParams = nullptr, *dParams = nullptr;
    double *hCat = nullptr, *dCat = nullptr;          // [2][totalCat]: rates then props
    double *hExplicit = nullptr, *dExplicit = nullptr;  // by node (allocated lazily on first EXPLICIT clock)
    uint8_t *hInTree = nullptr, *dInTree = nullptr, *hInTreeDev = nullptr;   // hInTreeDev: device address of the mapped host buffer
    int64_t inTreeBytes = 0;
    uint8_t *hFlags = nullptr, *dFlags = nullptr;
    // species + shared parameters, staged as one contiguous byte block that mirrors the StateBlock's species region
    uint8_t *hSpecies = nullptr;
    size_t speciesBytes = 0, speciesOffInBlock = 0;
    // state blocks
    uint8_t *blockMem[2] = {nullptr, nullptr};
    size_t blockBytes = 0;
    StateBlock blk[2];
    int curIdx = 0;
    // outputs
    double *dRed = nullptr, *dRedOwn = nullptr, *dOut = nullptr, *hOut = nullptr, *hOutDev = nullptr;
    uint32_t *dDone = nullptr;
    // polled download of the result vector (ReduceArgs::hostTag): mapped host words, the tag of the evaluation in flight
    unsigned long long *hWords = nullptr, *hWordsDev = nullptr;
    bool pollOk = false;
    uint32_t pollSeq = 0;
    unsigned long long pollTag = 0;   // != 0: the evaluation in flight publishes tagged words
    // peer-memory all-reduce
    void *commBuf = nullptr;
    size_t commSlotsBytes = 0;
    std::vector<void *> commOpened;
    CommArgs comm{};
    bool commConnected = false;
    int nRed = 0, nOut = 0;
    int popKind = SB2_POP_CONSTANT;
    // dirtiness
    std::vector<uint8_t> treeDirty, modelDirty, clockDirty;
    bool speciesDirty = true, catDirty = true, paramsDirty = true, explicitDirty = false, forceAll = true;
    bool memoFlush = false;          // sb2_mark_all_dirty: the next closed forms are all recomputed (see ReduceArgs::memoFlush)
    bool needPrepare = true, walkPending = false, tilesValid = false, reduced = false;
    // lazy store(): the physical copy cur -> stored is owed until the next k_prepare, which does it for the few loci that
    // differ (`diffLoci`: changed since the blocks were last equal) -- or the host copies the whole block when everything differs
    bool storePending = false, diffAll = true;
    std::vector<int32_t> diffLoci;
    std::vector<uint8_t> diffFlag;
    bool resultValid = false;   // hOut holds a full evaluation (coalescent AND likelihood) of exactly the staged state
    // host mirrors of the stored state (restore() must bring the staging copies back too)
    LocusParams *hParamsStored = nullptr;
    double *hCatStored = nullptr, *hExplicitStored = nullptr;
    uint8_t *hSpeciesStored = nullptr;
    int popKindStored = SB2_POP_CONSTANT;
    std::vector<int32_t> touched;        // loci whose parameters changed since the last store()
    std::vector<uint8_t> touchedFlag;
    int reuploadFirst = 0x7fffffff, reuploadLast = -1;   // parameter range to re-upload after a restore()
    // operator-side queries (allocated on first use): {result u64[4], leafClass[pad], mask[totalNodes]}
    uint8_t *dQuery = nullptr, *hQuery = nullptr;
    size_t queryLeafPad = 0;
    bool everPrepared = false;
    bool hasStored = false;              // sb2_store has been called at least once (sb2_restore needs a stored state)
    std::vector<int32_t> scratchStack;   // check_tree
    // profiling
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evPool;
    size_t evUsed = 0;

    int fail(int code, const char *fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define CUDA_TRY(e, call)                                                                        \
    do {                                                                                         \
        cudaError_t _st = (call);                                                                \
        if (_st != cudaSuccess) return (e)->fail(SB2_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_st)); \
    } while (0)

namespace {

template <typename T>
int dev_alloc(sb2_engine *e, T **p, size_t n)
{
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    cudaError_t st = cudaMalloc((void **)p, bytes);
    if (st != cudaSuccess) return e->fail(SB2_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(st));
    e->devAllocs.push_back(*p);
    e->devBytes += (int64_t)bytes;
    return SB2_OK;
}

template <typename T>
int pin_alloc(sb2_engine *e, T **p, size_t n)
{
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    cudaError_t st = cudaHostAlloc((void **)p, bytes, cudaHostAllocDefault);
    if (st != cudaSuccess) return e->fail(SB2_ERR_NOMEM, "cudaHostAlloc(%zu bytes): %s", bytes, cudaGetErrorString(st));
    memset(*p, 0, bytes);
    return SB2_OK;
}

// carve a StateBlock out of one allocation; returns bytes needed.  Species region is contiguous.
size_t carve_block(StateBlock *b, uint8_t *base, int64_t N, int64_t L, int S, int nLeaves, size_t *speciesOff, size_t *speciesBytes)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) & ~size_t(15); return r; };
    const size_t LS = (size_t)L * S;
    size_t gHeight = take(8 * N), bRate = take(8 * N), bTime = take(8 * N);
    size_t logL = take(8 * L), coal = take(8 * L);
    size_t brLogP = take(8 * LS), anG = take(8 * LS), anR = take(8 * LS);
    size_t gParent = take(4 * N), gLeft = take(4 * N), gRight = take(4 * N), assign = take(4 * N);
    size_t brN = take(4 * LS), brK = take(4 * LS);
    size_t curP = take(N), curM = take(N);
    // species region (kept contiguous: one H2D copy)
    const size_t sp0 = o;
    size_t sHeight = take(8 * S), sRates = take(8 * S), popA = take(8 * S), popB = take(8 * S), hyper = take(16);
    size_t sParent = take(4 * S), sLeft = take(4 * S), sRight = take(4 * S);
    const size_t sp1 = o;
    (void)nLeaves;
    if (speciesOff) *speciesOff = sp0;
    if (speciesBytes) *speciesBytes = sp1 - sp0;
    if (b) {
        b->gHeight = (double *)(base + gHeight); b->bRate = (double *)(base + bRate); b->bTime = (double *)(base + bTime);
        b->logL = (double *)(base + logL); b->coal = (double *)(base + coal);
        b->brLogP = (double *)(base + brLogP); b->anG = (double *)(base + anG); b->anR = (double *)(base + anR);
        b->gParent = (int32_t *)(base + gParent); b->gLeft = (int32_t *)(base + gLeft); b->gRight = (int32_t *)(base + gRight);
        b->assign = (int32_t *)(base + assign); b->brN = (int32_t *)(base + brN); b->brK = (int32_t *)(base + brK);
        b->curP = base + curP; b->curM = base + curM;
        b->sHeight = (double *)(base + sHeight); b->sRates = (double *)(base + sRates);
        b->popA = (double *)(base + popA); b->popB = (double *)(base + popB); b->hyper = (double *)(base + hyper);
        b->sParent = (int32_t *)(base + sParent); b->sLeft = (int32_t *)(base + sLeft); b->sRight = (int32_t *)(base + sRight);
    }
    return o;
}

int floor_log2(int x)
{
    int r = 0;
    while ((1 << (r + 1)) <= x) r++;
    return r;
}

// host view of the pinned species staging block (same carve as the device block's species region)
struct SpeciesView {
    double *sHeight, *sRates, *popA, *popB, *hyper;
    int32_t *sParent, *sLeft, *sRight;
};
SpeciesView species_view(sb2_engine *e)
{
    StateBlock tmp;
    uint8_t *base = reinterpret_cast<uint8_t *>(reinterpret_cast<uintptr_t>(e->hSpecies) - e->speciesOffInBlock);
    carve_block(&tmp, base, e->totalNodes, e->L, e->S, e->nSpeciesLeaves, nullptr, nullptr);
    SpeciesView v;
    v.sHeight = tmp.sHeight; v.sRates = tmp.sRates; v.popA = tmp.popA; v.popB = tmp.popB; v.hyper = tmp.hyper;
    v.sParent = tmp.sParent; v.sLeft = tmp.sLeft; v.sRight = tmp.sRight;
    return v;
}

int check_locus(sb2_engine *e, int32_t locus)
{
    if (!e->finalized) return e->fail(SB2_ERR_STATE, "engine not finalized");
    if (locus < 0 || locus >= e->L) return e->fail(SB2_ERR_ARG, "locus %d out of range [0,%d)", locus, e->L);
    return SB2_OK;
}

void touch(sb2_engine *e, int32_t locus)
{
    if (!e->touchedFlag[locus]) { e->touchedFlag[locus] = 1; e->touched.push_back(locus); }
}

StateBlock &cur(sb2_engine *e) { return e->blk[e->curIdx]; }
StateBlock &stored(sb2_engine *e) { return e->blk[1 - e->curIdx]; }

ReduceArgs reduce_args(sb2_engine *e, bool coalOnly = false)
{
    ReduceArgs r;
    r.useTiles = (!coalOnly && e->tilesValid) ? 1 : 0;
    r.loci = e->dLoci; r.nOps = e->dNOps; r.chunkSum = e->dChunkSum; r.patLogL = e->dPatLogL; r.weights = e->dWeights;
    r.cur = cur(e); r.red = e->dRed; r.out = e->dOut;
    r.nLoci = e->L; r.S = e->S; r.popKind = e->popKind; r.exact = e->cfg.exact_order;
    r.hostOut = e->hOutDev;
    r.hostWords = e->hWordsDev; r.hostTag = 0;
    r.memo = e->dMemo; r.memoFlush = e->memoFlush ? 1 : 0; r.pad2 = 0;
    return r;
}

// a FULL evaluation (every entry of the result vector is rewritten) may publish tagged words that the host polls
void arm_poll(sb2_engine *e, ReduceArgs &r)
{
    e->pollTag = 0;
    // (with peers connected too: the words are written after the exchange; a rank whose host has its result can be at most one
    // evaluation ahead of a peer, which the exchange's alternating buffer halves already allow for)
    static const bool pollUnderComm = getenv("SB2_NO_POLL_COMM") == nullptr;   // A/B knob
    if (!e->pollOk || (e->commConnected && !pollUnderComm)) return;
    e->pollSeq++;
    e->pollTag = (0x80000000ull | (e->pollSeq & 0x7fffffffull)) << 32;
    r.hostTag = e->pollTag;
}

int do_walk(sb2_engine *e);
int flush_pending_walk(sb2_engine *e);

// Structural check of a host-supplied binary tree in BEAST node numbering (leaves first): exactly one root, every
// internal node has two distinct children that name it as their parent, leaves have none, and a walk from the root
// reaches all n nodes.  Together these rule out cycles and detached components, on which the device-side rootward
// walks (k_prepare, k_tip_class_query) and the host-side descents would never terminate.  nLeaves < 0: leaves are
// recognised by left < 0 (species tree), otherwise nodes 0..nLeaves-1 must be the leaves (gene trees).
// Returns nullptr when the tree is sound, else what is wrong.  `stack` is scratch.
const char *check_tree(int n, int nLeaves, const int32_t *parent, const int32_t *left, const int32_t *right, std::vector<int32_t> &stack)
{
    int root = -1, roots = 0;
    for (int i = 0; i < n; i++) {
        if (parent[i] >= n) return "node index out of range";
        if (parent[i] < 0) { root = i; roots++; }
    }
    if (roots != 1) return "tree must have exactly one root";
    stack.clear();
    stack.push_back(root);
    int seen = 0;
    while (!stack.empty()) {
        const int i = stack.back();
        stack.pop_back();
        seen++;
        const int L = left[i], R = right[i];
        const bool leaf = nLeaves < 0 ? (L < 0) : (i < nLeaves);
        if (leaf) {
            if (L >= 0 || R >= 0) return nLeaves < 0 ? "left and right must both be -1 at a leaf" : "nodes 0..tips-1 must be the leaves";
            continue;
        }
        if (L < 0 || R < 0) return nLeaves < 0 ? "left and right must both be -1 at a leaf" : "nodes 0..tips-1 must be the leaves";
        if (L >= n || R >= n) return "node index out of range";
        if (L == R || parent[L] != i || parent[R] != i) return "parent / child arrays disagree";
        stack.push_back(L);
        stack.push_back(R);
    }
    if (seen != n) return "tree is not connected (cycle or detached nodes)";
    return nullptr;
}

// upload staged inputs and run k_prepare -- or, when the caller wants a whole evaluation (allowFused) and the shapes are
// latency-bound, the fused one-launch step k_step (*didFused: prepare, walk, reductions and publish are all in flight)
int join_kstate(sb2_engine *e)
{
    if (e->kPending) {
        CUDA_TRY(e, cudaStreamWaitEvent(e->stream, e->evKDone, 0));
        e->kPending = false;
    }
    return SB2_OK;
}

int do_prepare(sb2_engine *e, bool allowFused = false, bool *didFused = nullptr)
{
    { int jrc = join_kstate(e); if (jrc) return jrc; }   // a K-state kernel still in flight reads what is about to be uploaded
    if (didFused) *didFused = false;
    if (!e->needPrepare) return SB2_OK;
    // A walk scheduled by an earlier k_prepare (strict two-phase protocol: coalescent evaluated, likelihood never asked for)
    // must run before the next k_prepare re-schedules: that one would treat the already-flipped nodes as clean.
    if (e->walkPending) { int rc = flush_pending_walk(e); if (rc) return rc; }
    e->resultValid = false;
    cudaStream_t st = e->stream;
    const int L = e->L;
    int firstTree = L, lastTree = -1, firstPar = L, lastPar = -1;
    bool any = false;
    for (int l = 0; l < L; l++) {
        uint8_t f = 0;
        if (e->forceAll) f |= LF_RUN | LF_FORCE_ALL;
        if (e->speciesDirty) f |= LF_RUN;
        if (e->treeDirty[l]) { f |= LF_RUN | LF_TREE_IN; firstTree = std::min(firstTree, l); lastTree = l; }
        if (e->modelDirty[l]) { f |= LF_RUN | LF_FORCE_MATS; firstPar = std::min(firstPar, l); lastPar = l; }
        if (e->clockDirty[l]) { f |= LF_RUN; firstPar = std::min(firstPar, l); lastPar = l; }
        e->hFlags[l] = f;
        any |= f != 0;
    }
    int uniformFlags = e->hFlags[0];
    for (int l = 1; l < L; l++) if (e->hFlags[l] != e->hFlags[0]) { uniformFlags = -1; break; }
    // (a) the copy owed by store(): bring the stored block up to the current one where they differ, before this evaluation
    //     changes anything -- a handful of loci: their CTAs do it (kernel arguments); otherwise one whole-block copy
    int nSync = 0;
    int32_t syncLoci[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    static const bool eagerStore = getenv("SB2_EAGER_STORE") != nullptr;   // A/B knob: whole-block copy every time
    if (e->storePending) {
        if (e->diffAll || e->diffLoci.size() > 8 || eagerStore) {
            CUDA_TRY(e, cudaMemcpyAsync(e->blockMem[1 - e->curIdx], e->blockMem[e->curIdx], e->blockBytes, cudaMemcpyDeviceToDevice, st));
        } else {
            for (int32_t l : e->diffLoci) syncLoci[nSync++] = l;
        }
        for (int32_t l : e->diffLoci) e->diffFlag[l] = 0;
        e->diffLoci.clear(); e->diffAll = false; e->storePending = false;
    }
    // (b) what this evaluation is about to change
    if (e->forceAll || e->speciesDirty || e->catDirty || e->paramsDirty || e->explicitDirty || uniformFlags > 0) e->diffAll = true;
    else for (int l = 0; l < L; l++) if (e->hFlags[l] && !e->diffFlag[l]) { e->diffFlag[l] = 1; e->diffLoci.push_back(l); }
    // the common MCMC step touches one locus: its index and flags ride in the kernel arguments, no flags upload
    int singleLocus = -1, nFlagged = 0;
    if (uniformFlags < 0)
        for (int l = 0; l < L; l++) if (e->hFlags[l]) { nFlagged++; singleLocus = l; }
    if (nFlagged != 1) singleLocus = -1;
    if (uniformFlags < 0 && singleLocus < 0) CUDA_TRY(e, cudaMemcpyAsync(e->dFlags, e->hFlags, L, cudaMemcpyHostToDevice, st));
    if (e->speciesDirty)
        CUDA_TRY(e, cudaMemcpyAsync(e->blockMem[e->curIdx] + e->speciesOffInBlock, e->hSpecies, e->speciesBytes, cudaMemcpyHostToDevice, st));
    if (e->reuploadLast >= 0) { firstPar = std::min(firstPar, e->reuploadFirst); lastPar = std::max(lastPar, e->reuploadLast); }
    e->reuploadFirst = 0x7fffffff; e->reuploadLast = -1;
    if (lastPar >= 0 || e->paramsDirty) {
        if (e->paramsDirty) { firstPar = 0; lastPar = L - 1; }
        CUDA_TRY(e, cudaMemcpyAsync(e->dParams + firstPar, e->hParams + firstPar, sizeof(LocusParams) * (size_t)(lastPar - firstPar + 1), cudaMemcpyHostToDevice, st));
    }
    if (e->catDirty) CUDA_TRY(e, cudaMemcpyAsync(e->dCat, e->hCat, sizeof(double) * 2 * (size_t)e->totalCat, cudaMemcpyHostToDevice, st));
    if (e->explicitDirty && e->hExplicit)
        CUDA_TRY(e, cudaMemcpyAsync(e->dExplicit, e->hExplicit, sizeof(double) * (size_t)e->totalNodes, cudaMemcpyHostToDevice, st));
    // Uploads of up to 256 KB of trees ride zero-copy (each CTA reads its locus's tree from the mapped host buffer: one
    // coalesced burst per CTA, measured -10 us end to end at C2); bulk uploads go through the copy engine (zero-copy reads
    // of 1.2 MB measured +40 us at the C5 shard).  The small species block always goes through the copy engine: having
    // every CTA fetch it over PCIe measured slower than one 1 KB memcpy.
    static const bool noZeroCopy = getenv("SB2_NO_ZEROCOPY") != nullptr;
    const bool zeroCopyTrees = !noZeroCopy && lastTree >= 0 &&
        (e->loci[lastTree].d.treeOff + 20ll * e->loci[lastTree].d.nNodes - e->loci[firstTree].d.treeOff) <= (256ll << 10);
    if (lastTree >= 0 && !zeroCopyTrees) {
        const int64_t b0 = e->loci[firstTree].d.treeOff;
        const int64_t b1 = e->loci[lastTree].d.treeOff + 20ll * e->loci[lastTree].d.nNodes;
        CUDA_TRY(e, cudaMemcpyAsync(e->dInTree + b0, e->hInTree + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, st));
    }
    PrepareArgs a;
    a.loci = e->dLoci; a.params = e->dParams; a.flags = e->dFlags; a.inTree = zeroCopyTrees ? e->hInTreeDev : e->dInTree; a.tipSpecies = e->dTipSpecies;
    a.catRates = e->dCat; a.explicitRates = e->dExplicit ? e->dExplicit : e->dCat;
    a.cur = cur(e); a.stored = stored(e); a.ops = e->dOps; a.nOps = e->dNOps; a.mats = e->dMats; a.opMats = e->dOpMats;
    a.S = e->S; a.nSpeciesLeaves = e->nSpeciesLeaves; a.popKind = e->popKind; a.nLoci = L; a.stackDepth = e->stackDepth; a.uniformFlags = uniformFlags;
    a.workers = e->workers; a.laneOps = e->dLaneOps;
    a.singleLocus = singleLocus; a.singleFlags = singleLocus >= 0 ? e->hFlags[singleLocus] : 0;
    a.nSync = nSync;
    for (int j = 0; j < 8; j++) a.syncLoci[j] = syncLoci[j];
    a.pfTipStates = e->dStates; a.pfWeights = e->dWeights; a.pfTileLocus = e->dTileLocus; a.pfConstMask = e->dConstMask; a.pfCatProps = e->dCat + e->totalCat;
    // ---- one launch or three? ------------------------------------------------------------------------------------------
    // k_step serves single-locus steps (the MCMC's common move) and, for small problems, steps in which every locus runs;
    // streaming-size evaluations keep the three-kernel path, whose tile CTAs do not repeat the tree-shaped work.
    bool fused = false;
    int runCluster = e->fusedCluster;
    if (allowFused && e->fusedOk && e->fusedMode != 0 && !e->cfg.exact_order) {
        if (singleLocus >= 0) {
            fused = e->loci[singleLocus].d.kStates == 0;   // a K-state locus goes through k_prepare + k_kstate
            const LocusDesc &d = e->loci[singleLocus].d;
            const int tilePat = (STEP_WARPS / e->nCat) * 32;
            runCluster = (d.nPat + tilePat - 1) / tilePat;
        } else {
            fused = e->nK == 0 && (e->fusedMode == 1 || e->totalLanes <= e->fusedMaxLanes);
        }
    }
    e->lastFusedSingle = -1;
    if (fused) {
        StepArgs sa;
        sa.prep = a;
        sa.prep.workers = 1;
        sa.red = reduce_args(e, /*coalOnly=*/true);   // useTiles = 0: the leaders write the per-locus values into the state block
        arm_poll(e, sa.red);
        sa.comm = e->comm;
        sa.useComm = e->commConnected ? 1 : 0;
        if (e->commConnected) { e->comm.seq++; sa.comm.seq = e->comm.seq; }
        sa.tipStates = e->dStates; sa.weights = e->dWeights; sa.constMask = e->dConstMask; sa.catProps = e->dCat + e->totalCat;
        sa.partials = e->dPartials; sa.cumExp = e->dCumExp; sa.patLogL = e->dPatLogL; sa.doneCounter = e->dDone;
        sa.maxNodes = e->maxNodes; sa.nCat = e->nCat; sa.stackDepth = e->stackDepth; sa.maxTips = (e->maxNodes + 1) / 2;
        sa.dbg = nullptr;
        if (e->cfg.profile) {
            if (!e->hStamps) {
                CUDA_TRY(e, cudaHostAlloc((void **)&e->hStamps, sizeof(unsigned long long) * STEP_STAMPS, cudaHostAllocMapped));
                memset(e->hStamps, 0, sizeof(unsigned long long) * STEP_STAMPS);
                CUDA_TRY(e, cudaHostGetDevicePointer((void **)&e->hStampsDev, e->hStamps, 0));
            }
            sa.dbg = e->hStampsDev;
        }
        sa.nRun = 0; sa.runFlags = 0; sa.inlineTreeBytes = 0; sa.pad0 = 0;
        for (int j = 0; j < 9; j++) sa.runLoci[j] = 0;
        int nClusters = L;
        if (singleLocus >= 0) {
            sa.runLoci[0] = singleLocus; sa.runFlags = e->hFlags[singleLocus]; sa.nRun = 1;
            for (int j = 0; j < nSync; j++) if (syncLoci[j] != singleLocus) sa.runLoci[sa.nRun++] = syncLoci[j];
            nClusters = sa.nRun;
            e->lastFusedSingle = singleLocus;
            const LocusDesc &d1 = e->loci[singleLocus].d;
            if ((sa.runFlags & LF_TREE_IN) && 20 * d1.nNodes <= STEP_INLINE_TREE_MAX) {   // the staged tree rides in the arguments
                sa.inlineTreeBytes = 20 * d1.nNodes;
                memcpy(sa.inlineTree, e->hInTree + d1.treeOff, (size_t)sa.inlineTreeBytes);
            }
        }
        std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
        if (e->cfg.profile) {
            if (e->evUsed == e->evPool.size()) {
                cudaEvent_t x, y;
                CUDA_TRY(e, cudaEventCreate(&x));
                CUDA_TRY(e, cudaEventCreate(&y));
                e->evPool.push_back({x, y});
            }
            ev = e->evPool[e->evUsed++];
            CUDA_TRY(e, cudaEventRecord(ev.first, st));
        }
        CUDA_TRY(e, launch_step(sa, nClusters, runCluster, st));
        e->memoFlush = false;
        if (e->cfg.profile) CUDA_TRY(e, cudaEventRecord(ev.second, st));
    } else {
        CUDA_TRY(e, launch_prepare(a, e->maxNodes, st));
        bool kFlagged = false;
        for (int32_t l : e->kLoci) kFlagged |= (e->hFlags[l] & LF_RUN) != 0;
        if (kFlagged) {   // K-state likelihood-only loci: from scratch, on the tree k_prepare has just made current
            KStateArgs k;
            k.loci = e->dLoci; k.params = e->dParams; k.kLoci = e->dKLoci; k.flags = e->dFlags;
            k.uniformFlags = uniformFlags; k.singleLocus = singleLocus; k.singleFlags = a.singleFlags;
            k.tipStates = e->dStates; k.weights = e->dWeights; k.excluded = e->dExcluded;
            k.catRates = e->dCat; k.catProps = e->dCat + e->totalCat; k.explicitRates = e->dExplicit ? e->dExplicit : e->dCat;
            k.cur = cur(e); k.patLogL = e->dPatLogL; k.scratch = e->dKScratch; k.scratchExp = e->dKExp;
            k.maxNodes = e->maxNodes; k.maxCat = e->nCat; k.smemScratch = e->kSmemScratch;
            if (e->kStream) {   // concurrently with the pruning kernel of the gene loci
                CUDA_TRY(e, cudaEventRecord(e->evPrepared, st));
                CUDA_TRY(e, cudaStreamWaitEvent(e->kStream, e->evPrepared, 0));
                CUDA_TRY(e, launch_kstate(k, e->nK, e->maxK, e->kStream));
                CUDA_TRY(e, cudaEventRecord(e->evKDone, e->kStream));
                e->kPending = true;
            } else {
                CUDA_TRY(e, launch_kstate(k, e->nK, e->maxK, st));
            }
            e->launches++;
        }
    }
    if (didFused) *didFused = fused;
    e->launches++;
    std::fill(e->treeDirty.begin(), e->treeDirty.end(), 0);
    std::fill(e->modelDirty.begin(), e->modelDirty.end(), 0);
    std::fill(e->clockDirty.begin(), e->clockDirty.end(), 0);
    e->speciesDirty = e->catDirty = e->paramsDirty = e->explicitDirty = e->forceAll = false;
    e->needPrepare = false;
    e->everPrepared = true;
    e->walkPending = !fused;
    e->tilesValid = false;
    (void)any;
    return SB2_OK;
}

// fuse: 0 = plain; 1 = last CTA also runs k_reduce_local's body; 2 = ... and k_finish's + publish.  *fused tells the
// caller whether the launch happened (nothing is launched when no walk is pending).
// final reduction of the single-call evaluations: across GPUs over peer memory when connected, else local
int launch_final(sb2_engine *e, const ReduceArgs &r)
{
    { int jrc = join_kstate(e); if (jrc) return jrc; }
    static const bool skip = getenv("SB2_DEBUG_SKIP_REDUCE") != nullptr;   // timing experiments only: results are stale
    if (skip) { e->pollTag = 0; return SB2_OK; }
    if (e->commConnected) {
        e->comm.seq++;
        CUDA_TRY(e, launch_reduce_xgpu(r, e->comm, e->stream, &e->launches));
    } else {
        CUDA_TRY(e, launch_reduce_finish(r, e->stream, &e->launches));
    }
    e->memoFlush = false;
    return SB2_OK;
}

int do_walk(sb2_engine *e)
{
    if (!e->walkPending) return SB2_OK;
    WalkArgs a;
    a.loci = e->dLoci; a.params = e->dParams; a.tileLocus = e->dTileLocus; a.ops = e->dOps; a.nOps = e->dNOps;
    a.tipStates = e->dStates; a.weights = e->dWeights; a.constMask = e->dConstMask; a.catProps = e->dCat + e->totalCat;
    a.opMats = e->dOpMats; a.partials = e->dPartials; a.cumExp = e->dCumExp; a.patLogL = e->dPatLogL; a.chunkSum = e->dChunkSum;
    a.nCat = e->nCat; a.chunksPerTile = e->chunksPerTile; a.patPerThread = e->patPerThread; a.stackDepth = e->stackDepth;
    a.maxTips = (e->maxNodes + 1) / 2;
    a.workers = e->workers; a.laneOps = e->dLaneOps;
    static const int walkCtas = getenv("SB2_WALK_CTAS") ? atoi(getenv("SB2_WALK_CTAS")) : 0;   // A/B knob
    a.ctasPerSM = walkCtas;
    static const int debugMaxOps = getenv("SB2_DEBUG_MAX_OPS") ? atoi(getenv("SB2_DEBUG_MAX_OPS")) : -1;
    a.debugMaxOps = debugMaxOps;
    e->walkPending = false; e->tilesValid = true;
    a.tipsInSmem = treewalk_tips_in_smem(a.maxTips, 32 * e->chunksPerTile * e->patPerThread) ? 1 : 0;
    if (a.tipsInSmem && getenv("SB2_TIP_STAGE_LEGACY")) a.tipsInSmem = 2;
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    if (e->cfg.profile) {
        if (e->evUsed == e->evPool.size()) {
            cudaEvent_t x, y;
            CUDA_TRY(e, cudaEventCreate(&x));
            CUDA_TRY(e, cudaEventCreate(&y));
            e->evPool.push_back({x, y});
        }
        ev = e->evPool[e->evUsed++];
        CUDA_TRY(e, cudaEventRecord(ev.first, e->stream));
    }
    static const char *dbgTimesPath = getenv("SB2_DEBUG_WALK_TIMES");   // profiling only: per-tile %globaltimer stamps -> this file
    a.dbgTimes = nullptr;
    if (dbgTimesPath && e->totalTiles > 0) {
        if (!e->dDbgTimes) { int rc = dev_alloc(e, &e->dDbgTimes, 8 * (size_t)e->totalTiles); if (rc) return rc; }   // the engine's own: freed with it
        CUDA_TRY(e, cudaMemsetAsync(e->dDbgTimes, 0, 64 * (size_t)e->totalTiles, e->stream));
        a.dbgTimes = e->dDbgTimes;
    }
    static const bool skipWalk = getenv("SB2_DEBUG_SKIP_WALK") != nullptr;   // timing experiments only
    if (e->totalTiles == 0) return SB2_OK;   // K-state likelihood-only loci only: nothing for the 4-state pruning kernel
    if (!skipWalk) CUDA_TRY(e, launch_treewalk(a, (int)e->totalTiles, e->cfg.exact_order != 0, e->stream));
    if (dbgTimesPath && a.dbgTimes) {
        std::vector<unsigned long long> h(8 * (size_t)e->totalTiles);
        cudaStreamSynchronize(e->stream);
        cudaMemcpy(h.data(), e->dDbgTimes, 64 * (size_t)e->totalTiles, cudaMemcpyDeviceToHost);
        if (FILE *f = fopen(dbgTimesPath, "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
    }
    e->launches++;
    if (e->cfg.profile) CUDA_TRY(e, cudaEventRecord(ev.second, e->stream));
    CUDA_TRY(e, cudaGetLastError());
    return SB2_OK;
}

int flush_pending_walk(sb2_engine *e)
{
    // k_prepare already flipped buffer indices for the scheduled nodes; make sure their partials exist before the
    // state is frozen (only reachable when a host evaluates the coalescent but never the likelihood)
    if (e->walkPending) {
        int rc = do_walk(e);
        if (rc) return rc;
        // ... and that the cached per-locus log-likelihoods (cur.logL, what a clean locus reports) belong to them
        { int jrc = join_kstate(e); if (jrc) return jrc; }
        CUDA_TRY(e, launch_reduce_local(reduce_args(e), e->stream, &e->launches));
    }
    return SB2_OK;
}

int download(sb2_engine *e)
{
    if (e->pollTag) {
        // the final kernel writes every double of the result as two tagged 8-byte words into mapped host memory: spin on the
        // words themselves (no wait for the kernel to drain, no completion interrupt); the stream is queried now and then so
        // that a failed launch surfaces as an error instead of a hang
        const unsigned long long tag = e->pollTag;
        e->pollTag = 0;
        volatile unsigned long long *w = e->hWords;
        const int n = e->nOut;
        bool finished = false;
        for (int i = 0; i < n; i++) {
            unsigned long long w0, w1;
            for (unsigned spins = 0;; spins++) {
                w0 = w[2 * i]; w1 = w[2 * i + 1];
                if ((w0 & 0xffffffff00000000ull) == tag && (w1 & 0xffffffff00000000ull) == tag) break;
                if ((spins & 1023u) == 1023u) {
                    if (finished) return e->fail(SB2_ERR_CUDA, "result word %d never arrived although the stream is idle", i);
                    const cudaError_t st = cudaStreamQuery(e->stream);
                    if (st == cudaSuccess) finished = true;   // one more look at the word, then give up
                    else if (st != cudaErrorNotReady) return e->fail(SB2_ERR_CUDA, "evaluation failed: %s", cudaGetErrorString(st));
                }
            }
            const unsigned long long bits = (w0 & 0xffffffffull) | (w1 << 32);
            memcpy(&e->hOut[i], &bits, 8);
        }
        return SB2_OK;
    }
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));   // k_finish already published the result vector into host memory
    return SB2_OK;
}

void copy_out(sb2_engine *e, double *perLocusCoal, double *perLocusLik, double *perBranch, double *total3)
{
    const int L = e->L;
    if (total3) memcpy(total3, e->hOut, 3 * sizeof(double));
    if (perLocusCoal) memcpy(perLocusCoal, e->hOut + 3, sizeof(double) * L);
    if (perLocusLik) memcpy(perLocusLik, e->hOut + 3 + L, sizeof(double) * L);
    if (perBranch) memcpy(perBranch, e->hOut + 3 + 2 * L, sizeof(double) * e->S);
}

}  // namespace

// one full evaluation (coalescent + likelihood) into hOut; a no-op when hOut already holds it
static int eval_full(sb2_engine *e)
{
    int rc;
    bool fused = false;
    if ((rc = do_prepare(e, /*allowFused=*/true, &fused))) return rc;
    // cached result: only without peers -- with a connected communicator EVERY evaluation call is one exchange on every
    // rank, whatever this rank's local state is (a rank whose loci did not change must still contribute)
    if (!fused && e->resultValid && !e->walkPending && !e->commConnected) return SB2_OK;
    if (!fused) {
        if ((rc = do_walk(e))) return rc;
        ReduceArgs r = reduce_args(e);
        arm_poll(e, r);
        if ((rc = launch_final(e, r))) return rc;
    }
    if ((rc = download(e))) return rc;
    e->resultValid = true;
    return SB2_OK;
}

extern "C" {

int sb2_create(const sb2_config *cfg, sb2_engine **out)
{
    if (!out) return SB2_ERR_ARG;
    *out = nullptr;
    sb2_engine *e = new sb2_engine();
    if (cfg) e->cfg = *cfg;
    *out = e;   // handed back even on failure so that sb2_last_error works
    int nDev = 0;
    cudaError_t st = cudaGetDeviceCount(&nDev);
    if (st != cudaSuccess || nDev == 0)
        return e->fail(SB2_ERR_CUDA, "no CUDA device (%s); libsb2gpu has no CPU fallback", st != cudaSuccess ? cudaGetErrorString(st) : "device count 0");
    if (e->cfg.device < 0 || e->cfg.device >= nDev) return e->fail(SB2_ERR_ARG, "device %d out of range [0,%d)", e->cfg.device, nDev);
    CUDA_TRY(e, cudaSetDevice(e->cfg.device));
    cudaDeviceProp prop;
    CUDA_TRY(e, cudaGetDeviceProperties(&prop, e->cfg.device));
    if (prop.major != 10) return e->fail(SB2_ERR_CUDA, "device %d is sm_%d%d; libsb2gpu is built for sm_100a only", e->cfg.device, prop.major, prop.minor);
    CUDA_TRY(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->ownStream = true;
    return SB2_OK;
}

void sb2_destroy(sb2_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->kStream) cudaStreamSynchronize(e->kStream);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->kStream) cudaStreamDestroy(e->kStream);
    if (e->evPrepared) cudaEventDestroy(e->evPrepared);
    if (e->evKDone) cudaEventDestroy(e->evKDone);
    for (void *p : e->devAllocs) cudaFree(p);
    for (void *p : {(void *)e->hParams, (void *)e->hCat, (void *)e->hExplicit, (void *)e->hInTree, (void *)e->hFlags, (void *)e->hSpecies, (void *)e->hOut,
                    (void *)e->hStamps, (void *)e->hWords, (void *)e->hParamsStored, (void *)e->hCatStored, (void *)e->hExplicitStored, (void *)e->hSpeciesStored, (void *)e->hQuery})
        if (p) cudaFreeHost(p);
    for (void *p : e->commOpened) cudaIpcCloseMemHandle(p);
    for (auto &ev : e->evPool) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    if (e->ownStream && e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

const char *sb2_last_error(const sb2_engine *e) { return e ? e->err.c_str() : "null engine"; }

int sb2_add_locus(sb2_engine *e, int32_t nTips, int32_t nPatterns, const uint8_t *tipStates, const int32_t *patternWeights,
                  int32_t nCategories, const int32_t *tipToSpecies, double ploidy)
{
    if (e && !tipToSpecies) return e->fail(SB2_ERR_ARG, "sb2_add_locus needs tipToSpecies (likelihood-only loci: sb2_add_locus_k)");
    return sb2_add_locus_k(e, nTips, nPatterns, 4, tipStates, patternWeights, nCategories, tipToSpecies, ploidy, 0, nullptr);
}

int sb2_add_locus_k(sb2_engine *e, int32_t nTips, int32_t nPatterns, int32_t nStates, const uint8_t *tipStates, const int32_t *patternWeights,
                    int32_t nCategories, const int32_t *tipToSpecies, double ploidy, int32_t nExcluded, const int32_t *excludedPatterns)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    if (e->finalized) return e->fail(SB2_ERR_STATE, "sb2_add_locus after sb2_finalize");
    if (nTips < 2 || nPatterns < 1 || nCategories < 1 || nCategories > 32 || !tipStates || !patternWeights)
        return e->fail(SB2_ERR_ARG, "bad locus shape: tips %d patterns %d categories %d (1..32)", nTips, nPatterns, nCategories);
    // The 4-state pruning path pairs a TreeLikelihood with a GeneTree.  Anything else -- another state count, an ascertained
    // alignment, a likelihood whose tree is not a gene tree -- is a K-state likelihood-only locus (kstate.cu).
    const bool kstate = !tipToSpecies;
    if (nStates < 2 || nStates > KSTATE_MAX) return e->fail(SB2_ERR_ARG, "nStates %d outside 2..%d", nStates, KSTATE_MAX);
    if (!kstate && (nStates != 4 || nExcluded != 0))
        return e->fail(SB2_ERR_ARG, "a locus with a gene-tree embedding (tipToSpecies) must be 4-state and not ascertained; K-state / ascertained data are likelihood-only (tipToSpecies = NULL)");
    if (nExcluded < 0 || (nExcluded > 0 && !excludedPatterns)) return e->fail(SB2_ERR_ARG, "bad excluded-pattern list");
    for (int i = 0; i < nExcluded; i++)
        if (excludedPatterns[i] < 0 || excludedPatterns[i] >= nPatterns) return e->fail(SB2_ERR_ARG, "excluded pattern %d out of range", excludedPatterns[i]);
    if (nTips > 8192) return e->fail(SB2_ERR_ARG, "nTips %d > 8192: cumulative int16 exponents could overflow", nTips);
    if (!e->loci.empty() && e->loci[0].d.nCat != nCategories)
        return e->fail(SB2_ERR_ARG, "all loci of one engine must share the category count (%d vs %d)", e->loci[0].d.nCat, nCategories);
    HostLocus hl;
    memset(&hl.d, 0, sizeof hl.d);
    hl.d.nTips = nTips; hl.d.nNodes = 2 * nTips - 1; hl.d.nInt = nTips - 1; hl.d.nPat = nPatterns; hl.d.nCat = nCategories;
    hl.d.ploidy = ploidy;
    hl.d.kStates = kstate ? nStates : 0; hl.d.nExcl = nExcluded;
    hl.states.assign(tipStates, tipStates + (size_t)nTips * nPatterns);
    hl.weights.assign(patternWeights, patternWeights + nPatterns);
    if (tipToSpecies) hl.tipSpecies.assign(tipToSpecies, tipToSpecies + nTips);
    else hl.tipSpecies.assign(nTips, 0);
    if (nExcluded) hl.excluded.assign(excludedPatterns, excludedPatterns + nExcluded);
    // TreeLikelihood.calcConstantPatternIndices (useAmbiguities=false): states shared by every unambiguous tip
    hl.constMask.assign(nPatterns, 0xF);
    for (int t = 0; t < nTips; t++)
        for (int p = 0; p < nPatterns; p++) {
            const uint8_t s = tipStates[(size_t)t * nPatterns + p];
            if (s < 4) hl.constMask[p] &= (uint8_t)(1u << s);
        }
    e->loci.push_back(std::move(hl));
    return (int)e->loci.size() - 1;
}

int sb2_finalize(sb2_engine *e, int32_t nSpeciesNodes, int32_t nSpeciesLeaves)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    if (e->finalized) return e->fail(SB2_ERR_STATE, "already finalized");
    if (e->loci.empty()) return e->fail(SB2_ERR_STATE, "no loci");
    if (nSpeciesNodes < 1 || nSpeciesLeaves < 1 || nSpeciesLeaves > nSpeciesNodes) return e->fail(SB2_ERR_ARG, "bad species tree size");
    e->S = nSpeciesNodes; e->nSpeciesLeaves = nSpeciesLeaves; e->L = (int)e->loci.size();
    const int L = e->L;
    e->nCat = e->loci[0].d.nCat;
    // launch geometry of the pruning kernel: a warp = 32*R patterns of one category, a CTA = nCat * chunks warps.
    // Large problems are instruction-issue bound: R=2 amortises op decode and matrix loads over two patterns;
    // small ones are latency bound and want as many independent warps as possible (R=1).
    {
        int64_t threads1 = 0;
        for (auto &hl : e->loci) if (!hl.d.kStates) threads1 += (int64_t)hl.d.nPat * hl.d.nCat;
        // measured on B200 (same box A/B, round 2 with aligned rows): R=2 wins from ~100k (pattern, category) lanes up
        // (C2 x 200 loci = 112k: 30.7 vs 36.9 us; C5 x 103 = 139k: 122.6 vs 134.6; C5 x 2000: 1,961 vs 2,109), R=1 below
        // (C3 x 50 = 91k: 43.2 vs 44.2; C2 x 100 = 56k: 20.6 vs 21.8; C2 = 28k: 16.7 vs 18.4): small problems are latency
        // bound and want as many independent warps as possible
        e->patPerThread = threads1 > 100000 ? 2 : 1;
        e->chunksPerTile = std::max(1, 2 / e->nCat);
        if (const char *v = getenv("SB2_PATTERNS_PER_THREAD")) e->patPerThread = atoi(v);
        if (const char *v = getenv("SB2_CHUNKS_PER_TILE")) e->chunksPerTile = atoi(v);
        // Tree-level parallelism (worker warps per chunk, see sb2_device.cuh) is OPT-IN (SB2_WORKERS=2|4): measured on B200 at C2
        // (same box): the pruning kernel gains ~1 us (18.8 -> 17.7) because half of its time is fixed cost, not the walk, while
        // k_prepare's worker assignment costs ~4 us -> the step gets slower (50.6 -> 54.9 us).  Kept for very short chains of
        // large trees; results are bit-identical either way (tests/test_gpu_parity.py).
        int maxInt = 0;
        for (auto &hl : e->loci) if (!hl.d.kStates) maxInt = std::max(maxInt, hl.d.nTips - 1);
        e->workers = 1;
        if (const char *v = getenv("SB2_WORKERS")) {
            const int wv = atoi(v);
            const bool ok = wv == 1 || (maxInt <= WALK_MAX_WORKER_OPS && e->patPerThread == 1 && ((e->nCat == 1 && (wv == 2 || wv == 4)) || (e->nCat == 4 && wv == 2)));
            if (!ok) return e->fail(SB2_ERR_ARG, "SB2_WORKERS=%d not available for this shape", wv);
            e->workers = wv;
        }
        if (e->workers > 1) e->chunksPerTile = 1;
        if (e->patPerThread != 1 && e->patPerThread != 2 && e->patPerThread != 4) return e->fail(SB2_ERR_ARG, "SB2_PATTERNS_PER_THREAD must be 1, 2 or 4");
        if (e->chunksPerTile < 1 || 32 * e->nCat * e->chunksPerTile > 256) return e->fail(SB2_ERR_ARG, "nCat * chunks per tile must stay <= 8 warps");
    }
    const int tilePat = 32 * e->chunksPerTile * e->patPerThread;
    int64_t nodeBase = 0, tipBase = 0, patBase = 0, catBase = 0, tileBase = 0, opBase = 0, stateBase = 0, partBase = 0, expBase = 0, matBase = 0, treeOff = 0, chunkBase = 0;
    int64_t kBase = 0, exclBase = 0;
    e->stackDepth = 1;
    for (auto &hl : e->loci) {
        LocusDesc &d = hl.d;
        for (int t = 0; t < d.nTips && !d.kStates; t++)
            if (hl.tipSpecies[t] < 0 || hl.tipSpecies[t] >= nSpeciesLeaves) return e->fail(SB2_ERR_ARG, "tipToSpecies must name a species LEAF (node number below the leaf count)");
        d.nodeBase = (int32_t)nodeBase; d.tipBase = (int32_t)tipBase; d.patBase = (int32_t)patBase; d.catBase = (int32_t)catBase;
        const bool ks = d.kStates != 0;   // K-state likelihood-only locus: no tiles, no schedule, no resident partials
        d.tileBase = (int32_t)tileBase; d.nTiles = ks ? 0 : (d.nPat + tilePat - 1) / tilePat; d.opBase = (int32_t)opBase;
        d.stackDepth = floor_log2(d.nInt) + 1;
        d.chunkBase = (int32_t)chunkBase; d.nChunks = ks ? 0 : (d.nPat + 31) / 32;
        chunkBase += d.nChunks;
        d.stateBase = stateBase; d.partBase = partBase; d.expBase = expBase; d.matBase = matBase; d.treeOff = treeOff;
        if (!ks) e->stackDepth = std::max(e->stackDepth, d.stackDepth);
        e->maxNodes = std::max(e->maxNodes, d.nNodes);
        nodeBase += d.nNodes; tipBase += d.nTips; patBase += d.nPat; catBase += d.nCat; tileBase += d.nTiles; opBase += ks ? 0 : d.nInt;
        d.stateStride = (d.nPat + 127) & ~127;
        d.patStride = (d.nPat + 63) & ~63;
        stateBase += (int64_t)d.nTips * d.stateStride;
        if (ks) {
            d.kBase = kBase; d.exclBase = (int32_t)exclBase;
            kBase += (int64_t)d.nInt * d.nCat * d.nPat * d.kStates;
            kBase = (kBase + 119) / 120 * 120;   // a multiple of every K: the exponent array is indexed by kBase / K
            exclBase += d.nExcl;
            e->kLoci.push_back((int32_t)(&hl - e->loci.data()));
            e->maxK = std::max(e->maxK, d.kStates);
        } else {
            partBase += 2ll * d.nInt * d.nCat * d.patStride * 4;
            expBase += 2ll * d.nInt * d.nCat * d.patStride;
            expBase = (expBase + 63) & ~63ll;
            if (2ll * d.nInt * d.nCat * d.patStride >= 0x7fffffffll) return e->fail(SB2_ERR_ARG, "locus too large for 32-bit element offsets");
            matBase += 2ll * d.nNodes * d.nCat * 16;
        }
        treeOff += 20ll * d.nNodes;
        treeOff = (treeOff + 7) & ~7ll;
    }
    if (nodeBase > 0x7fffffff || patBase > 0x7fffffff) return e->fail(SB2_ERR_ARG, "problem too large for 32-bit node/pattern indices");
    e->totalNodes = nodeBase; e->totalTips = tipBase; e->totalPat = patBase; e->totalCat = catBase; e->totalTiles = tileBase;
    e->totalInt = opBase; e->inTreeBytes = treeOff;
    if (prepare_smem_bytes(e->maxNodes, e->S) > 227 * 1024) return e->fail(SB2_ERR_ARG, "gene tree too large for the per-locus shared-memory kernel");
    e->stackDepth = std::min(e->stackDepth, 4);   // deeper (rare) results are re-read from their global buffer (OPK_SPILL)
    if (const char *v = getenv("SB2_STACK_DEPTH")) e->stackDepth = std::min(15, std::max(1, atoi(v)));
    if (e->workers > 1) e->stackDepth = 3;   // the multi-worker instantiations are compiled for 3 local slots + WALK_MAILBOX
    if (treewalk_smem_bytes(e->nCat, e->chunksPerTile, e->patPerThread, e->stackDepth, (e->maxNodes + 1) / 2, e->workers) > 227 * 1024)
        return e->fail(SB2_ERR_ARG, "pruning stack does not fit shared memory");

    int rc;
    // static data
    std::vector<LocusDesc> descs(L);
    std::vector<uint8_t> states((size_t)stateBase, (uint8_t)4), cmask((size_t)patBase);
    std::vector<int32_t> weights((size_t)patBase), tipsp((size_t)tipBase), tileLocus((size_t)tileBase);
    for (int l = 0; l < L; l++) {
        const HostLocus &hl = e->loci[l];
        descs[l] = hl.d;
        for (int t = 0; t < hl.d.nTips; t++)
            memcpy(states.data() + hl.d.stateBase + (size_t)t * hl.d.stateStride, hl.states.data() + (size_t)t * hl.d.nPat, hl.d.nPat);
        memcpy(cmask.data() + hl.d.patBase, hl.constMask.data(), hl.constMask.size());
        memcpy(weights.data() + hl.d.patBase, hl.weights.data(), hl.weights.size() * 4);
        memcpy(tipsp.data() + hl.d.tipBase, hl.tipSpecies.data(), hl.tipSpecies.size() * 4);
        for (int t = 0; t < hl.d.nTiles; t++) tileLocus[hl.d.tileBase + t] = l;
    }
    if ((rc = dev_alloc(e, &e->dLoci, L))) return rc;
    if ((rc = dev_alloc(e, &e->dStates, states.size()))) return rc;
    if ((rc = dev_alloc(e, &e->dConstMask, cmask.size()))) return rc;
    if ((rc = dev_alloc(e, &e->dWeights, weights.size()))) return rc;
    if ((rc = dev_alloc(e, &e->dTipSpecies, tipsp.size()))) return rc;
    if ((rc = dev_alloc(e, &e->dTileLocus, tileLocus.size()))) return rc;
    CUDA_TRY(e, cudaMemcpy(e->dLoci, descs.data(), sizeof(LocusDesc) * L, cudaMemcpyHostToDevice));
    CUDA_TRY(e, cudaMemcpy(e->dStates, states.data(), states.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(e, cudaMemcpy(e->dConstMask, cmask.data(), cmask.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(e, cudaMemcpy(e->dWeights, weights.data(), weights.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(e, cudaMemcpy(e->dTipSpecies, tipsp.data(), tipsp.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(e, cudaMemcpy(e->dTileLocus, tileLocus.data(), tileLocus.size() * 4, cudaMemcpyHostToDevice));
    for (auto &hl : e->loci) { std::vector<uint8_t>().swap(hl.states); }

    // dynamic data
    if ((rc = dev_alloc(e, &e->dPartials, (size_t)partBase))) return rc;
    if ((rc = dev_alloc(e, &e->dCumExp, (size_t)expBase))) return rc;
    if ((rc = dev_alloc(e, &e->dMats, (size_t)matBase))) return rc;
    if ((rc = dev_alloc(e, &e->dPatLogL, (size_t)patBase))) return rc;
    if ((rc = dev_alloc(e, &e->dChunkSum, (size_t)chunkBase))) return rc;
    if ((rc = dev_alloc(e, &e->dOps, (size_t)opBase))) return rc;
    if ((rc = dev_alloc(e, &e->dOpMats, (size_t)opBase * e->nCat * 32))) return rc;
    if ((rc = dev_alloc(e, &e->dNOps, L))) return rc;
    e->nK = (int)e->kLoci.size();
    if (e->nK && e->totalTiles > 0 && !getenv("SB2_KSTATE_INLINE")) {   // SB2_KSTATE_INLINE: A/B knob, k_kstate in the main stream
        CUDA_TRY(e, cudaStreamCreateWithFlags(&e->kStream, cudaStreamNonBlocking));
        CUDA_TRY(e, cudaEventCreateWithFlags(&e->evPrepared, cudaEventDisableTiming));
        CUDA_TRY(e, cudaEventCreateWithFlags(&e->evKDone, cudaEventDisableTiming));
    }
    if (e->nK) {
        for (int32_t l : e->kLoci) {
            const LocusDesc &d = e->loci[l].d;
            const size_t b = kstate_scratch_bytes(d.nInt, d.nCat, d.nPat, d.kStates);
            if (b <= 96 * 1024) e->kSmemScratch = std::max(e->kSmemScratch, b);
        }
        e->kSmemScratch = (e->kSmemScratch + 15) & ~size_t(15);
        if (kstate_smem_bytes(e->maxNodes, e->nCat, e->maxK, e->kSmemScratch) > 227 * 1024) {
            e->kSmemScratch = 0;   // trees this large leave no room: every K-state locus uses the HBM scratch
            if (kstate_smem_bytes(e->maxNodes, e->nCat, e->maxK, 0) > 227 * 1024) return e->fail(SB2_ERR_ARG, "K-state locus too large for the per-locus shared-memory kernel");
        }
        std::vector<int32_t> excl((size_t)std::max<int64_t>(exclBase, 1), 0);
        for (auto &hl : e->loci) if (hl.d.nExcl) memcpy(excl.data() + hl.d.exclBase, hl.excluded.data(), 4 * (size_t)hl.d.nExcl);
        if ((rc = dev_alloc(e, &e->dKLoci, (size_t)e->nK))) return rc;
        if ((rc = dev_alloc(e, &e->dExcluded, excl.size()))) return rc;
        if ((rc = dev_alloc(e, &e->dKScratch, (size_t)kBase))) return rc;
        if ((rc = dev_alloc(e, &e->dKExp, (size_t)kBase / 2 + 1))) return rc;
        CUDA_TRY(e, cudaMemcpy(e->dKLoci, e->kLoci.data(), 4 * (size_t)e->nK, cudaMemcpyHostToDevice));
        CUDA_TRY(e, cudaMemcpy(e->dExcluded, excl.data(), 4 * excl.size(), cudaMemcpyHostToDevice));
    }
    CUDA_TRY(e, cudaMemset(e->dNOps, 0, sizeof(int32_t) * L));
    CUDA_TRY(e, cudaMemset(e->dCumExp, 0, sizeof(int16_t) * (size_t)std::max<int64_t>(expBase, 1)));
    if ((rc = dev_alloc(e, &e->dParams, L))) return rc;
    if ((rc = pin_alloc(e, &e->hParams, L))) return rc;
    if ((rc = pin_alloc(e, &e->hParamsStored, L))) return rc;
    if ((rc = dev_alloc(e, &e->dCat, 2 * (size_t)catBase))) return rc;
    if ((rc = pin_alloc(e, &e->hCat, 2 * (size_t)catBase))) return rc;
    if ((rc = pin_alloc(e, &e->hCatStored, 2 * (size_t)catBase))) return rc;
    if ((rc = dev_alloc(e, &e->dInTree, (size_t)treeOff))) return rc;
    {   // incoming gene trees: mapped pinned memory, so that k_prepare can read a SMALL proposal's trees straight from the
        // host (zero-copy) instead of waiting for a separate H2D copy on the latency path
        const size_t bytes = std::max<size_t>((size_t)treeOff, 16);
        cudaError_t st = cudaHostAlloc((void **)&e->hInTree, bytes, cudaHostAllocMapped);
        if (st != cudaSuccess) return e->fail(SB2_ERR_NOMEM, "cudaHostAlloc(mapped trees): %s", cudaGetErrorString(st));
        memset(e->hInTree, 0, bytes);
        CUDA_TRY(e, cudaHostGetDevicePointer((void **)&e->hInTreeDev, e->hInTree, 0));
    }
    if ((rc = dev_alloc(e, &e->dFlags, L))) return rc;
    if ((rc = pin_alloc(e, &e->hFlags, L))) return rc;
    for (int l = 0; l < L; l++) {   // defaults: HKY kappa 1 (JC69), one category of rate 1, strict clock rate 1
        LocusParams &p = e->hParams[l];
        p.subst[0] = 1.0; p.subst[1] = p.subst[2] = p.subst[3] = p.subst[4] = 0.25;
        p.clockRate = 1.0; p.pInv = 0.0; p.substKind = SB2_SUBST_HKY; p.clockKind = SB2_CLOCK_STRICT;
    }
    for (int64_t i = 0; i < catBase; i++) { e->hCat[i] = 1.0; e->hCat[catBase + i] = 1.0 / e->nCat; }

    // state blocks
    e->blockBytes = carve_block(nullptr, nullptr, nodeBase, L, e->S, nSpeciesLeaves, &e->speciesOffInBlock, &e->speciesBytes);
    for (int i = 0; i < 2; i++) {
        if ((rc = dev_alloc(e, &e->blockMem[i], e->blockBytes))) return rc;
        CUDA_TRY(e, cudaMemset(e->blockMem[i], 0, e->blockBytes));
        carve_block(&e->blk[i], e->blockMem[i], nodeBase, L, e->S, nSpeciesLeaves, nullptr, nullptr);
    }
    if ((rc = pin_alloc(e, &e->hSpecies, e->speciesBytes))) return rc;
    if ((rc = pin_alloc(e, &e->hSpeciesStored, e->speciesBytes))) return rc;
    {
        SpeciesView v = species_view(e);
        for (int i = 0; i < e->S; i++) { v.sRates[i] = 1.0; v.popA[i] = 1.0; v.popB[i] = 1.0; v.sParent[i] = v.sLeft[i] = v.sRight[i] = -1; }
    }
    // outputs
    e->nRed = 3 + 3 * e->S;
    e->nOut = 3 + 2 * L + e->S;
    if ((rc = dev_alloc(e, &e->dRedOwn, e->nRed))) return rc;
    e->dRed = e->dRedOwn;
    if ((rc = dev_alloc(e, &e->dOut, e->nOut))) return rc;
    if (!getenv("SB2_NO_MEMO")) {   // A/B knob
        if ((rc = dev_alloc(e, &e->dMemo, (size_t)MEMO_STRIDE * e->S))) return rc;
        CUDA_TRY(e, cudaMemset(e->dMemo, 0xFF, sizeof(double) * MEMO_STRIDE * (size_t)e->S));   // keys that match nothing
    }
    {   // result vector: host memory the kernels write directly (zero-copy), so a download is just a stream sync
        // (one extra slot behind the vector: the peer-exchange wait of the last evaluation, sb2_exchange_wait_us)
        cudaError_t st = cudaHostAlloc((void **)&e->hOut, sizeof(double) * (e->nOut + 1), cudaHostAllocMapped);
        if (st != cudaSuccess) return e->fail(SB2_ERR_NOMEM, "cudaHostAlloc(mapped): %s", cudaGetErrorString(st));
        memset(e->hOut, 0, sizeof(double) * (e->nOut + 1));
        CUDA_TRY(e, cudaHostGetDevicePointer((void **)&e->hOutDev, e->hOut, 0));
    }
    {   // polled download: two tagged words per double; small result vectors only (the host checks every word)
        const char *v = getenv("SB2_POLL");
        e->pollOk = (v ? atoi(v) != 0 : true) && e->nOut <= 4096;
        if (e->pollOk) {
            cudaError_t st = cudaHostAlloc((void **)&e->hWords, sizeof(unsigned long long) * 2 * (size_t)e->nOut, cudaHostAllocMapped);
            if (st != cudaSuccess) return e->fail(SB2_ERR_NOMEM, "cudaHostAlloc(mapped words): %s", cudaGetErrorString(st));
            memset(e->hWords, 0, sizeof(unsigned long long) * 2 * (size_t)e->nOut);
            CUDA_TRY(e, cudaHostGetDevicePointer((void **)&e->hWordsDev, e->hWords, 0));
        }
    }
    if ((rc = dev_alloc(e, &e->dLaneOps, (size_t)L * (WALK_MAX_WORKERS + 1)))) return rc;
    CUDA_TRY(e, cudaMemset(e->dLaneOps, 0, sizeof(int32_t) * (size_t)L * (WALK_MAX_WORKERS + 1)));
    if ((rc = dev_alloc(e, &e->dDone, 1))) return rc;
    CUDA_TRY(e, cudaMemset(e->dDone, 0, sizeof(uint32_t)));
    {   // communication buffer: slots [2][COMM_MAX_RANKS][nRed] doubles, then flags [2][COMM_MAX_RANKS] u64 (own cudaMalloc: IPC-exportable)
        e->commSlotsBytes = 2 * sizeof(unsigned long long) * 2 * COMM_MAX_RANKS * (size_t)e->nRed;   // two {32 data bits, 32 flag bits} words per double
        unsigned char *cb = nullptr;
        if ((rc = dev_alloc(e, &cb, e->commSlotsBytes + sizeof(unsigned long long) * 2 * COMM_MAX_RANKS))) return rc;
        e->commBuf = cb;
        CUDA_TRY(e, cudaMemset(cb, 0, e->commSlotsBytes + sizeof(unsigned long long) * 2 * COMM_MAX_RANKS));
    }
    CUDA_TRY(e, cudaMemset(e->dOut, 0, sizeof(double) * e->nOut));

    e->treeDirty.assign(L, 0); e->modelDirty.assign(L, 0); e->clockDirty.assign(L, 0); e->touchedFlag.assign(L, 0);
    e->diffFlag.assign(L, 0);
    memcpy(e->hParamsStored, e->hParams, sizeof(LocusParams) * L);
    memcpy(e->hCatStored, e->hCat, 16 * (size_t)catBase);
    memcpy(e->hSpeciesStored, e->hSpecies, e->speciesBytes);
    {   // the fused one-launch step: eligible when every locus fits a cluster and the CTA's shared memory holds a whole schedule
        const int tilePat = e->nCat <= STEP_WARPS ? (STEP_WARPS / e->nCat) * 32 : 0;
        int maxPat = 0;
        e->totalLanes = 0;
        for (auto &hl : e->loci) if (!hl.d.kStates) { maxPat = std::max(maxPat, hl.d.nPat); e->totalLanes += (int64_t)hl.d.nPat * hl.d.nCat; }
        e->fusedCluster = tilePat ? (maxPat + tilePat - 1) / tilePat : 0;
        e->fusedOk = tilePat > 0 && e->fusedCluster <= STEP_MAX_CLUSTER && e->workers == 1 &&
                     step_smem_bytes(e->maxNodes, e->S, e->nCat, e->stackDepth, (e->maxNodes + 1) / 2) <= STEP_DYNAMIC_SMEM_MAX;
        if (const char *v = getenv("SB2_FUSED")) e->fusedMode = atoi(v) ? 1 : 0;
        if (const char *v = getenv("SB2_FUSED_MAX_LANES")) e->fusedMaxLanes = atoll(v);
    }
    e->finalized = true;
    e->forceAll = true;
    CUDA_TRY(e, cudaDeviceSynchronize());
    return SB2_OK;
}

int sb2_set_species_tree(sb2_engine *e, const int32_t *parent, const int32_t *left, const int32_t *right, const double *height)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!parent || !left || !right || !height) return e->fail(SB2_ERR_ARG, "null species tree array");
    SpeciesView v = species_view(e);
    if (const char *why = check_tree(e->S, -1, parent, left, right, e->scratchStack)) return e->fail(SB2_ERR_ARG, "species tree: %s", why);
    memcpy(v.sParent, parent, 4 * e->S); memcpy(v.sLeft, left, 4 * e->S); memcpy(v.sRight, right, 4 * e->S);
    memcpy(v.sHeight, height, 8 * e->S);
    e->speciesDirty = true; e->needPrepare = true;
    return SB2_OK;
}

int sb2_set_gene_trees(sb2_engine *e, int32_t firstLocus, int32_t nLoci, const int32_t *parent, const int32_t *left,
                       const int32_t *right, const double *height)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (firstLocus < 0 || nLoci < 0 || firstLocus + nLoci > e->L) return e->fail(SB2_ERR_ARG, "locus range [%d,%d) out of range", firstLocus, firstLocus + nLoci);
    if (!parent || !left || !right || !height) return e->fail(SB2_ERR_ARG, "null gene tree array");
    size_t o = 0;
    for (int l = firstLocus; l < firstLocus + nLoci; l++) {
        const LocusDesc &d = e->loci[l].d;
        const int G = d.nNodes;
        if (const char *why = check_tree(G, d.nTips, parent + o, left + o, right + o, e->scratchStack)) return e->fail(SB2_ERR_ARG, "locus %d: gene %s", l, why);
        uint8_t *dst = e->hInTree + d.treeOff;
        memcpy(dst, height + o, 8 * (size_t)G);
        memcpy(dst + 8 * (size_t)G, parent + o, 4 * (size_t)G);
        memcpy(dst + 12 * (size_t)G, left + o, 4 * (size_t)G);
        memcpy(dst + 16 * (size_t)G, right + o, 4 * (size_t)G);
        e->treeDirty[l] = 1;
        o += G;
    }
    e->needPrepare = true;
    return SB2_OK;
}

int sb2_set_gene_tree(sb2_engine *e, int32_t locus, const int32_t *parent, const int32_t *left, const int32_t *right, const double *height)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    return sb2_set_gene_trees(e, locus, 1, parent, left, right, height);
}

int sb2_set_subst_model(sb2_engine *e, int32_t locus, int32_t kind, const double *params)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    int rc = check_locus(e, locus);
    if (rc) return rc;
    if (!params || (kind != SB2_SUBST_HKY && kind != SB2_SUBST_EIGEN)) return e->fail(SB2_ERR_ARG, "bad substitution model");
    const int K = e->loci[locus].d.kStates;
    if (K && kind != SB2_SUBST_EIGEN) return e->fail(SB2_ERR_ARG, "a K-state locus takes an eigen system {eval[K], evec[K*K], ievc[K*K], pi[K]}");
    touch(e, locus);
    LocusParams &p = e->hParams[locus];
    p.substKind = kind;
    memcpy(p.subst, params, sizeof(double) * (kind == SB2_SUBST_HKY ? 5 : (K ? 2 * K + 2 * K * K : 40)));
    e->modelDirty[locus] = 1; e->needPrepare = true;
    return SB2_OK;
}

int sb2_set_site_model(sb2_engine *e, int32_t locus, const double *catRates, const double *catProps, double pInv)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    int rc = check_locus(e, locus);
    if (rc) return rc;
    if (!catRates || !catProps) return e->fail(SB2_ERR_ARG, "null site model array");
    if (e->loci[locus].d.kStates && pInv != 0.0) return e->fail(SB2_ERR_ARG, "proportionInvariant is not supported on K-state loci");
    touch(e, locus);
    const LocusDesc &d = e->loci[locus].d;
    memcpy(e->hCat + d.catBase, catRates, 8 * d.nCat);
    memcpy(e->hCat + e->totalCat + d.catBase, catProps, 8 * d.nCat);
    e->hParams[locus].pInv = pInv;
    e->modelDirty[locus] = 1; e->catDirty = true; e->needPrepare = true;
    return SB2_OK;
}

int sb2_set_clock(sb2_engine *e, int32_t locus, int32_t kind, double meanRate, const double *ratesOrNull)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    int rc = check_locus(e, locus);
    if (rc) return rc;
    if (kind < 0 || kind > 2) return e->fail(SB2_ERR_ARG, "bad clock kind %d", kind);
    if (kind == SB2_CLOCK_SPECIES_RATES && e->loci[locus].d.kStates)
        return e->fail(SB2_ERR_ARG, "StarBeastClock needs a gene-tree embedding: a likelihood-only locus takes a strict or explicit clock");
    if (kind == SB2_CLOCK_EXPLICIT) {
        if (!ratesOrNull) return e->fail(SB2_ERR_ARG, "explicit clock needs rates");
        if (!e->hExplicit) {
            if ((rc = pin_alloc(e, &e->hExplicit, (size_t)e->totalNodes))) return rc;
            if ((rc = pin_alloc(e, &e->hExplicitStored, (size_t)e->totalNodes))) return rc;
            if ((rc = dev_alloc(e, &e->dExplicit, (size_t)e->totalNodes))) return rc;
            for (int64_t i = 0; i < e->totalNodes; i++) e->hExplicit[i] = e->hExplicitStored[i] = 1.0;
        }
    }
    touch(e, locus);
    if (kind == SB2_CLOCK_EXPLICIT) {
        const LocusDesc &d = e->loci[locus].d;
        memcpy(e->hExplicit + d.nodeBase, ratesOrNull, 8 * (size_t)d.nNodes);
        e->explicitDirty = true;
    }
    e->hParams[locus].clockKind = kind;
    e->hParams[locus].clockRate = meanRate;
    e->clockDirty[locus] = 1; e->needPrepare = true;
    return SB2_OK;
}

int sb2_set_species_rates(sb2_engine *e, const double *rates)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!rates) return e->fail(SB2_ERR_ARG, "null rates");
    memcpy(species_view(e).sRates, rates, 8 * e->S);
    e->speciesDirty = true; e->needPrepare = true;
    return SB2_OK;
}

int sb2_set_pop_model(sb2_engine *e, int32_t kind, const double *a, const double *b, double alpha, double beta)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    SpeciesView v = species_view(e);
    if (kind == SB2_POP_CONSTANT) {
        if (!a) return e->fail(SB2_ERR_ARG, "ConstantPopulations needs populationSizes");
        memcpy(v.popA, a, 8 * e->S);
    } else if (kind == SB2_POP_LWCR) {
        if (!a || !b) return e->fail(SB2_ERR_ARG, "LinearWithConstantRoot needs tip and top population sizes");
        memcpy(v.popA, a, 8 * e->nSpeciesLeaves);
        memcpy(v.popB, b, 8 * (e->S - 1));
    } else if (kind == SB2_POP_ANALYTIC) {
        v.hyper[0] = alpha; v.hyper[1] = beta;
    } else {
        return e->fail(SB2_ERR_ARG, "bad population model kind %d", kind);
    }
    e->popKind = kind;
    e->speciesDirty = true; e->needPrepare = true;
    return SB2_OK;
}

int sb2_eval_begin(sb2_engine *e)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    int rc;
    if ((rc = do_prepare(e))) return rc;
    if ((rc = do_walk(e))) return rc;
    { int jrc = join_kstate(e); if (jrc) return jrc; }
    CUDA_TRY(e, launch_reduce_local(reduce_args(e), e->stream, &e->launches));
    e->reduced = true;
    return SB2_OK;
}

int sb2_eval_end(sb2_engine *e, double *perLocusCoal, double *perLocusLik, double *perBranch, double *total3)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!e->reduced) return e->fail(SB2_ERR_STATE, "sb2_eval_end without sb2_eval_begin");
    CUDA_TRY(e, launch_finish(reduce_args(e), e->stream, &e->launches));
    e->memoFlush = false;
    e->pollTag = 0;
    int rc = download(e);
    if (rc) return rc;
    copy_out(e, perLocusCoal, perLocusLik, perBranch, total3);
    e->reduced = false;
    return SB2_OK;
}

int sb2_eval_posterior(sb2_engine *e, double *perLocusCoal, double *perLocusLik, double *perBranch, double *total3)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    int rc;
    if ((rc = eval_full(e))) return rc;
    copy_out(e, perLocusCoal, perLocusLik, perBranch, total3);
    return SB2_OK;
}

int sb2_eval_coalescent(sb2_engine *e, double *perLocus, double *perBranch, double *total)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    int rc;
    // Phase 1 of the reference's posterior.  By default the likelihood phase is evaluated SPECULATIVELY in the same device
    // pass (one host round trip per MCMC step instead of two: -10 us per step at C2); its result is only handed out by
    // sb2_eval_likelihood, which the reference's CompoundDistribution never calls when this phase returns -infinity.
    // SB2_NO_SPECULATION=1 restores the strict protocol: no pruning kernel is launched before the coalescent is known.
    const bool speculate = getenv("SB2_NO_SPECULATION") == nullptr;   // read per call: tests flip it
    if (speculate && (e->needPrepare || e->walkPending || e->resultValid)) {
        if ((rc = eval_full(e))) return rc;
    } else {
        if ((rc = do_prepare(e))) return rc;
        // the pruning kernel has not run for the re-scheduled loci yet, so the likelihood entries of this pass are the
        // cached ones (coalOnly); sb2_eval_likelihood runs the walk
        ReduceArgs r = reduce_args(e, /*coalOnly=*/true);
        e->pollTag = 0;
        if ((rc = launch_final(e, r))) return rc;
        if ((rc = download(e))) return rc;
    }
    copy_out(e, perLocus, nullptr, perBranch, nullptr);
    if (total) *total = e->hOut[0];
    return SB2_OK;
}

int sb2_eval_likelihood(sb2_engine *e, double *perLocus, double *total)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    int rc;
    if ((rc = eval_full(e))) return rc;
    copy_out(e, nullptr, perLocus, nullptr, nullptr);
    if (total) *total = e->hOut[1];
    return SB2_OK;
}

int sb2_reduce_vector(sb2_engine *e, void **devPtr, int32_t *nDoubles)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (devPtr) *devPtr = e->dRed;
    if (nDoubles) *nDoubles = e->nRed;
    return SB2_OK;
}

int sb2_bind_reduce_vector(sb2_engine *e, void *devPtr)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    e->dRed = devPtr ? (double *)devPtr : e->dRedOwn;
    return SB2_OK;
}

int sb2_read_reduce_vector(sb2_engine *e, double *host)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!host) return e->fail(SB2_ERR_ARG, "null host buffer");
    CUDA_TRY(e, cudaMemcpyAsync(host, e->dRed, sizeof(double) * (size_t)e->nRed, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    return SB2_OK;
}

int sb2_write_reduce_vector(sb2_engine *e, const double *host)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!host) return e->fail(SB2_ERR_ARG, "null host buffer");
    CUDA_TRY(e, cudaMemcpyAsync(e->dRed, host, sizeof(double) * (size_t)e->nRed, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));   // `host` may be reused by the caller as soon as this returns
    return SB2_OK;
}

int sb2_comm_handle(sb2_engine *e, void *handle64)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CUDA_TRY(e, cudaIpcGetMemHandle(&h, e->commBuf));
    memcpy(handle64, &h, 64);
    return SB2_OK;
}

int sb2_comm_connect(sb2_engine *e, int32_t rank, int32_t nRanks, const void *handles)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (nRanks < 1 || nRanks > COMM_MAX_RANKS || rank < 0 || rank >= nRanks || !handles) return e->fail(SB2_ERR_ARG, "bad rank %d / %d (max %d ranks)", rank, nRanks, COMM_MAX_RANKS);
    for (int r = 0; r < nRanks; r++) {
        unsigned char *base;
        if (r == rank) base = (unsigned char *)e->commBuf;
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const unsigned char *)handles + 64 * (size_t)r, 64);
            void *p = nullptr;
            CUDA_TRY(e, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            e->commOpened.push_back(p);
            base = (unsigned char *)p;
        }
        e->comm.peerSlots[r] = (unsigned long long *)base;
        e->comm.peerFlags[r] = (unsigned long long *)(base + e->commSlotsBytes);
    }
    e->comm.rank = rank; e->comm.nRanks = nRanks; e->comm.nRed = e->nRed; e->comm.seq = 0;
    e->commConnected = nRanks > 1;
    return SB2_OK;
}

int sb2_set_stream(sb2_engine *e, void *cudaStream)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->ownStream && e->stream) cudaStreamDestroy(e->stream);
    e->stream = (cudaStream_t)cudaStream;
    e->ownStream = false;
    return SB2_OK;
}

int sb2_store(sb2_engine *e)
{
    if (e && e->finalized) { int jrc = join_kstate(e); if (jrc) return jrc; }
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    // Inputs staged since the last evaluation belong to the state that is being stored (the host mirrors below include
    // them): bring the device state up to them first, or a later restore() would drop them without ever uploading them.
    int rc = SB2_OK;
    if (e->needPrepare) {
        if ((rc = do_prepare(e))) return rc;
        if ((rc = do_walk(e))) return rc;
        // the rank-local sums also refresh the cached per-locus values (cur.logL) that a restore() falls back on; no
        // cross-GPU exchange here: ranks need not agree on whether they had anything staged
        { int jrc = join_kstate(e); if (jrc) return jrc; }
        CUDA_TRY(e, launch_reduce_local(reduce_args(e), e->stream, &e->launches));
    }
    if ((rc = flush_pending_walk(e))) return rc;
    e->hasStored = true;
    e->storePending = true;   // the device copy is owed to the next k_prepare (do_prepare), which knows how little differs
    // host staging mirrors of what just became the stored state
    for (int32_t l : e->touched) {
        const LocusDesc &d = e->loci[l].d;
        e->hParamsStored[l] = e->hParams[l];
        memcpy(e->hCatStored + d.catBase, e->hCat + d.catBase, 8 * (size_t)d.nCat);
        memcpy(e->hCatStored + e->totalCat + d.catBase, e->hCat + e->totalCat + d.catBase, 8 * (size_t)d.nCat);
        if (e->hExplicit) memcpy(e->hExplicitStored + d.nodeBase, e->hExplicit + d.nodeBase, 8 * (size_t)d.nNodes);
        e->touchedFlag[l] = 0;
    }
    e->touched.clear();
    memcpy(e->hSpeciesStored, e->hSpecies, e->speciesBytes);
    e->popKindStored = e->popKind;
    return SB2_OK;
}

int sb2_restore(sb2_engine *e)
{
    if (e && e->finalized) { int jrc = join_kstate(e); if (jrc) return jrc; }
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!e->hasStored) return e->fail(SB2_ERR_STATE, "sb2_restore without a prior sb2_store");
    // swap current / stored, like the reference's restore() methods -- unless nothing was evaluated since store(): then the
    // device state IS the stored state (and the stored block's owed copy has not happened yet)
    if (!e->storePending) e->curIdx = 1 - e->curIdx;
    // staged proposal inputs belong to the rejected state: drop them and bring the host mirrors back
    std::fill(e->treeDirty.begin(), e->treeDirty.end(), 0);
    std::fill(e->modelDirty.begin(), e->modelDirty.end(), 0);
    std::fill(e->clockDirty.begin(), e->clockDirty.end(), 0);
    for (int32_t l : e->touched) {
        const LocusDesc &d = e->loci[l].d;
        e->hParams[l] = e->hParamsStored[l];
        memcpy(e->hCat + d.catBase, e->hCatStored + d.catBase, 8 * (size_t)d.nCat);
        memcpy(e->hCat + e->totalCat + d.catBase, e->hCatStored + e->totalCat + d.catBase, 8 * (size_t)d.nCat);
        if (e->hExplicit) { memcpy(e->hExplicit + d.nodeBase, e->hExplicitStored + d.nodeBase, 8 * (size_t)d.nNodes); e->explicitDirty = true; }
        e->touchedFlag[l] = 0;
        e->reuploadFirst = std::min(e->reuploadFirst, (int)l);
        e->reuploadLast = std::max(e->reuploadLast, (int)l);
        e->catDirty = true;   // device copies of the parameters still hold the rejected values: re-upload, no recompute
    }
    e->touched.clear();
    memcpy(e->hSpecies, e->hSpeciesStored, e->speciesBytes);
    e->popKind = e->popKindStored;
    e->speciesDirty = false;
    e->forceAll = false;
    e->needPrepare = e->reuploadLast >= 0;
    e->walkPending = false; e->tilesValid = false; e->reduced = false; e->resultValid = false;
    return SB2_OK;
}

int sb2_accept(sb2_engine *e)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    return flush_pending_walk(e);
}

int sb2_mark_all_dirty(sb2_engine *e)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    e->forceAll = true; e->needPrepare = true; e->memoFlush = true;
    return SB2_OK;
}

int sb2_get_branch_rates(sb2_engine *e, int32_t locus, double *rates)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    int rc = check_locus(e, locus);
    if (rc) return rc;
    const LocusDesc &d = e->loci[locus].d;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    CUDA_TRY(e, cudaMemcpy(rates, cur(e).bRate + d.nodeBase, 8 * (size_t)d.nNodes, cudaMemcpyDeviceToHost));
    return SB2_OK;
}

int sb2_get_embedding(sb2_engine *e, int32_t locus, int32_t *lineageCounts, int32_t *coalCounts, int32_t *assignment, double *perBranchLogP)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    int rc = check_locus(e, locus);
    if (rc) return rc;
    const LocusDesc &d = e->loci[locus].d;
    const size_t o = (size_t)locus * e->S;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    if (lineageCounts) CUDA_TRY(e, cudaMemcpy(lineageCounts, cur(e).brN + o, 4 * (size_t)e->S, cudaMemcpyDeviceToHost));
    if (coalCounts) CUDA_TRY(e, cudaMemcpy(coalCounts, cur(e).brK + o, 4 * (size_t)e->S, cudaMemcpyDeviceToHost));
    if (assignment) CUDA_TRY(e, cudaMemcpy(assignment, cur(e).assign + d.nodeBase, 4 * (size_t)d.nNodes, cudaMemcpyDeviceToHost));
    if (perBranchLogP) CUDA_TRY(e, cudaMemcpy(perBranchLogP, cur(e).brLogP + o, 8 * (size_t)e->S, cudaMemcpyDeviceToHost));
    return SB2_OK;
}

int sb2_get_partials(sb2_engine *e, int32_t locus, int32_t node, double *partials, int32_t *exponents)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    int rc = check_locus(e, locus);
    if (rc) return rc;
    const LocusDesc &d = e->loci[locus].d;
    if (node < d.nTips || node >= d.nNodes) return e->fail(SB2_ERR_ARG, "node %d is not an internal node", node);
    if (d.kStates) return e->fail(SB2_ERR_ARG, "K-state loci keep no resident partials");
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    uint8_t buf;
    CUDA_TRY(e, cudaMemcpy(&buf, cur(e).curP + d.nodeBase + node, 1, cudaMemcpyDeviceToHost));
    const size_t idx = (size_t)buf * d.nInt + (node - d.nTips);
    // device rows are padded to patStride patterns; the caller's buffers are compact [category][pattern]
    const size_t rowElems = (size_t)d.nCat * d.patStride;
    if (partials)
        CUDA_TRY(e, cudaMemcpy2D(partials, 32 * (size_t)d.nPat, e->dPartials + d.partBase + idx * rowElems * 4, 32 * (size_t)d.patStride,
                                 32 * (size_t)d.nPat, (size_t)d.nCat, cudaMemcpyDeviceToHost));
    if (exponents) {
        std::vector<int16_t> tmp(rowElems);
        CUDA_TRY(e, cudaMemcpy(tmp.data(), e->dCumExp + d.expBase + idx * rowElems, 2 * rowElems, cudaMemcpyDeviceToHost));
        for (int c = 0; c < d.nCat; c++)
            for (int p = 0; p < d.nPat; p++) exponents[(size_t)c * d.nPat + p] = tmp[(size_t)c * d.patStride + p];
    }
    return SB2_OK;
}

int sb2_get_matrices(sb2_engine *e, int32_t locus, int32_t node, double *matrices)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    int rc = check_locus(e, locus);
    if (rc) return rc;
    const LocusDesc &d = e->loci[locus].d;
    if (node < 0 || node >= d.nNodes) return e->fail(SB2_ERR_ARG, "node %d out of range", node);
    if (d.kStates) return e->fail(SB2_ERR_ARG, "K-state loci keep no resident matrices");
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    uint8_t buf;
    CUDA_TRY(e, cudaMemcpy(&buf, cur(e).curM + d.nodeBase + node, 1, cudaMemcpyDeviceToHost));
    CUDA_TRY(e, cudaMemcpy(matrices, e->dMats + d.matBase + ((size_t)buf * d.nNodes + node) * d.nCat * 16, 8 * (size_t)d.nCat * 16, cudaMemcpyDeviceToHost));
    return SB2_OK;
}

int sb2_get_pattern_logl(sb2_engine *e, int32_t locus, double *patternLogL)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    int rc = check_locus(e, locus);
    if (rc) return rc;
    const LocusDesc &d = e->loci[locus].d;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    CUDA_TRY(e, cudaMemcpy(patternLogL, e->dPatLogL + d.patBase, 8 * (size_t)d.nPat, cudaMemcpyDeviceToHost));
    return SB2_OK;
}

// ---- operator-side queries over the resident gene trees -----------------------------------------------------------------

static double dec_min(unsigned long long u)
{
    const unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// leafClass: class bits per species leaf.  out3 = {tipward, rootward, maxHeight}; mask may be null.
static int run_query(sb2_engine *e, int mode, const uint8_t *leafClass, uint8_t *mask, double *out3, int64_t *count)
{
    if (!e->everPrepared) return e->fail(SB2_ERR_STATE, "no evaluated state: evaluate the posterior before querying the gene trees");
    for (int l = 0; l < e->L; l++)
        if (e->treeDirty[l]) return e->fail(SB2_ERR_STATE, "locus %d has a staged gene tree that was never evaluated: the query reads the last evaluated (or restored) state", l);
    if (!e->dQuery) {
        e->queryLeafPad = ((size_t)e->nSpeciesLeaves + 15) & ~(size_t)15;
        const size_t bytes = 32 + e->queryLeafPad + (size_t)e->totalNodes;
        int rc;
        if ((rc = dev_alloc(e, &e->dQuery, bytes))) return rc;
        if ((rc = pin_alloc(e, &e->hQuery, bytes))) return rc;
    }
    unsigned long long *hdr = reinterpret_cast<unsigned long long *>(e->hQuery);
    hdr[0] = hdr[1] = hdr[2] = 0xFFF0000000000000ull;   // enc(+inf)
    hdr[3] = 0;
    memcpy(e->hQuery + 32, leafClass, (size_t)e->nSpeciesLeaves);
    cudaStream_t st = e->stream;
    CUDA_TRY(e, cudaMemcpyAsync(e->dQuery, e->hQuery, 32 + e->queryLeafPad, cudaMemcpyHostToDevice, st));
    QueryArgs a;
    const StateBlock &c = cur(e);
    a.loci = e->dLoci; a.tipSpecies = e->dTipSpecies; a.leafClass = e->dQuery + 32;
    a.gParent = c.gParent; a.gLeft = c.gLeft; a.gRight = c.gRight; a.gHeight = c.gHeight;
    a.mask = mask ? e->dQuery + 32 + e->queryLeafPad : nullptr;
    a.result = reinterpret_cast<unsigned long long *>(e->dQuery);
    a.mode = mode; a.maxNodes = e->maxNodes;
    if (8 * (size_t)e->maxNodes > 227 * 1024) return e->fail(SB2_ERR_ARG, "gene trees of %d nodes exceed the query kernel's shared memory", e->maxNodes);
    CUDA_TRY(e, launch_tip_class_query(a, e->L, st));
    e->launches++;
    CUDA_TRY(e, cudaMemcpyAsync(e->hQuery, e->dQuery, mask ? 32 + e->queryLeafPad + (size_t)e->totalNodes : 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(e, cudaStreamSynchronize(st));
    for (int i = 0; i < 3; i++) out3[i] = dec_min(hdr[i]);
    if (count) *count = (int64_t)hdr[3];
    if (mask) memcpy(mask, e->hQuery + 32 + e->queryLeafPad, (size_t)e->totalNodes);
    return SB2_OK;
}

int sb2_connecting_nodes(sb2_engine *e, int32_t speciesNode, uint8_t *mask, double *freedoms, int64_t *count)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!freedoms) return e->fail(SB2_ERR_ARG, "null freedoms");
    SpeciesView v = species_view(e);
    if (speciesNode < 0 || speciesNode >= e->S || v.sLeft[speciesNode] < 0) return e->fail(SB2_ERR_ARG, "species node %d is not an internal node", speciesNode);
    // findDescendants (CoordinatedOperator.java:27-43): leaves under the left child -> 1, under the right child -> 2, others -> 4
    std::vector<uint8_t> cls((size_t)e->nSpeciesLeaves, 4);
    std::vector<int32_t> stack;
    for (int side = 0; side < 2; side++) {
        stack.assign(1, side ? v.sRight[speciesNode] : v.sLeft[speciesNode]);
        while (!stack.empty()) {
            const int32_t n = stack.back();
            stack.pop_back();
            if (v.sLeft[n] < 0) {
                if (n >= e->nSpeciesLeaves) return e->fail(SB2_ERR_ARG, "species leaf %d is not numbered below the leaf count", n);
                cls[n] = (uint8_t)(1 << side);
            } else { stack.push_back(v.sLeft[n]); stack.push_back(v.sRight[n]); }
        }
    }
    double out3[3];
    int rc = run_query(e, QUERY_CONNECTING, cls.data(), mask, out3, count);
    if (rc) return rc;
    freedoms[0] = out3[0]; freedoms[1] = out3[1];
    return SB2_OK;
}

int sb2_max_height(sb2_engine *e, const uint8_t *leafIsRight, double *maxHeight)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!leafIsRight || !maxHeight) return e->fail(SB2_ERR_ARG, "null argument");
    std::vector<uint8_t> cls((size_t)e->nSpeciesLeaves);
    for (int i = 0; i < e->nSpeciesLeaves; i++) cls[i] = leafIsRight[i] ? 2 : 1;
    double out3[3];
    int rc = run_query(e, QUERY_MAX_HEIGHT, cls.data(), nullptr, out3, nullptr);
    if (rc) return rc;
    *maxHeight = out3[2];
    return SB2_OK;
}

int32_t sb2_num_loci(const sb2_engine *e) { return e ? (int32_t)e->loci.size() : 0; }
int64_t sb2_launch_count(const sb2_engine *e) { return e ? e->launches : 0; }

int64_t sb2_last_node_updates(sb2_engine *e)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return -1;
    std::vector<int32_t> n(e->L);
    cudaStreamSynchronize(e->stream);
    if (cudaMemcpy(n.data(), e->dNOps, 4 * (size_t)e->L, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (e->lastFusedSingle >= 0) return n[e->lastFusedSingle];   // a single-locus k_step refreshes only that locus's entry
    int64_t t = 0;
    for (int v : n) t += v;
    return t;
}

int sb2_kernel_time(sb2_engine *e, double *ms, int64_t *launches)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    double t = 0.0;
    for (size_t i = 0; i < e->evUsed; i++) {
        float f = 0.f;
        CUDA_TRY(e, cudaEventElapsedTime(&f, e->evPool[i].first, e->evPool[i].second));
        t += f;
    }
    if (ms) *ms = t;
    if (launches) *launches = (int64_t)e->evUsed;
    e->evUsed = 0;
    return SB2_OK;
}

int sb2_set_profile(sb2_engine *e, int32_t on)
{
    if (e) cudaSetDevice(e->cfg.device);   // several engines (one per GPU) may be driven by one host thread
    if (!e) return SB2_ERR_ARG;
    e->cfg.profile = on ? 1 : 0;
    return SB2_OK;
}

int64_t sb2_device_bytes(const sb2_engine *e) { return e ? e->devBytes : 0; }

int sb2_step_phase_times(sb2_engine *e, double *us, int32_t n)
{
    if (e) cudaSetDevice(e->cfg.device);
    if (!e || !e->finalized) return e ? e->fail(SB2_ERR_STATE, "engine not finalized") : SB2_ERR_ARG;
    if (!us || n < 1) return e->fail(SB2_ERR_ARG, "null output");
    if (!e->hStamps) return e->fail(SB2_ERR_STATE, "no fused step was launched in profile mode");
    CUDA_TRY(e, cudaStreamSynchronize(e->stream));
    for (int i = 0; i < n; i++) us[i] = (i < STEP_STAMPS && e->hStamps[i] >= e->hStamps[0]) ? 1e-3 * (double)(e->hStamps[i] - e->hStamps[0]) : 0.0;
    return SB2_OK;
}

double sb2_exchange_wait_us(const sb2_engine *e) { return (e && e->finalized && e->commConnected) ? e->hOut[e->nOut] : 0.0; }

}  // extern "C"
