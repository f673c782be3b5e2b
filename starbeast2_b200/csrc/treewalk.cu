// k_treewalk: the Felsenstein-pruning kernel (the roofline kernel of the path).
//
// Replaces BeerLikelihoodCore4.calculateStatesStatesPruning / calculateStatesPartialsPruning /
// calculatePartialsPartialsPruning, scalePartials, integratePartials, calculateLogLikelihoods and
// TreeLikelihood.calcLogP (BEAST.base >= 2.7.3, named by /root/reference/version.xml:2 and
// tutorial/CanisPhylogeny/Canis-Phylogeny.xml:684; restated from SURVEY.md Appendix B).
//
// B200-first decomposition: the pruning recursion is independent across site patterns, so the grid is
// (locus x tile of patterns) and ONE CTA walks the locus's whole dirty subtree for its tile, in the post-order
// schedule k_prepare emitted.  Results that a later op of the same walk consumes stay on the SM: they sit in a
// small shared-memory stack (heavy-child-first order keeps it shallow; the rare deeper result is re-read from the
// node's global buffer, which the same thread has just written).  Every result is streamed to HBM exactly once
// (256-bit stores) because later incremental (dirty-node-only) evaluations read clean siblings from there
// (256-bit non-allocating loads).  HBM traffic of a full evaluation is therefore partials written once + tip
// states + exponents, instead of write + two re-reads per node.
//
// Forwarding: every result is consumed exactly once, by its parent; in post-order the parent follows its
// last-evaluated child immediately, so that operand is taken from registers (OPK_PREV) and only the first-evaluated
// child of a node with two dirty children is parked in the shared-memory stack.
//
// Thread mapping: a warp owns 32*R consecutive patterns of ONE rate category, a thread R of them (stride 32, so
// every global / shared access is coalesced / conflict free).  Warps never wait for each other inside the walk:
// each warp streams its own category's transition matrices through a private cp.async ring, the op schedule and
// the tile's tip states are staged in shared memory once, and rescaling is per (pattern, category).
//
// Rescaling: every node, every (pattern, category), by an exact power of two (the exponent of the largest of
// the 4 values); scaled values keep their mantissas, so rescaling never perturbs a result and evaluation is
// independent of the update history.  The cumulative exponent of a subtree is stored as int16 next to the
// partials; categories are brought to a common exponent only at the root.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "walk_common.cuh"

namespace sb2 {

constexpr int OPS_BATCH = 32;   // ops staged in shared memory per half batch
constexpr int MAT_RING = 4;     // transition-matrix pairs in flight per warp (cp.async ring)
constexpr int TIP_SMEM_MAX = 48 * 1024;

bool treewalk_tips_in_smem(int maxTips, int tilePat) { return (size_t)maxTips * tilePat <= TIP_SMEM_MAX; }

size_t treewalk_smem_bytes(int nCat, int chunksPerTile, int R, int stackDepth, int maxTips, int workers)
{
    const size_t warps = (size_t)nCat * chunksPerTile * workers;
    const size_t tilePat = (size_t)chunksPerTile * 32 * R;
    const size_t slots = (size_t)stackDepth + (workers > 1 ? WALK_MAILBOX : 0);
    size_t b = slots * warps * R * 32 * 32;                // partials stack (two double2 planes)
    b += (size_t)warps * MAT_RING * 2 * 16 * 8;            // per-warp matrix rings
    b += 2ull * OPS_BATCH * sizeof(Op);                    // op schedule, two half batches
    b += slots * warps * R * 32 * 2;                       // exponent stack (int16)
    b += 16 * 8;                                           // the ones column
    if (workers > 1) b += ((slots * warps * 4 + 15) & ~size_t(15));   // mailbox flags
    if (treewalk_tips_in_smem(maxTips, (int)tilePat)) b += (size_t)maxTips * tilePat + 16;   // tip states of the tile
    return (b + 15) & ~size_t(15);
}

__device__ __forceinline__ unsigned ld_acquire_shared(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_shared(unsigned *p, unsigned v)
{
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// NC > 0: the launch geometry (categories, chunks per tile, stack depth, tips staged in shared memory) is a compile-time
// constant, so every shared-memory offset and stride folds into an immediate; NC == 0 is the generic run-time path.
// NW > 1 (tree-level parallelism for latency-bound shapes, see sb2_device.cuh): NW worker warps share one pattern chunk;
// worker 0 walks the main path, the helpers the side subtrees k_prepare gave them, handing each finished subtree over
// through a dedicated stack slot (mailbox) guarded by a release / acquire flag in shared memory.
template <bool EXACT, int R, int MINB, int NC, int CH, int DEPTH, int NW>
__global__ void __launch_bounds__(NC ? 32 * NC * CH * NW : 256, MINB) k_treewalk(WalkArgs A)
{
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int tile = blockIdx.x;
    auto stamp = [&](int i) {   // profiling only
        if (A.dbgTimes && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); A.dbgTimes[(size_t)tile * 8 + i] = t; }
    };
    stamp(0);
    if (A.dbgTimes && threadIdx.x == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); A.dbgTimes[(size_t)tile * 8 + 6] = sm; }
    const int l = A.tileLocus[tile];
    const LocusDesc d = A.loci[l];
    const int C = NC ? NC : A.nCat, chunks = NC ? CH : A.chunksPerTile, depth = NC ? DEPTH : A.stackDepth;
    const int inner = C * chunks;                 // warps of one worker
    const int warps = inner * NW;
    const int slots = depth + (NW > 1 ? WALK_MAILBOX : 0);
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int worker = NW > 1 ? w / inner : 0, wi = NW > 1 ? w - worker * inner : w;
    const int c = wi % C, chunk = wi / C;
    const int nPat = d.nPat;
    const int tilePat = chunks * 32 * R;
    const int p0 = (tile - d.tileBase) * tilePat;
    const int pbase = p0 + chunk * 32 * R + lane;   // pattern r of this thread: pbase + 32 r
    const bool tipsInSmem = NC ? true : (A.tipsInSmem != 0);

    // shared memory carve-up; everything is addressed as an index off one of these bases so that the compiler keeps
    // the accesses in the shared window (LDS/STS), never generic
    double2 *st2 = reinterpret_cast<double2 *>(smem_raw);                       // stack: [2 planes][depth][warps][R][32] double2
    const int planeStride = slots * warps * R * 32;                             //        plane 0 = states 0,1; plane 1 = states 2,3
    double *smd = reinterpret_cast<double *>(st2 + 2 * planeStride);            // ring [warps][MAT_RING][32] | ones[16]
    const int onesIdx = warps * MAT_RING * 32;
    Op *sops = reinterpret_cast<Op *>(smd + onesIdx + 16);                      // [2 * OPS_BATCH]
    int16_t *sexp = reinterpret_cast<int16_t *>(sops + 2 * OPS_BATCH);          // [slots][warps][R][32]
    unsigned *sflag = reinterpret_cast<unsigned *>(sexp + planeStride);         // [slots][warps] mailbox flags (NW > 1 only)
    uint8_t *sstate = reinterpret_cast<uint8_t *>(sflag + (NW > 1 ? ((slots * warps + 3) & ~3) : 0));   // [nTips][tilePat]; 16-byte aligned by construction
    const uint8_t *states = A.tipStates + d.stateBase;

    // ---- static inputs first: under programmatic dependent launch this part overlaps k_prepare's tail ----
    if (tid < 16) smd[onesIdx + tid] = 1.0;
    if (NW > 1) for (int i = tid; i < slots * warps; i += blockDim.x) sflag[i] = 0u;
    if (tipsInSmem) {
        // the tile's tip states, every tip, as one burst of asynchronous 16-byte copies (rows are padded to 128 bytes with
        // 'missing', so a tile never leaves its row); they ride in the first cp.async group, which the first op waits for
        if (A.tipsInSmem == 2) {   // A/B knob: plain byte loads
            const int total = d.nTips * tilePat;
            for (int idx = tid; idx < total; idx += blockDim.x) {
                const int t = idx / tilePat, pp = idx - t * tilePat;
                sstate[idx] = states[(size_t)t * d.stateStride + p0 + pp];
            }
        } else {
            const int per = tilePat >> 4, total = d.nTips * per;
            for (int idx = tid; idx < total; idx += blockDim.x) {
                const int t = idx / per, q = idx - t * per;
                cp_async16(sstate + t * tilePat + q * 16, states + (size_t)t * d.stateStride + p0 + q * 16);
            }
        }
    }
    // nOps / ops / opMats are k_prepare's outputs: they must not be touched before that grid has completed
    stamp(1);
    cudaGridDependencySynchronize();
    stamp(2);
    cudaTriggerProgrammaticLaunchCompletion();   // the (single-CTA) reduction kernel may queue up behind us already
    const Op *ops = A.ops + d.opBase;
    const double *matStream = A.opMats + ((size_t)d.opBase * C + c) * 32 + lane;   // + k * C*32: this lane's 8 bytes of op k's pair
    const int ringIdx = w * MAT_RING * 32;
    const int matStride = C * 32;
    // Single-worker walks start at op 0 whatever nOps turns out to be: fetch the first batch of the schedule and the first
    // matrix pairs in the same wave of loads as nOps itself (bounded by the locus's allocation, so a clean locus merely
    // reads stale entries) instead of one dependent round trip later.
    int4 specOp = make_int4(0, 0, 0, 0);
    if (NW == 1) {
        if (tid < OPS_BATCH && tid < d.nInt) specOp = reinterpret_cast<const int4 *>(ops)[tid];
        for (int j = 0; j < MAT_RING - 1; j++) {
            if (j < d.nInt) cp_async8(smd + ringIdx + j * 32 + lane, matStream + (size_t)j * matStride);
            cp_async_commit();
        }
    }
    const int nOps = A.nOps[l];
    // this worker's share of the schedule: ops [kBeg, kEnd) of the locus's list
    const int kBeg = NW > 1 ? A.laneOps[l * (WALK_MAX_WORKERS + 1) + worker] : 0;
    int kEnd = NW > 1 ? A.laneOps[l * (WALK_MAX_WORKERS + 1) + worker + 1] : nOps;
    if (A.debugMaxOps >= 0 && NW == 1) kEnd = min(kEnd, kBeg + A.debugMaxOps);   // timing experiments only: results are garbage
    if (nOps == 0) {         // locus clean: cached logL stands
        cp_async_commit();
        cp_async_wait<0>();
        return;
    }
    const bool incremental = nOps < d.nInt;         // only then can operands be clean nodes resident in HBM

    double *partials = A.partials + d.partBase + ((size_t)c * d.patStride + pbase) * 4;   // + idx*4 (+ r*128)
    int16_t *cumExp = A.cumExp + d.expBase + (size_t)c * d.patStride + pbase;              // + idx   (+ r*32)
    // keep the three per-thread base pointers in registers: under the 64-register cap ptxas otherwise re-derives them
    // (64-bit multiplies) inside every iteration of the walk
    asm volatile("" : "+l"(partials));
    asm volatile("" : "+l"(cumExp));
    asm volatile("" : "+l"(matStream));
    const int stateIdx = chunk * 32 * R + lane;
    const int stackIdx = w * R * 32 + lane;                 // + slot*slotStride + r*32
    const int slotStride = warps * R * 32;
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; r++) valid[r] = pbase + 32 * r < nPat;

    auto load_batch = [&](int b) {
        const int first = b * OPS_BATCH, cnt = min(OPS_BATCH, nOps - first);
        const int4 *src = reinterpret_cast<const int4 *>(ops + first);
        int4 *dst = reinterpret_cast<int4 *>(sops + (b & 1) * OPS_BATCH);
        for (int t = tid; t < cnt; t += blockDim.x) dst[t] = src[t];
    };
    // this warp's matrix pair of op k (256 contiguous bytes of the op-ordered stream) -> ring slot k % MAT_RING
    auto issue_mats = [&](int k) {
        if (k < kEnd) cp_async8(smd + ringIdx + (k & (MAT_RING - 1)) * 32 + lane, matStream + (size_t)k * matStride);
        cp_async_commit();
    };

    double prev[R][4];   // result of the previous op (scaled) and its cumulative exponent
    int prevCum[R];
#pragma unroll
    for (int r = 0; r < R; r++) { prev[r][0] = prev[r][1] = prev[r][2] = prev[r][3] = 0.0; prevCum[r] = 0; }

    if (c == 0 && worker == 0) {   // what the root epilogue will read, pulled into L1 while the walk runs
        const LocusParams *prm = &A.params[l];
        prefetch_l1(prm->subst);        // HKY frequencies
        prefetch_l1(prm->subst + 32);   // eigen-system frequencies, pInv, substKind
        prefetch_l1(A.catProps + d.catBase);
#pragma unroll
        for (int r = 0; r < R; r++) if (valid[r]) prefetch_l1(A.weights + d.patBase + pbase + 32 * r);
    }
    if (NW == 1) {
        if (tid < OPS_BATCH) reinterpret_cast<int4 *>(sops)[tid] = specOp;   // batch 0, fetched above
    } else {
        load_batch(0);
        if (nOps > OPS_BATCH) load_batch(1);   // multi-worker walks are limited to 2 * OPS_BATCH ops: all resident, no reloads
        for (int j = 0; j < MAT_RING - 1; j++) issue_mats(kBeg + j);
    }
    cp_async_wait<MAT_RING - 2>();   // group 0 (tip states + op 0's matrices) has landed for this thread ...
    __syncthreads();                 // ... and for everybody else; the op schedule is visible too

    stamp(3);
    for (int k = kBeg; k < kEnd; k++) {
        if (NW == 1 && (k & (OPS_BATCH - 1)) == OPS_BATCH / 2 && (k / OPS_BATCH + 1) * OPS_BATCH < nOps) {
            __syncthreads();                       // every warp has left the previous batch's half
            load_batch(k / OPS_BATCH + 1);
            __syncthreads();
        }
        cp_async_wait<MAT_RING - 2>();             // this lane's pieces of op k's matrices have landed
        __syncwarp();                              // ... and the other lanes'; the warp is also done with op k-1
        issue_mats(k + MAT_RING - 1);              // reuses the ring slot of op k-1
        if (incremental && k + 2 < kEnd) {         // clean children live in HBM: pull their sectors towards L2 two ops ahead
            const Op q = sops[(k + 2) & (2 * OPS_BATCH - 1)];
#pragma unroll
            for (int r = 0; r < R; r++) {
                if ((q.ctl & 7) == OPK_GLOBAL && valid[r]) prefetch_l2(partials + (size_t)q.aOff * 4 + r * 128);
                if (((q.ctl >> 3) & 7) == OPK_GLOBAL && valid[r]) prefetch_l2(partials + (size_t)q.bOff * 4 + r * 128);
            }
        }

        const Op o = sops[k & (2 * OPS_BATCH - 1)];
        const int mIdx = ringIdx + (k & (MAT_RING - 1)) * 32;
        double out[R][4];
        int cum[R];
#pragma unroll
        for (int r = 0; r < R; r++) cum[r] = 0;

#pragma unroll
        for (int side = 0; side < 2; side++) {
            const unsigned kind = (o.ctl >> (3 * side)) & 7;
            const unsigned off = side ? o.bOff : o.aOff;
            const int m = mIdx + side * 16;
            if (kind == OPK_TIP) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int st = tipsInSmem ? sstate[off * tilePat + stateIdx + 32 * r]
                                              : states[(size_t)off * d.stateStride + pbase + 32 * r];
                    const int col = (st < 4) ? (m + st) : onesIdx;   // missing / ambiguous: factor 1
                    if (side == 0) { out[r][0] = smd[col]; out[r][1] = smd[col + 4]; out[r][2] = smd[col + 8]; out[r][3] = smd[col + 12]; }
                    else { out[r][0] *= smd[col]; out[r][1] *= smd[col + 4]; out[r][2] *= smd[col + 8]; out[r][3] *= smd[col + 12]; }
                }
            } else {
                if (kind == OPK_PREV) {   // the previous op's result, forwarded in registers
#pragma unroll
                    for (int r = 0; r < R; r++) cum[r] += prevCum[r];
                    if (side == 0) rows_times<EXACT, R, true>(smd + m, prev, out);
                    else rows_times<EXACT, R, false>(smd + m, prev, out);
                    continue;
                }
                double a[R][4];
                if (kind == OPK_STACK || (NW > 1 && kind == OPK_REMOTE)) {
                    int si0 = ((o.ctl >> (6 + 4 * side)) & 15) * slotStride + stackIdx;
                    if (NW > 1 && kind == OPK_REMOTE) {   // a helper's finished subtree: wait for its mailbox flag, read its slot
                        const int src = (int)off * inner + wi;
                        const unsigned *flag = sflag + ((o.ctl >> (6 + 4 * side)) & 15) * warps + src;
                        while (ld_acquire_shared(flag) == 0u) { }
                        si0 += (src - w) * R * 32;
                    }
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const double2 lo = st2[si0 + 32 * r], hi = st2[planeStride + si0 + 32 * r];
                        a[r][0] = lo.x; a[r][1] = lo.y; a[r][2] = hi.x; a[r][3] = hi.y;
                        cum[r] += sexp[si0 + 32 * r];
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.0;
                        if (valid[r]) {
                            const double *src = partials + (size_t)off * 4 + r * 128;
                            if (kind == OPK_GLOBAL) ld256_nc(src, a[r][0], a[r][1], a[r][2], a[r][3]);
                            else ld256_coherent(src, a[r][0], a[r][1], a[r][2], a[r][3]);   // OPK_SPILL: written by this thread in this launch
                            cum[r] += cumExp[(size_t)off + 32 * r];
                        }
                    }
                }
                if (side == 0) rows_times<EXACT, R, true>(smd + m, a, out);
                else rows_times<EXACT, R, false>(smd + m, a, out);
            }
        }

        // ---- exact power-of-two rescale, store ----
        double *dst = partials + (size_t)o.outOff * 4;
        int16_t *dstE = cumExp + (size_t)o.outOff;
        const int outSlot = (o.ctl >> 14) & 255;
        const bool toStack = outSlot < slots;
        const int so0 = outSlot * slotStride + stackIdx;
#pragma unroll
        for (int r = 0; r < R; r++) {
            double o0, o1, o2, o3;
            const int ce = cum[r] + rescale4(out[r], o0, o1, o2, o3);
            if (toStack) {
                st2[so0 + 32 * r] = make_double2(o0, o1);
                st2[planeStride + so0 + 32 * r] = make_double2(o2, o3);
                sexp[so0 + 32 * r] = (int16_t)ce;
            }
            if (valid[r]) {
                st256(dst + r * 128, o0, o1, o2, o3);
                dstE[32 * r] = (int16_t)ce;
            }
            prev[r][0] = o0; prev[r][1] = o1; prev[r][2] = o2; prev[r][3] = o3;
            prevCum[r] = ce;
        }
        if (NW > 1 && (o.ctl & OP_SIGNAL)) {   // hand the finished subtree to worker 0
            __syncwarp();
            if (lane == 0) st_release_shared(sflag + outSlot * warps + w, 1u);
        }
    }
    cp_async_wait<0>();
    stamp(4);
    __syncthreads();

    // root: integratePartials over categories, (+pInv on constant patterns), frequencies, log, pattern weight; then the sum
    // over each 32-pattern chunk (fixed-shape shuffle tree).  The locus's log-likelihood is the sum of its chunk sums in
    // pattern order (k_reduce_*), whatever the launch geometry: the fused one-launch step (step.cu) adds the same numbers in
    // the same order.
    if (c == 0 && worker == 0) {
        const LocusParams *prm = &A.params[l];
        const double *freqs = (prm->substKind == 0) ? prm->subst + 1 : prm->subst + 36;
        const double *props = A.catProps + d.catBase;
        const double pInv = prm->pInv;
#pragma unroll
        for (int r = 0; r < R; r++) {
            double contrib = 0.0;
            if (valid[r]) {
                const int p = pbase + 32 * r;
                const int mask = (pInv > 0.0) ? A.constMask[d.patBase + p] : 0;
                const double lk = root_pattern_logl<EXACT>(st2, planeStride, sexp, chunk * C * R * 32 + 32 * r + lane, R * 32, C, props, freqs, pInv, mask);
                A.patLogL[d.patBase + p] = lk;
                contrib = __dmul_rn(lk, (double)A.weights[d.patBase + p]);
            }
            contrib = chunk_sum(contrib);
            const int ch = (p0 >> 5) + chunk * R + r;   // 32-pattern chunk of the locus
            if (lane == 0 && ch < d.nChunks) A.chunkSum[d.chunkBase + ch] = contrib;
        }
    }
    stamp(5);
}

template <bool EXACT, int R, int MINB, int NC, int CH, int DEPTH, int NW = 1>
static cudaError_t launch_one(const WalkArgs &a, int nTiles, size_t smem, int threads, cudaStream_t st)
{
    // resident CTAs per SM can be capped by padding the dynamic shared memory (WalkArgs::ctasPerSM, 0 = whatever fits)
    if (a.ctasPerSM > 0) smem = std::max(smem, (size_t)(227 * 1024 / a.ctasPerSM - 1024) & ~size_t(15));
    static size_t configured[MAX_DEVICES + 1] = {};
    cudaError_t rc = ensure_dynamic_smem(k_treewalk<EXACT, R, MINB, NC, CH, DEPTH, NW>, smem, configured);
    if (rc != cudaSuccess) return rc;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nTiles); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_treewalk<EXACT, R, MINB, NC, CH, DEPTH, NW>, a);
}

template <bool EXACT>
static cudaError_t launch_exact(const WalkArgs &a, int nTiles, size_t smem, int threads, cudaStream_t st)
{
    static const bool forceGeneric = getenv("SB2_GENERIC") != nullptr;   // A/B knob
    const bool std4 = a.stackDepth == 4 && a.tipsInSmem == 1 && !forceGeneric;
    if (a.workers > 1) {   // tree-level parallelism: engine.cu only selects it for these shapes
        if (a.nCat == 1 && a.workers == 4) return launch_one<EXACT, 1, 4, 1, 1, 3, 4>(a, nTiles, smem, threads, st);
        else if (a.nCat == 1 && a.workers == 2) return launch_one<EXACT, 1, 4, 1, 1, 3, 2>(a, nTiles, smem, threads, st);
        else if (a.nCat == 4 && a.workers == 2) return launch_one<EXACT, 1, 4, 4, 1, 3, 2>(a, nTiles, smem, threads, st);
        return cudaErrorInvalidValue;   // engine.cu only selects the shapes above; never abort() inside a JVM
    }
    if (a.patPerThread == 1) {
        if (std4 && a.nCat == 4 && a.chunksPerTile == 1) return launch_one<EXACT, 1, 4, 4, 1, 4>(a, nTiles, smem, threads, st);        // GTR+G4 shapes
        else if (std4 && a.nCat == 1 && a.chunksPerTile == 2) return launch_one<EXACT, 1, 4, 1, 2, 4>(a, nTiles, smem, threads, st);   // no rate heterogeneity
        else return launch_one<EXACT, 1, 4, 0, 0, 0>(a, nTiles, smem, threads, st);
    } else if (a.patPerThread == 2) {
        if (a.stackDepth == 3 && a.tipsInSmem == 1 && a.nCat == 4 && a.chunksPerTile == 1) return launch_one<EXACT, 2, 5, 4, 1, 3>(a, nTiles, smem, threads, st);   // 5 CTAs / SM
        else if (std4 && a.nCat == 4 && a.chunksPerTile == 1) return launch_one<EXACT, 2, 2, 4, 1, 4>(a, nTiles, smem, threads, st);
        else if (std4 && a.nCat == 1 && a.chunksPerTile == 2) return launch_one<EXACT, 2, 2, 1, 2, 4>(a, nTiles, smem, threads, st);
        else return launch_one<EXACT, 2, 2, 0, 0, 0>(a, nTiles, smem, threads, st);
    } else {
        return launch_one<EXACT, 4, 1, 0, 0, 0>(a, nTiles, smem, threads, st);
    }
}

cudaError_t launch_treewalk(const WalkArgs &a, int nTiles, bool exact, cudaStream_t st)
{
    const size_t smem = treewalk_smem_bytes(a.nCat, a.chunksPerTile, a.patPerThread, a.stackDepth, a.maxTips, a.workers);
    const int threads = 32 * a.nCat * a.chunksPerTile * a.workers;
    return exact ? launch_exact<true>(a, nTiles, smem, threads, st) : launch_exact<false>(a, nTiles, smem, threads, st);
}

}  // namespace sb2
