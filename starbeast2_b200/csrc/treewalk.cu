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
// each warp streams its own category's schedule -- op descriptors AND transition-matrix pairs, one OpRec per op, laid out
// [locus][category][op] by k_prepare -- through a private double-buffered ring of OPREC_CHUNK-op chunks, each chunk ONE
// bulk asynchronous copy (cp.async.bulk, issued by an elected lane, completing on the buffer's mbarrier); the tile's
// tip states arrive the same way, one bulk copy per tip row; rescaling is per (pattern, category).
//
// Parked results: a first-evaluated child waits for its sibling's subtree either in the thread's REGISTERS (OPK_REG: whenever
// that subtree parks nothing itself -- most parks of a random tree; the two register sets then simply swap roles, nothing is
// copied) or in the shared-memory stack.
//
// Rescaling: every node, every (pattern, category), by an exact power of two (the exponent of the largest of
// the 4 values); scaled values keep their mantissas, so rescaling never perturbs a result and evaluation is
// independent of the update history.  The cumulative exponent of a subtree is stored as int16 next to the
// partials; categories are brought to a common exponent only at the root.
#include <math.h>
#include <stdlib.h>

#include "walk_common.cuh"

namespace sb2 {

constexpr int CHUNK = OPREC_CHUNK;          // ops per bulk copy
constexpr int MAT_STRIDE = OPMAT_STRIDE;    // doubles per matrix of an OpRec (sb2_device.cuh)
constexpr int TIP_SMEM_MAX = 48 * 1024;

bool treewalk_tips_in_smem(int maxTips, int tilePat) { return (size_t)maxTips * tilePat <= TIP_SMEM_MAX; }

size_t treewalk_smem_bytes(int nCat, int chunksPerTile, int R, int stackDepth, int maxTips, int workers)
{
    const size_t warps = (size_t)nCat * chunksPerTile * workers;
    const size_t tilePat = (size_t)chunksPerTile * 32 * R;
    const size_t slots = (size_t)stackDepth + (workers > 1 ? WALK_MAILBOX : 0);
    size_t b = slots * warps * R * 32 * 32;                // partials stack (two double2 planes)
    b += warps * 2 * CHUNK * sizeof(OpRec);                // per-warp schedule rings: two chunk buffers
    b += warps * 2 * 8 + 32;                               // their mbarriers + the tip-state tile's
    b += slots * warps * R * 32 * 2;                       // exponent stack (int16)
    if (workers > 1) b += ((slots * warps * 4 + 15) & ~size_t(15));   // mailbox flags
    if (treewalk_tips_in_smem(maxTips, (int)tilePat)) b += (size_t)maxTips * tilePat + 16;   // tip states of the tile
    return (b + 31) & ~size_t(15);
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
    OpRec *ring = reinterpret_cast<OpRec *>(st2 + 2 * planeStride);             // [warps][2][CHUNK]
    unsigned long long *sbar = reinterpret_cast<unsigned long long *>(ring + warps * 2 * CHUNK);   // [warps][2] chunk-buffer mbarriers | [4]: tip tile
    unsigned long long *tipBar = sbar + warps * 2;
    int16_t *sexp = reinterpret_cast<int16_t *>(tipBar + 4);                    // [slots][warps][R][32]
    unsigned *sflag = reinterpret_cast<unsigned *>(sexp + planeStride);         // [slots][warps] mailbox flags (NW > 1 only)
    const int stateOff = (int)(((reinterpret_cast<unsigned char *>(sflag + (NW > 1 ? slots * warps : 0)) - smem_raw) + 15) & ~15);
    uint8_t *sstate = smem_raw + stateOff;                                      // [nTips][tilePat], 16-byte aligned
    const uint8_t *states = A.tipStates + d.stateBase;
    OpRec *myRing = ring + w * 2 * CHUNK;
    unsigned long long *myBar = sbar + w * 2;

    // ---- static inputs first: under programmatic dependent launch this part overlaps k_prepare's tail ----
    if (NW > 1) for (int i = tid; i < slots * warps; i += blockDim.x) sflag[i] = 0u;
    // mbarriers: one per chunk buffer of every warp (one arrival: the elected lane's expect_tx), one for the tip-state tile
    // (every thread arrives with the bytes of the rows it copies)
    if (lane == 0) {
        mbar_init(myBar, 1); mbar_init(myBar + 1, 1);
        if (tid == 0) mbar_init(tipBar, blockDim.x);
    }
    mbar_init_fence();
    __syncthreads();
    if (tipsInSmem) {
        // the tile's tip states: one bulk copy per tip row (rows are padded to 128 bytes with 'missing', so a tile never
        // leaves its row), all completing on one mbarrier that every thread waits for before its first op
        if (A.tipsInSmem == 2) {   // A/B knob: plain byte loads
            const int total = d.nTips * tilePat;
            for (int idx = tid; idx < total; idx += blockDim.x) {
                const int t = idx / tilePat, pp = idx - t * tilePat;
                sstate[idx] = states[(size_t)t * d.stateStride + p0 + pp];
            }
            mbar_arrive_expect_tx(tipBar, 0);
        } else {
            const int mine = tid < d.nTips ? (d.nTips - tid + (int)blockDim.x - 1) / (int)blockDim.x : 0;
            mbar_arrive_expect_tx(tipBar, (unsigned)(mine * tilePat));
            for (int t = tid; t < d.nTips; t += blockDim.x)
                bulk_g2s(sstate + t * tilePat, states + (size_t)t * d.stateStride + p0, (unsigned)tilePat, tipBar);
        }
    }
    // the schedule is k_prepare's output: it must not be touched before that grid has completed
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();   // the (single-CTA) reduction kernel may queue up behind us already
    const OpRec *stream = A.opStream + (size_t)d.opBase * C + (size_t)c * d.nInt;   // this category's records of the locus
    // chunk j of this warp's list (ops kBeg + j*CHUNK ...) -> chunk buffer j & 1, one bulk copy by the elected lane completing
    // on the buffer's mbarrier (phase parity (j >> 1) & 1).  A chunk is copied whole, up to the locus's allocation (nInt
    // records): entries past the op count are stale but harmless, so the first chunks can be fetched before nOps is known.
    auto issue_chunk = [&](int kFirst, int j) {
        const int first = kFirst + j * CHUNK;
        if (lane == 0 && first < d.nInt) {
            const unsigned bytes = (unsigned)(min(CHUNK, d.nInt - first) * (int)sizeof(OpRec));
            mbar_arrive_expect_tx(myBar + (j & 1), bytes);
            bulk_g2s(myRing + (j & 1) * CHUNK, stream + first, bytes, myBar + (j & 1));
        }
    };
    // Single-worker walks start at op 0 whatever nOps turns out to be: fetch the first two chunks in the same wave of loads as
    // nOps itself instead of one dependent round trip later.
    if (NW == 1) { issue_chunk(0, 0); issue_chunk(0, 1); }
    const int nOps = A.nOps[l];
    // this worker's share of the schedule: ops [kBeg, kEnd) of the locus's list
    const int kBeg = NW > 1 ? A.laneOps[l * (WALK_MAX_WORKERS + 1) + worker] : 0;
    int kEnd = NW > 1 ? A.laneOps[l * (WALK_MAX_WORKERS + 1) + worker + 1] : nOps;
    if (A.debugMaxOps >= 0 && NW == 1) kEnd = min(kEnd, kBeg + A.debugMaxOps);   // timing experiments only: results are garbage
    if (NW > 1) { issue_chunk(kBeg, 0); issue_chunk(kBeg, 1); }
    const int nChunks = (kEnd - kBeg + CHUNK - 1) / CHUNK;
    const int issued = min(2, (d.nInt - kBeg + CHUNK - 1) / CHUNK);   // chunk copies in flight so far
    if (nOps == 0) {         // locus clean: cached logL stands; the copies in flight must land before the CTA's shared memory goes away
        if (tipsInSmem) mbar_wait(tipBar, 0);
        for (int j = 0; j < issued; j++) mbar_wait(myBar + j, 0);
        return;
    }
    const bool incremental = nOps < d.nInt;         // only then can operands be clean nodes resident in HBM

    double *partials = A.partials + d.partBase + ((size_t)c * nPat + pbase) * 4;   // + idx*4 (+ r*128)
    int16_t *cumExp = A.cumExp + d.expBase + (size_t)c * nPat + pbase;              // + idx   (+ r*32)
    // keep the per-thread base pointers in registers: ptxas otherwise re-derives them (64-bit multiplies) inside every
    // iteration of the walk
    asm volatile("" : "+l"(partials));
    asm volatile("" : "+l"(cumExp));
    const int stateIdx = chunk * 32 * R + lane;
    const int stackIdx = w * R * 32 + lane;                 // + slot*slotStride + r*32
    const int slotStride = warps * R * 32;
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; r++) valid[r] = pbase + 32 * r < nPat;

    // two register sets: one receives every op's result (and forwards it to the next op), the other is the register parking
    // slot; they swap roles whenever a result is parked in registers (flip)
    double xv[R][4], yv[R][4];
    int xc[R], yc[R];
#pragma unroll
    for (int r = 0; r < R; r++) { xv[r][0] = xv[r][1] = xv[r][2] = xv[r][3] = 0.0; yv[r][0] = yv[r][1] = yv[r][2] = yv[r][3] = 0.0; xc[r] = yc[r] = 0; }

    if (c == 0 && worker == 0) {   // what the root epilogue will read, pulled into L1 while the walk runs
        const LocusParams *prm = &A.params[l];
        prefetch_l1(prm->subst);        // HKY frequencies
        prefetch_l1(prm->subst + 32);   // eigen-system frequencies, pInv, substKind
        prefetch_l1(A.catProps + d.catBase);
#pragma unroll
        for (int r = 0; r < R; r++) if (valid[r]) prefetch_l1(A.weights + d.patBase + pbase + 32 * r);
    }
    if (tipsInSmem) mbar_wait(tipBar, 0);   // the tip-state tile has landed

    // One pruning op.  `fwd` holds the previous op's (scaled) result and receives this op's; `park` is the register parking
    // slot (OPK_REG).  Returns true when the result is parked in registers (OP_REG_SLOT): the caller swaps the sets' roles.
    // Software pipelining across ops: the op's descriptor `o` and the state bytes of its tip operands (sa, sb) were fetched
    // while the previous op was computing; this op in turn fetches the state bytes of `on`, the next op of the chunk.
    auto tip_states = [&](const Op &q, int (&ta)[R], int (&tb)[R]) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            if ((q.ctl & 7) == OPK_TIP)
                ta[r] = tipsInSmem ? sstate[q.aOff * tilePat + stateIdx + 32 * r] : states[(size_t)q.aOff * d.stateStride + pbase + 32 * r];
            if (((q.ctl >> 3) & 7) == OPK_TIP)
                tb[r] = tipsInSmem ? sstate[q.bOff * tilePat + stateIdx + 32 * r] : states[(size_t)q.bOff * d.stateStride + pbase + 32 * r];
        }
    };
    auto op_body = [&](const Op &o, const OpRec *rec, const Op &on, int (&sa)[R], int (&sb)[R],
                       double (&fwd)[R][4], int (&fwdCum)[R], double (&park)[R][4], int (&parkCum)[R]) -> bool {
        const double *mA = rec->m[0], *mB = rec->m[1];
        const unsigned aKind = o.ctl & 7, bKind = (o.ctl >> 3) & 7;
        double out[R][4];
        int cum[R];
#pragma unroll
        for (int r = 0; r < R; r++) cum[r] = 0;

        // a tip's factor: column `state` of its matrix.  k_prepare stores a tip's matrix transposed (a column = 32 contiguous
        // bytes = two LDS.128 that the whole warp shares bank-conflict free) with a fifth column of ones for missing /
        // ambiguous data (state 4: the engine normalises every such code to 4), so the state byte IS the column index.
        auto tip_column = [&](int st, const double *m, double2 &v01, double2 &v23) {
            v01 = *reinterpret_cast<const double2 *>(m + st * 4);
            v23 = *reinterpret_cast<const double2 *>(m + st * 4 + 2);
        };
        // a parked (shared-memory stack / helper mailbox) or resident (HBM) operand
        auto fetch = [&](unsigned kind, unsigned off, unsigned slot, double (&a)[R][4]) {
            if (kind == OPK_STACK || (NW > 1 && kind == OPK_REMOTE)) {
                int si0 = slot * slotStride + stackIdx;
                if (NW > 1 && kind == OPK_REMOTE) {   // a helper's finished subtree: wait for its mailbox flag, read its slot
                    const int src = (int)off * inner + wi;
                    const unsigned *flag = sflag + slot * warps + src;
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
        };

        // ---- operand a (canonical order: a tip whenever the op has one; never the register-forwarded result) ----
        if (aKind == OPK_TIP) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                double2 v01, v23;
                tip_column(sa[r], mA, v01, v23);
                out[r][0] = v01.x; out[r][1] = v01.y; out[r][2] = v23.x; out[r][3] = v23.y;
            }
        } else if (aKind == OPK_REG) {
#pragma unroll
            for (int r = 0; r < R; r++) cum[r] += parkCum[r];
            rows_times<EXACT, R, true>(mA, park, out);
        } else {
            double a[R][4];
            fetch(aKind, o.aOff, (o.ctl >> 6) & 15, a);
            rows_times<EXACT, R, true>(mA, a, out);
        }
        int na[R], nb[R];
#pragma unroll
        for (int r = 0; r < R; r++) { na[r] = 4; nb[r] = 4; }
        tip_states(on, na, nb);   // the next op's tip states: in flight while this op finishes
        // ---- operand b: the previous op's result forwarded in registers, or a second tip, or (incremental walks) a clean node ----
        if (bKind == OPK_PREV) {
#pragma unroll
            for (int r = 0; r < R; r++) cum[r] += fwdCum[r];
            rows_times<EXACT, R, false>(mB, fwd, out);
        } else if (bKind == OPK_TIP) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                double2 v01, v23;
                tip_column(sb[r], mB, v01, v23);
                out[r][0] = __dmul_rn(out[r][0], v01.x); out[r][1] = __dmul_rn(out[r][1], v01.y);
                out[r][2] = __dmul_rn(out[r][2], v23.x); out[r][3] = __dmul_rn(out[r][3], v23.y);
            }
        } else {
            double a[R][4];
            fetch(bKind, o.bOff, (o.ctl >> 10) & 15, a);
            rows_times<EXACT, R, false>(mB, a, out);
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
            fwd[r][0] = o0; fwd[r][1] = o1; fwd[r][2] = o2; fwd[r][3] = o3;
            fwdCum[r] = ce;
        }
        if (NW > 1 && (o.ctl & OP_SIGNAL)) {   // hand the finished subtree to worker 0
            __syncwarp();
            if (lane == 0) st_release_shared(sflag + outSlot * warps + w, 1u);
        }
#pragma unroll
        for (int r = 0; r < R; r++) { sa[r] = na[r]; sb[r] = nb[r]; }
        return outSlot == (int)OP_REG_SLOT;
    };

    bool flip = false;
    for (int j = 0; j < nChunks; j++) {
        mbar_wait(myBar + (j & 1), (j >> 1) & 1);          // chunk j (descriptors and matrix pairs of up to CHUNK ops) has landed
        const OpRec *recs = myRing + (j & 1) * CHUNK;
        const int n = min(CHUNK, kEnd - kBeg - j * CHUNK);
        if (incremental) {                                  // clean children live in HBM: pull the chunk's sectors towards L2
            for (int i = 0; i < n; i++) {
                const Op q = recs[i].op;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if ((q.ctl & 7) == OPK_GLOBAL && valid[r]) prefetch_l2(partials + (size_t)q.aOff * 4 + r * 128);
                    if (((q.ctl >> 3) & 7) == OPK_GLOBAL && valid[r]) prefetch_l2(partials + (size_t)q.bOff * 4 + r * 128);
                }
            }
        }
        Op o = recs[0].op;
        int sa[R], sb[R];
#pragma unroll
        for (int r = 0; r < R; r++) { sa[r] = 4; sb[r] = 4; }
        tip_states(o, sa, sb);
#pragma unroll 1
        for (int i = 0; i < n; i++) {
            const Op on = recs[min(i + 1, n - 1)].op;       // the next op's descriptor (the last op of a chunk re-reads its own)
            // the op itself, on (forward, parked) = (X, Y) or (Y, X): a result parked in registers simply stays where it was
            // written and the two register sets swap roles -- no copies
            const bool parked = flip ? op_body(o, recs + i, on, sa, sb, yv, yc, xv, xc) : op_body(o, recs + i, on, sa, sb, xv, xc, yv, yc);
            flip ^= parked;
            o = on;
        }
        __syncwarp();                                       // every lane is done with chunk j: its buffer is free
        if (j + 2 < nChunks) issue_chunk(kBeg, j + 2);
    }
    for (int j = nChunks; j < issued; j++) mbar_wait(myBar + j, 0);   // speculative fetches beyond a short op list must land before the CTA retires
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
}

template <bool EXACT, int R, int MINB, int NC, int CH, int DEPTH, int NW = 1>
static cudaError_t launch_one(const WalkArgs &a, int nTiles, size_t smem, int threads, cudaStream_t st)
{
    static size_t configured[MAX_DEVICES] = {};
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
        else if (std4 && a.nCat == 4 && a.chunksPerTile == 1) return launch_one<EXACT, 2, 4, 4, 1, 4>(a, nTiles, smem, threads, st);
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
