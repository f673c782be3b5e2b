// k_step: ONE launch per posterior evaluation for the latency-bound shapes (a single changed locus per MCMC step --
// "regime B" -- and small problems with every locus dirty), instead of  k_prepare -> k_treewalk -> k_reduce_*.
//
// One thread-block CLUSTER per locus, one CTA per 8-warp tile of site patterns.  What the three-kernel path hands from
// kernel to kernel through HBM and two grid-wide dependencies stays inside the cluster:
//   * every CTA of the cluster runs the locus's tree-shaped work for itself (prepare_body.cuh: embedding in the species tree,
//     StarBeastClock rates, dirtiness, post-order schedule, transition matrices) straight into its OWN shared memory -- it is
//     a few microseconds of latency however many CTAs repeat it, and repeating it needs no exchange at all;
//   * only the cluster's leader (rank 0) writes the locus's state to HBM, and only after a hardware cluster barrier tells it
//     that every CTA has read what it needs of the old state (followers arrive and move on: nobody but the leader waits);
//   * each CTA then walks the schedule for its tile of patterns out of shared memory (same arithmetic, bit for bit, as
//     k_treewalk: walk_common.cuh) and sends its 32-pattern chunk sums to the leader through distributed shared memory;
//   * after the second cluster barrier the leader adds the chunk sums in pattern order (the order k_reduce_* uses);
//   * the closing reductions (reduce_body.cuh, unchanged) run in the same launch: by the leader itself when a single locus
//     ran -- every other locus's cached values are already in HBM -- or by the leader that takes the last ticket when all ran.
//     With peers connected that CTA also does the all-reduce over NVLink peer memory.
//
// The two paths share the engine's whole state (state blocks, partials, exponents, matrices), so evaluations may alternate
// freely between them; tests/test_gpu_fused.py checks that they agree bit for bit.
//
// Compiled with -fmad=false like prepare.cu (the reference's scalar arithmetic lives in prepare_body.cuh / reduce_body.cuh);
// the pruning arithmetic uses explicit fma (walk_common.cuh), so it does not depend on that flag.
#include "prepare_body.cuh"
#include "reduce_body.cuh"
#include "walk_common.cuh"

namespace sb2 {

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_count_x() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
// store a double into the same shared-memory variable of another CTA of the cluster (distributed shared memory)
__device__ __forceinline__ void st_cluster_f64(double *localAddr, unsigned rank, double v)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(localAddr);
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(r), "d"(v) : "memory");
}

struct StepSmem {
    size_t prep, sops, smd, st2, sexp, schunk, sstate, total;
    int onesIdx, planeStride;
};

__host__ __device__ inline StepSmem step_carve(int maxNodes, int S, int C, int depth, int maxTips, int tilePat)
{
    StepSmem m;
    const int maxInt = (maxNodes - 1) / 2;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 31) & ~size_t(31); return r; };
    m.prep = take(prep_carve(nullptr, nullptr, maxNodes, S));
    m.sops = take((size_t)maxInt * sizeof(Op));
    m.onesIdx = maxInt * C * 32;
    m.smd = take(((size_t)m.onesIdx + 16) * 8);                       // matrices [op][category][side][16] | ones[16]
    m.planeStride = depth * STEP_WARPS * 32;
    m.st2 = take(2ull * m.planeStride * 16);                          // stack: [2 planes][slot][warp][32] double2
    m.sexp = take((size_t)m.planeStride * 2);
    m.schunk = take((size_t)STEP_MAX_CLUSTER * STEP_WARPS * 8);       // the locus's chunk sums (used in the leader only)
    m.sstate = take((size_t)maxTips * tilePat + 16);
    m.total = o;
    return m;
}

size_t step_smem_bytes(int maxNodes, int S, int nCat, int stackDepth, int maxTips)
{
    const int chunks = STEP_WARPS / nCat;
    return step_carve(maxNodes, S, nCat, stackDepth, maxTips, chunks * 32).total;
}

// The walk of one tile: the schedule and the matrices come from shared memory, everything else is k_treewalk's R = 1 walk.
__device__ __forceinline__ void step_walk(const StepArgs &A, const LocusDesc &d, const LocusParams *prm, int tile, int nOps, const StepSmem &m,
                                          unsigned char *smem_raw, double *chunkOut, int weight, int constMask)
{
    const int C = A.nCat, chunks = STEP_WARPS / C, depth = A.stackDepth;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const bool active = w < C * chunks;                 // category counts that do not divide 8 leave warps idle
    const int c = w % C, chunk = w / C;
    const int tilePat = chunks * 32, nPat = d.nPat;
    const int p0 = tile * tilePat;
    const int p = p0 + chunk * 32 + lane;
    const bool valid = active && p < nPat;
    const Op *sops = reinterpret_cast<const Op *>(smem_raw + m.sops);
    const double *smd = reinterpret_cast<const double *>(smem_raw + m.smd);
    double2 *st2 = reinterpret_cast<double2 *>(smem_raw + m.st2);
    int16_t *sexp = reinterpret_cast<int16_t *>(smem_raw + m.sexp);
    const uint8_t *sstate = smem_raw + m.sstate;
    const int planeStride = m.planeStride, onesIdx = m.onesIdx;
    constexpr int slotStride = STEP_WARPS * 32;
    const int stackIdx = w * 32 + lane, stateIdx = chunk * 32 + lane;
    double *partials = A.partials + d.partBase + ((size_t)c * d.patStride + p) * 4;
    int16_t *cumExp = A.cumExp + d.expBase + (size_t)c * d.patStride + p;

    if (active) {
        if (nOps < d.nInt && valid) {
            // incremental: clean siblings live in HBM / L2 -- pull all of them into L1 before the dependent chain starts, and
            // those of the partner warp in the other half of the CTA too (the halves reach the walk at different times: whoever
            // comes first fetches for both)
            const int dp = (C == 1 || C == 2 || C == 4) ? (((w ^ (STEP_WARPS / 2)) - w) / C) * 32 : 0;   // same category, partner chunk
            for (int k = 0; k < nOps; k++) {
                const Op q = sops[k];
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    if (((q.ctl >> (3 * side)) & 7) != OPK_GLOBAL) continue;
                    const size_t off = side ? q.bOff : q.aOff;
                    prefetch_l1(partials + off * 4); prefetch_l1(cumExp + off);
                    if (dp != 0 && p + dp >= 0 && p + dp < nPat) prefetch_l1(partials + (off + dp) * 4);
                }
            }
        }
        double prev[1][4] = {{0.0, 0.0, 0.0, 0.0}};
        int prevCum = 0;
        for (int k = 0; k < nOps; k++) {
            const Op o = sops[k];
            const int mIdx = (k * C + c) * 32;
            double out[1][4];
            int cum = 0;
#pragma unroll
            for (int side = 0; side < 2; side++) {
                const unsigned kind = (o.ctl >> (3 * side)) & 7;
                const unsigned off = side ? o.bOff : o.aOff;
                const int mm = mIdx + side * 16;
                if (kind == OPK_TIP) {
                    const int st = sstate[off * tilePat + stateIdx];
                    const int col = (st < 4) ? (mm + st) : onesIdx;   // missing / ambiguous: factor 1
                    if (side == 0) { out[0][0] = smd[col]; out[0][1] = smd[col + 4]; out[0][2] = smd[col + 8]; out[0][3] = smd[col + 12]; }
                    else {
                        out[0][0] = __dmul_rn(out[0][0], smd[col]); out[0][1] = __dmul_rn(out[0][1], smd[col + 4]);
                        out[0][2] = __dmul_rn(out[0][2], smd[col + 8]); out[0][3] = __dmul_rn(out[0][3], smd[col + 12]);
                    }
                } else if (kind == OPK_PREV) {
                    cum += prevCum;
                    if (side == 0) rows_times<false, 1, true>(smd + mm, prev, out);
                    else rows_times<false, 1, false>(smd + mm, prev, out);
                } else {
                    double a[1][4] = {{0.0, 0.0, 0.0, 0.0}};
                    if (kind == OPK_STACK) {
                        const int si = ((o.ctl >> (6 + 4 * side)) & 15) * slotStride + stackIdx;
                        const double2 lo = st2[si], hi = st2[planeStride + si];
                        a[0][0] = lo.x; a[0][1] = lo.y; a[0][2] = hi.x; a[0][3] = hi.y;
                        cum += sexp[si];
                    } else if (valid) {
                        const double *src = partials + (size_t)off * 4;
                        if (kind == OPK_GLOBAL) ld256_nc(src, a[0][0], a[0][1], a[0][2], a[0][3]);
                        else ld256_coherent(src, a[0][0], a[0][1], a[0][2], a[0][3]);   // OPK_SPILL: written by this thread in this launch
                        cum += cumExp[(size_t)off];
                    }
                    if (side == 0) rows_times<false, 1, true>(smd + mm, a, out);
                    else rows_times<false, 1, false>(smd + mm, a, out);
                }
            }
            double o0, o1, o2, o3;
            const int ce = cum + rescale4(out[0], o0, o1, o2, o3);
            const int outSlot = (o.ctl >> 14) & 255;
            if (outSlot < depth) {
                const int so = outSlot * slotStride + stackIdx;
                st2[so] = make_double2(o0, o1);
                st2[planeStride + so] = make_double2(o2, o3);
                sexp[so] = (int16_t)ce;
            }
            if (valid) {
                st256(partials + (size_t)o.outOff * 4, o0, o1, o2, o3);
                cumExp[(size_t)o.outOff] = (int16_t)ce;
            }
            prev[0][0] = o0; prev[0][1] = o1; prev[0][2] = o2; prev[0][3] = o3;
            prevCum = ce;
        }
    }
    __syncthreads();   // every category's root sits in stack slot 0
    if (active && c == 0) {
        const double *freqs = (prm->substKind == 0) ? prm->subst + 1 : prm->subst + 36;   // the CTA's shared-memory copy
        const double *props = A.catProps + d.catBase;
        const double pInv = prm->pInv;
        double contrib = 0.0;
        if (valid) {
            const int mask = (pInv > 0.0) ? constMask : 0;
            const double lk = root_pattern_logl<false>(st2, planeStride, sexp, chunk * C * 32 + lane, 32, C, props, freqs, pInv, mask);
            A.patLogL[d.patBase + p] = lk;
            contrib = __dmul_rn(lk, (double)weight);
        }
        contrib = chunk_sum(contrib);
        const int ch = (p0 >> 5) + chunk;
        if (lane == 0 && ch < d.nChunks) st_cluster_f64(chunkOut + ch, 0, contrib);   // into the leader's shared memory
    }
}

__global__ void __launch_bounds__(STEP_THREADS, 2) k_step(const __grid_constant__ StepArgs A)
{
    extern __shared__ __align__(32) unsigned char smem_raw[];
    __shared__ LocusParams sprm;
    __shared__ int s_nOps;
    __shared__ unsigned s_last;
    const int tid = threadIdx.x;
    const unsigned rank = cluster_ctarank(), cid = cluster_id_x();
    auto stamp = [&](int i) {   // profile mode only: where the time of one step goes (sb2_step_phase_times)
        if (A.dbg && tid == 0 && rank == 0 && cid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); A.dbg[i] = t; }
    };
    stamp(0);
    const bool leader = rank == 0;
    const bool single = A.nRun > 0;
    const int l = single ? A.runLoci[cid] : (int)cid;
    const uint8_t fl = single ? (uint8_t)(cid == 0 ? A.runFlags : 0)
                              : (A.prep.uniformFlags >= 0 ? (uint8_t)A.prep.uniformFlags : A.prep.flags[l]);
    bool mine = false;   // lazy store(): this locus's slices of the stored block are brought up to date by its leader
    for (int j = 0; j < A.prep.nSync; j++) mine |= A.prep.syncLoci[j] == l;
    const bool run = (fl & LF_RUN) != 0;
    const LocusDesc d = A.prep.loci[l];
    const int C = A.nCat, tilePat = (STEP_WARPS / C) * 32;
    const StepSmem m = step_carve(A.maxNodes, A.prep.S, C, A.stackDepth, A.maxTips, tilePat);
    double *schunk = reinterpret_cast<double *>(smem_raw + m.schunk);
    const bool hasTile = run && (int)rank * tilePat < d.nPat;

    if (run && !hasTile) {   // a CTA beyond this locus's last tile (the cluster is sized for the widest locus): barriers only
        cluster_arrive_release(); cluster_wait_acquire();
        cluster_arrive_release(); cluster_wait_acquire();
        return;
    }
    int weight = 0, constMask = 0;
    if (run) {
        {   // static inputs of the root epilogue, per thread: fetched now, needed after the walk
            const int C2 = A.nCat, w = tid >> 5, chunk = w / C2, p = (int)rank * tilePat + chunk * 32 + (tid & 31);
            if (w % C2 == 0 && w < C2 * (STEP_WARPS / C2) && p < d.nPat) { weight = A.weights[d.patBase + p]; constMask = A.constMask[d.patBase + p]; }
            if (tid < C2) prefetch_l1(A.catProps + d.catBase + tid);
        }
        if (hasTile) {   // the tile's tip states (static input): asynchronous 16-byte copies that land while the tree work runs
            uint8_t *sstate = smem_raw + m.sstate;
            const uint8_t *states = A.tipStates + d.stateBase + (size_t)rank * tilePat;
            const int per = min(tilePat, d.stateStride - (int)rank * tilePat) >> 4, total = d.nTips * per;   // never past the padded row
            for (int idx = tid; idx < total; idx += STEP_THREADS) {
                const int t = idx / per, q = idx - t * per;
                cp_async16(sstate + t * tilePat + q * 16, states + (size_t)t * d.stateStride + q * 16);
            }
        }
        cp_async_commit();
        if (tid < 16) reinterpret_cast<double *>(smem_raw + m.smd)[m.onesIdx + tid] = 1.0;
        if (mine && leader) sync_stored_slices(A.prep, l);   // reads cur, writes stored; nobody reads this locus's stored slices in this launch
        PrepFused F;
        F.leader = leader; F.storedIsCur = mine;
        F.sops = reinterpret_cast<Op *>(smem_raw + m.sops);
        F.smats = reinterpret_cast<double *>(smem_raw + m.smd);
        F.nOpsOut = &s_nOps;
        F.dbg = (leader && cid == 0) ? A.dbg : nullptr;
        F.treeSrc = (single && cid == 0 && A.inlineTreeBytes > 0) ? A.inlineTree : nullptr;
        F.splitReturn = true;
        // cluster barrier 1 inside: all arrive, the leader waits
        if (prepare_locus<true>(A.prep, A.maxNodes, l, fl, smem_raw + m.prep, sprm, F) == PREP_HALVES && tid < STEP_THREADS / 2)
            asm volatile("bar.sync 4, %0;" ::"n"(STEP_THREADS) : "memory");   // the likelihood half's schedule and matrices
        stamp(1);
        const int nOps = s_nOps;
        if (hasTile && nOps > 0) step_walk(A, d, &sprm, (int)rank, nOps, m, smem_raw, schunk, weight, constMask);
        stamp(2);
        if (!leader) cluster_wait_acquire();   // barrier 1 (completed long ago)
        cluster_arrive_release();              // barrier 2: the chunk sums are in the leader's shared memory
        cluster_wait_acquire();
        if (!leader) return;
        stamp(3);
        if (nOps > 0 && tid == 0) {            // the locus's log-likelihood: chunk sums in pattern order, as k_reduce_* adds them
            double lik = 0.0;
            for (int ch = 0; ch < d.nChunks; ch++) lik += schunk[ch];
            A.prep.cur.logL[l] = lik;
        }
    } else {
        if (!leader) return;
        if (mine) sync_stored_slices(A.prep, l);
        if (!single && tid == 0) A.prep.nOps[l] = 0;
    }
    // ---- closing reductions, leaders only -----------------------------------------------------------------------------------
    if (single) {
        if (cid != 0) return;                  // sync-only clusters
        __syncthreads();                       // this CTA's own writes of the locus's new values
    } else {
        __threadfence();                       // this locus's values are visible device-wide before the ticket is taken
        __syncthreads();
        if (tid == 0) {
            const unsigned ticket = atomicAdd(A.doneCounter, 1u);
            s_last = (ticket == cluster_count_x() - 1);
            if (s_last) *A.doneCounter = 0;
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
    }
    stamp(4);
    double sums3[3];
    reduce_local_body(A.red, -1, 0, sums3);    // red.useTiles == 0: per-locus values come from the state block
    if (A.useComm || A.red.popKind == 2) { __threadfence_block(); __syncthreads(); }   // sums that travel through the reduce vector
    stamp(6);
    if (A.useComm) xgpu_exchange_body(A.red, A.comm);
    finish_body(A.red, A.useComm ? nullptr : sums3);
    stamp(7);
    publish_body(A.red);
    stamp(5);
}

cudaError_t launch_step(const StepArgs &a, int nClusters, int clusterSize, cudaStream_t st)
{
    const size_t smem = step_smem_bytes(a.maxNodes, a.prep.S, a.nCat, a.stackDepth, a.maxTips);
    static size_t configured[MAX_DEVICES + 1] = {};
    static bool nonPortable[MAX_DEVICES] = {};
    cudaError_t rc = ensure_dynamic_smem(k_step, smem, configured);
    if (rc != cudaSuccess) return rc;
    if (clusterSize > 8) {
        int dev = 0;
        if ((rc = cudaGetDevice(&dev)) != cudaSuccess) return rc;
        if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
        if (!nonPortable[dev]) {
            if ((rc = cudaFuncSetAttribute(k_step, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) return rc;
            nonPortable[dev] = true;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nClusters * clusterSize); cfg.blockDim = dim3(STEP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = clusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_step, a);
}

}  // namespace sb2
