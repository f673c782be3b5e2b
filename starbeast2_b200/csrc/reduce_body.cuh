// Device bodies of the reductions (see reduce.cu for the description); shared by the stand-alone reduction kernels
// and by k_treewalk, whose last-finishing CTA runs them in place (one launch fewer on the latency path).
// Must be compiled with -fmad=false semantics where it matters: all sums here are plain adds.
#pragma once
#include <math.h>

#include "sb2_device.cuh"

namespace sb2 {

// deterministic block sum: fixed assignment of elements to threads, fixed-shape tree over threads
__device__ __forceinline__ double block_sum(double v, double *scratch)
{
    const int tid = threadIdx.x;
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (tid < 32) {
        t = (tid < (blockDim.x >> 5)) ? scratch[tid] : 0.0;
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    }
    __syncthreads();
    if (tid == 0) scratch[0] = t;
    __syncthreads();
    t = scratch[0];
    __syncthreads();
    return t;
}

// three sums in one pass (same fixed shape as block_sum, a third of the barriers)
__device__ __forceinline__ void block_sum3(double &a, double &b, double &c, double *scratch)
{
    const int tid = threadIdx.x, nw = blockDim.x >> 5;
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); c += __shfl_xor_sync(0xffffffffu, c, off);
    }
    __syncthreads();
    if ((tid & 31) == 0) { scratch[tid >> 5] = a; scratch[32 + (tid >> 5)] = b; scratch[64 + (tid >> 5)] = c; }
    __syncthreads();
    double x = 0.0, y = 0.0, z = 0.0;
    if (tid < 32) {
        if (tid < nw) { x = scratch[tid]; y = scratch[32 + tid]; z = scratch[64 + tid]; }
        for (int off = 16; off > 0; off >>= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, off); y += __shfl_xor_sync(0xffffffffu, y, off); z += __shfl_xor_sync(0xffffffffu, z, off);
        }
    }
    __syncthreads();
    if (tid == 0) { scratch[0] = x; scratch[1] = y; scratch[2] = z; }
    __syncthreads();
    a = scratch[0]; b = scratch[1]; c = scratch[2];
    __syncthreads();
}

// one double of the result vector towards the host (see ReduceArgs::hostTag)
__device__ __forceinline__ void host_put(const ReduceArgs &A, int idx, double v)
{
    if (A.hostTag) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        volatile unsigned long long *w = A.hostWords + 2 * (size_t)idx;
        w[0] = A.hostTag | (bits & 0xffffffffull);
        w[1] = A.hostTag | (bits >> 32);
    } else if (A.hostOut) {
        A.hostOut[idx] = v;
    }
}

// preChunkBase / preNChunks: chunkBase / nChunks of locus threadIdx.x, fetched by the caller before it waited for the pruning
// kernel (static data: one dependent round trip less on the latency path); preChunkBase < 0 = not provided
// sums3 (optional): receives {coalescent, likelihood, nIncompatible} in EVERY thread, so that a finish_body in the same CTA
// need not read them back from the reduce vector in HBM
__device__ __forceinline__ void reduce_local_body(const ReduceArgs &A, int preChunkBase = -1, int preNChunks = 0, double *sums3 = nullptr)
{
    __shared__ double scratch[96];
    const int tid = threadIdx.x, nt = blockDim.x, L = A.nLoci, S = A.S;
    double *perLocusCoal = A.out + 3, *perLocusLik = A.out + 3 + L;
    double likPart = 0.0, coalPart = 0.0, incPart = 0.0;
    for (int l = tid; l < L; l += nt) {
        double lik;
        const double coal = __ldcg(A.cur.coal + l);   // issued before the dependent chunk-sum loads
        if (A.useTiles && A.nOps[l] > 0) {
            const LocusDesc &d = A.loci[l];
            lik = 0.0;
            if (A.exact) {
                for (int p = 0; p < d.nPat; p++) lik = __dadd_rn(lik, __dmul_rn(__ldcg(A.patLogL + d.patBase + p), (double)A.weights[d.patBase + p]));   // calcLogP order
            } else {
                // the locus's 32-pattern chunk sums, added in pattern order (the order does not depend on the launch
                // geometry, and the fused step of step.cu adds the same numbers the same way)
                const bool pre = preChunkBase >= 0 && l == tid;
                const int chunkBase = pre ? preChunkBase : d.chunkBase, nChunks = pre ? preNChunks : d.nChunks;
                for (int t0 = 0; t0 < nChunks; t0 += 8) {   // loads in flight together, adds in chunk order
                    double v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = (t0 + j < nChunks) ? __ldcg(A.chunkSum + chunkBase + t0 + j) : 0.0;
#pragma unroll
                    for (int j = 0; j < 8; j++) if (t0 + j < nChunks) lik += v[j];
                }
            }
            A.cur.logL[l] = lik;
        } else {
            lik = __ldcg(A.cur.logL + l);   // clean locus: cached value (CompoundDistribution uses getCurrentLogP)
        }
        perLocusLik[l] = lik;
        perLocusCoal[l] = coal;
        host_put(A, 3 + L + l, lik); host_put(A, 3 + l, coal);   // zero-copy result vector, written as soon as known
        likPart += lik;
        coalPart += coal;
        incPart += (coal == -INFINITY) ? 1.0 : 0.0;
    }
    double likSum, coalSum;
    if (A.exact) {
        __syncthreads();
        likSum = 0.0; coalSum = 0.0;
        if (tid == 0) {
            for (int l = 0; l < L; l++) { likSum += perLocusLik[l]; coalSum += perLocusCoal[l]; }
            scratch[0] = likSum; scratch[1] = coalSum;
        }
        __syncthreads();
        likSum = scratch[0]; coalSum = scratch[1];
        __syncthreads();
        incPart = block_sum(incPart, scratch);
    } else {
        block_sum3(likPart, coalPart, incPart, scratch);
        likSum = likPart; coalSum = coalPart;
    }
    const double inc = incPart;
    if (tid == 0) { A.red[0] = coalSum; A.red[1] = likSum; A.red[2] = inc; }
    if (sums3) { sums3[0] = coalSum; sums3[1] = likSum; sums3[2] = inc; }

    // analytic integration: per-branch sums over this rank's loci
    if (A.popKind == 2 && A.exact) {   // the reference's order: loci in list order, one branch at a time
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
        for (int b = warp; b < S; b += nw) {
            double q = 0.0, g = 0.0, r = 0.0;
            if (lane == 0) {
                for (int l = 0; l < L; l++) {
                    const size_t o = (size_t)l * S + b;
                    q += __ldcg(A.cur.brK + o); g += __ldcg(A.cur.anG + o); r += __ldcg(A.cur.anR + o);
                }
                A.red[3 + 3 * b] = q; A.red[3 + 3 * b + 1] = g; A.red[3 + 3 * b + 2] = r;
            }
        }
    } else if (A.popKind == 2) {
        // The three arrays are [locus][branch]: a warp reads whole rows (lanes = 32 consecutive branches: coalesced), warp w
        // the loci l = w (mod warps), several rows in flight; the warps' partial sums then meet in shared memory and are
        // added in warp order.  Fixed shape: deterministic, whatever runs before or after.
        __shared__ double part[3][8][64];
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;   // nw <= 8 (RED_THREADS / STEP_THREADS = 256)
        for (int b0 = 0; b0 < S; b0 += 64) {   // two blocks of 32 branches per pass: twice the loads in flight per dependent step
            const int bA = b0 + lane, bB = b0 + 32 + lane;
            const bool hasA = bA < S, hasB = bB < S;
            const int cA = hasA ? bA : 0, cB = hasB ? bB : 0;   // clamped: loads stay unconditional, results are masked
            double qA = 0.0, gA = 0.0, rA = 0.0, qB = 0.0, gB = 0.0, rB = 0.0;
            int l = warp;
            for (; l + 3 * nw < L; l += 4 * nw) {   // four rows in flight
                double vq[4], vg[4], vr[4], wq[4], wg[4], wr[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const size_t o = (size_t)(l + j * nw) * S;
                    vq[j] = (double)__ldcg(A.cur.brK + o + cA); vg[j] = __ldcg(A.cur.anG + o + cA); vr[j] = __ldcg(A.cur.anR + o + cA);
                    wq[j] = (double)__ldcg(A.cur.brK + o + cB); wg[j] = __ldcg(A.cur.anG + o + cB); wr[j] = __ldcg(A.cur.anR + o + cB);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) { qA += vq[j]; gA += vg[j]; rA += vr[j]; qB += wq[j]; gB += wg[j]; rB += wr[j]; }
            }
            for (; l < L; l += nw) {
                const size_t o = (size_t)l * S;
                qA += (double)__ldcg(A.cur.brK + o + cA); gA += __ldcg(A.cur.anG + o + cA); rA += __ldcg(A.cur.anR + o + cA);
                qB += (double)__ldcg(A.cur.brK + o + cB); gB += __ldcg(A.cur.anG + o + cB); rB += __ldcg(A.cur.anR + o + cB);
            }
            __syncthreads();   // the previous pair of blocks has been consumed
            part[0][warp][lane] = qA; part[1][warp][lane] = gA; part[2][warp][lane] = rA;
            part[0][warp][32 + lane] = qB; part[1][warp][32 + lane] = gB; part[2][warp][32 + lane] = rB;
            __syncthreads();
            if (tid < 192) {
                const int a = tid >> 6, bb = b0 + (tid & 63);
                double sum = 0.0;
                for (int w = 0; w < nw; w++) sum += part[a][w][tid & 63];
                if (bb < S) A.red[3 + 3 * bb + a] = sum;
            }
        }
    } else {
        for (int i = tid; i < 3 * S; i += nt) A.red[3 + i] = 0.0;
    }
}

constexpr int FINISH_SMEM_TERMS = 512;   // per-branch terms kept in shared memory for the final in-order sum (larger species trees: the rest from HBM)
// sums3 (optional): the three leading entries of the reduce vector, already in registers (single-GPU, same CTA)
__device__ __forceinline__ void finish_body(const ReduceArgs &A, const double *sums3 = nullptr)
{
    const int tid = threadIdx.x, nt = blockDim.x, L = A.nLoci, S = A.S;
    double *perBranch = A.out + 3 + 2 * L;
    double coal = sums3 ? sums3[0] : A.red[0];
    const double lik = sums3 ? sums3[1] : A.red[1];
    const bool incompatible = (sums3 ? sums3[2] : A.red[2]) > 0.0;
    if (A.popKind == 2) {
        __shared__ double sterm[FINISH_SMEM_TERMS];   // per-branch terms for the final in-order sum (larger species trees: the rest from HBM)
        __shared__ int smiss[FINISH_SMEM_TERMS];      // branches whose term has to be computed
        __shared__ int s_nMiss;
        const double alpha = A.cur.hyper[0], beta = A.cur.hyper[1];
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
        const bool useMemo = A.memo != nullptr && S <= FINISH_SMEM_TERMS;   // (branch numbers travel in 16 bits below)
        if (tid == 0) s_nMiss = 0;
        __syncthreads();
        // Phase 1, one THREAD per branch: has this branch's term been computed for exactly these values before?  (keys compared as
        // bit patterns.)  Every branch's loads are in flight together: ONE round trip for the whole species tree, where a warp
        // per branch was one round trip per branch.
        if (useMemo) {
            const long long ka = __double_as_longlong(alpha), kb = __double_as_longlong(beta);
            for (int b = tid; b < S; b += nt) {
                double *m = A.memo + (size_t)MEMO_STRIDE * b;
                const long long kq = __double_as_longlong(A.red[3 + 3 * b]), kg = __double_as_longlong(A.red[3 + 3 * b + 1]),
                                kr = __double_as_longlong(A.red[3 + 3 * b + 2]);
                int hit = -1;
#pragma unroll
                for (int en = 1; en >= 0; en--) {
                    const double *k = m + 6 * en;
                    if (__double_as_longlong(k[0]) == kq && __double_as_longlong(k[1]) == kg && __double_as_longlong(k[2]) == kr &&
                        __double_as_longlong(k[3]) == ka && __double_as_longlong(k[4]) == kb) hit = en;
                }
                if (hit >= 0 && !A.memoFlush) {
                    const double logP = m[6 * hit + 5];
                    m[12] = (double)(1 - hit);   // the other entry is the next victim
                    perBranch[b] = logP; sterm[b] = logP; host_put(A, 3 + 2 * L + b, logP);
                } else {
                    // to be computed (any order: every branch's term is independent); the entry it will be written to is chosen
                    // here, where the memo's words are already in registers: the one that holds this key (forced recomputation),
                    // else the one that was not used last
                    const int v = hit >= 0 ? hit : ((m[12] == 1.0) ? 1 : 0);
                    smiss[atomicAdd(&s_nMiss, 1)] = b | (v << 16);
                }
            }
            __syncthreads();
        }
        // Phase 2, one WARP per branch that missed (all of them without the memo); no block barriers inside
        const int nWork = useMemo ? s_nMiss : S;
        for (int i = warp; i < nWork; i += nw) {
            const int b = useMemo ? (smiss[i] & 0xffff) : i, victim = useMemo ? (smiss[i] >> 16) : 0;
            const double qd = A.red[3 + 3 * b];
            const long long q = (long long)(qd + 0.5);
            const double g = A.red[3 + 3 * b + 1], r = A.red[3 + 3 * b + 2];
            double lgr = 0.0;   // logGammaRatio, :227-230
            if (A.exact) {
                if (lane == 0) for (long long i2 = 0; i2 < q; i2++) lgr += log(alpha + (double)i2);   // the reference's order
            } else {
                for (long long i2 = lane; i2 < q; i2 += 32) lgr += log(alpha + (double)i2);
                for (int off = 16; off > 0; off >>= 1) lgr += __shfl_xor_sync(0xffffffffu, lgr, off);
            }
            const double branchLogR = -r;
            // explicit roundings: the result must not depend on the translation unit's FMA contraction setting
            const double logP = __dadd_rn(__dsub_rn(__dadd_rn(branchLogR, __dmul_rn(alpha, log(beta))),
                                                    __dmul_rn(__dadd_rn(alpha, (double)q), log(__dadd_rn(beta, g)))), lgr);   // :232
            if (lane == 0) {
                perBranch[b] = logP; host_put(A, 3 + 2 * L + b, logP);
                if (b < FINISH_SMEM_TERMS) sterm[b] = logP;
                if (useMemo) {   // refill the entry chosen in phase 1
                    double *m = A.memo + (size_t)MEMO_STRIDE * b;
                    double *k = m + 6 * victim;
                    k[0] = qd; k[1] = g; k[2] = r; k[3] = alpha; k[4] = beta; k[5] = logP;
                    m[12] = (double)(1 - victim);
                }
            }
        }
        __syncthreads();
        // node-number order, one by one (MultispeciesCoalescent.java:194) -- from shared memory: read back from the result vector
        // in HBM, the 47 dependent adds of C4 were 47 dependent L2 round trips
        if (tid == 0) for (int b = 0; b < S; b++) coal += (b < FINISH_SMEM_TERMS) ? sterm[b] : perBranch[b];
    } else {
        for (int b = tid; b < S; b += nt) { perBranch[b] = 0.0; host_put(A, 3 + 2 * L + b, 0.0); }
    }
    if (incompatible) coal = -INFINITY;   // MultispeciesCoalescent.java:153
    if (tid == 0) {
        A.out[0] = coal; A.out[1] = lik; A.out[2] = coal + lik;
        host_put(A, 0, coal); host_put(A, 1, lik); host_put(A, 2, coal + lik);
    }
}


// All-reduce of the reduce vector over peer memory (see reduce.cu:k_reduce_xgpu); called by every thread of ONE CTA after
// reduce_local_body + __syncthreads; on return A.red holds the global sums (or NaN if a peer never showed up).
__device__ __forceinline__ void xgpu_exchange_body(const ReduceArgs &A, const CommArgs &Cm)
{
    const int tid = threadIdx.x, nt = blockDim.x, nR = Cm.nRanks, n = Cm.nRed;
    const int half = (int)(Cm.seq & 1ull);
    // Flag-in-data exchange: every double travels as two 8-byte words {32 data bits | 32-bit sequence tag}.  An 8-byte store
    // is atomic, so a receiver that sees the tag of this evaluation in a word also sees its data: no fence, no separate flag,
    // one one-way NVLink trip between "my sums are ready" and "I can add yours".
    const unsigned long long tag = (0x80000000ull | (Cm.seq & 0x7fffffffull)) << 32;
    for (int idx = tid; idx < nR * n; idx += nt) {
        const int r = idx / n, i = idx - r * n;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(A.red[i]);
        volatile unsigned long long *dst = Cm.peerSlots[r] + (((size_t)half * COMM_MAX_RANKS + Cm.rank) * n + i) * 2;
        dst[0] = tag | (bits & 0xffffffffull);
        dst[1] = tag | (bits >> 32);
    }
    __shared__ int s_poison;
    if (tid == 0) s_poison = 0;
    __syncthreads();
    unsigned long long waitT0 = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(waitT0));
    // One thread per (element, rank) word pair: every rank's words are polled at the same time (one L2 round trip for all of
    // them once they are there, instead of one per rank in turn); the ranks of an element sit in neighbouring lanes and are
    // added in rank order by shuffles -- every rank adds the same numbers in the same order.
    int nRp = 1;
    while (nRp < nR) nRp <<= 1;          // lanes per element: a power of two <= COMM_MAX_RANKS (16), so groups never straddle a warp
    const int lane = tid & 31;
    for (int base = 0; base < n * nRp; base += nt) {   // uniform trip count: all lanes take part in the shuffles
        const int p = base + tid, i = p / nRp, r = p - i * nRp;
        const bool act = i < n && r < nR;
        double v = 0.0;
        if (act) {
            const long long t0 = clock64();
            const volatile unsigned long long *src = Cm.peerSlots[Cm.rank] + (((size_t)half * COMM_MAX_RANKS + r) * n + i) * 2;
            unsigned long long w0, w1;
            for (;;) {
                w0 = src[0]; w1 = src[1];
                if ((w0 & 0xffffffff00000000ull) == tag && (w1 & 0xffffffff00000000ull) == tag) break;
                if (clock64() - t0 > 4000000000ll) { s_poison = 1; w0 = w1 = 0; break; }   // ~2 s at 2 GHz: a peer is gone
            }
            v = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        }
        __syncwarp();
        const int g0 = lane & ~(nRp - 1);
        double sum = __shfl_sync(0xffffffffu, v, g0);
        for (int rr = 1; rr < nR; rr++) sum += __shfl_sync(0xffffffffu, v, g0 + rr);
        if (act && r == 0) A.red[i] = sum;
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0 && A.hostOut) {   // time spent waiting for the peers' words (rank skew + one NVLink trip), for sb2_exchange_wait_us
        unsigned long long waitT1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(waitT1));
        A.hostOut[3 + 2 * A.nLoci + A.S] = (double)(waitT1 - waitT0) * 1e-3;
    }
    if (s_poison) { for (int i = tid; i < n; i += nt) A.red[i] = NAN; __threadfence_block(); __syncthreads(); }
}

// The result vector lives twice: in device memory (`out`) and in host-mapped memory (`hostOut`, zero-copy download: the
// host only has to synchronise the stream).  Both copies are written where the values are produced, so nothing is left to do
// here but make sure the block's writes are ordered before the kernel ends.
__device__ __forceinline__ void publish_body(const ReduceArgs &A)
{
    if (!A.hostTag) __threadfence_system();   // tagged words need no fence: each carries its own proof of arrival
}

}  // namespace sb2
