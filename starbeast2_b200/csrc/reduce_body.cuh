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

// preTileBase / preNTiles: tileBase / nTiles of locus threadIdx.x, fetched by the caller before it waited for the pruning
// kernel (static data: one dependent round trip less on the latency path); preTileBase < 0 = not provided
__device__ __forceinline__ void reduce_local_body(const ReduceArgs &A, int preTileBase = -1, int preNTiles = 0)
{
    __shared__ double scratch[96];
    const int tid = threadIdx.x, nt = blockDim.x, L = A.nLoci, S = A.S;
    double *perLocusCoal = A.out + 3, *perLocusLik = A.out + 3 + L;
    double likPart = 0.0, coalPart = 0.0, incPart = 0.0;
    for (int l = tid; l < L; l += nt) {
        double lik;
        const double coal = A.cur.coal[l];   // issued before the dependent tile loads
        if (A.useTiles && A.nOps[l] > 0) {
            const LocusDesc &d = A.loci[l];
            lik = 0.0;
            if (A.exact) {
                for (int p = 0; p < d.nPat; p++) lik = __dadd_rn(lik, __dmul_rn(__ldcg(A.patLogL + d.patBase + p), (double)A.weights[d.patBase + p]));   // calcLogP order
            } else {
                const bool pre = preTileBase >= 0 && l == tid;
                const int tileBase = pre ? preTileBase : d.tileBase, nTiles = pre ? preNTiles : d.nTiles;
                for (int t0 = 0; t0 < nTiles; t0 += 8) {   // loads in flight together, adds in tile order
                    double v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = (t0 + j < nTiles) ? __ldcg(A.tileSum + tileBase + t0 + j) : 0.0;
#pragma unroll
                    for (int j = 0; j < 8; j++) if (t0 + j < nTiles) lik += v[j];
                }
            }
            A.cur.logL[l] = lik;
        } else {
            lik = A.cur.logL[l];   // clean locus: cached value (CompoundDistribution uses getCurrentLogP)
        }
        perLocusLik[l] = lik;
        perLocusCoal[l] = coal;
        if (A.hostOut) { A.hostOut[3 + L + l] = lik; A.hostOut[3 + l] = coal; }   // zero-copy result vector, written as soon as known
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

    // analytic integration: per-branch sums over this rank's loci
    if (A.popKind == 2) {
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
        for (int b = warp; b < S; b += nw) {
            double q = 0.0, g = 0.0, r = 0.0;
            if (A.exact) {
                if (lane == 0)
                    for (int l = 0; l < L; l++) {
                        const size_t o = (size_t)l * S + b;
                        q += A.cur.brK[o]; g += A.cur.anG[o]; r += A.cur.anR[o];
                    }
            } else {
                for (int l = lane; l < L; l += 32) {
                    const size_t o = (size_t)l * S + b;
                    q += A.cur.brK[o]; g += A.cur.anG[o]; r += A.cur.anR[o];
                }
                for (int off = 16; off > 0; off >>= 1) {
                    q += __shfl_xor_sync(0xffffffffu, q, off);
                    g += __shfl_xor_sync(0xffffffffu, g, off);
                    r += __shfl_xor_sync(0xffffffffu, r, off);
                }
            }
            if (lane == 0) { A.red[3 + 3 * b] = q; A.red[3 + 3 * b + 1] = g; A.red[3 + 3 * b + 2] = r; }
        }
    } else {
        for (int i = tid; i < 3 * S; i += nt) A.red[3 + i] = 0.0;
    }
}

__device__ __forceinline__ void finish_body(const ReduceArgs &A)
{
    const int tid = threadIdx.x, nt = blockDim.x, L = A.nLoci, S = A.S;
    double *perBranch = A.out + 3 + 2 * L;
    double coal = A.red[0];
    const double lik = A.red[1];
    const bool incompatible = A.red[2] > 0.0;
    if (A.popKind == 2) {
        // one warp per species branch (no block barriers on this latency path); the per-branch terms are then added one by
        // one in node-number order (MultispeciesCoalescent.java:194)
        const double alpha = A.cur.hyper[0], beta = A.cur.hyper[1];
        const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
        for (int b = warp; b < S; b += nw) {
            const long long q = (long long)(A.red[3 + 3 * b] + 0.5);
            const double g = A.red[3 + 3 * b + 1], r = A.red[3 + 3 * b + 2];
            double lgr = 0.0;   // logGammaRatio, :227-230
            if (A.exact) {
                if (lane == 0) for (long long i = 0; i < q; i++) lgr += log(alpha + (double)i);   // the reference's order
            } else {
                for (long long i = lane; i < q; i += 32) lgr += log(alpha + (double)i);
                for (int off = 16; off > 0; off >>= 1) lgr += __shfl_xor_sync(0xffffffffu, lgr, off);
            }
            const double branchLogR = -r;
            // explicit roundings: this body is also inlined into k_treewalk, whose translation unit allows FMA contraction
            const double logP = __dadd_rn(__dsub_rn(__dadd_rn(branchLogR, __dmul_rn(alpha, log(beta))),
                                                    __dmul_rn(__dadd_rn(alpha, (double)q), log(__dadd_rn(beta, g)))), lgr);   // :232
            if (lane == 0) { perBranch[b] = logP; if (A.hostOut) A.hostOut[3 + 2 * L + b] = logP; }
        }
        __syncthreads();
        if (tid == 0) for (int b = 0; b < S; b++) coal += perBranch[b];
    } else {
        for (int b = tid; b < S; b += nt) { perBranch[b] = 0.0; if (A.hostOut) A.hostOut[3 + 2 * L + b] = 0.0; }
    }
    if (incompatible) coal = -INFINITY;   // MultispeciesCoalescent.java:153
    if (tid == 0) {
        A.out[0] = coal; A.out[1] = lik; A.out[2] = coal + lik;
        if (A.hostOut) { A.hostOut[0] = coal; A.hostOut[1] = lik; A.hostOut[2] = coal + lik; }
    }
}


// The result vector lives twice: in device memory (`out`) and in host-mapped memory (`hostOut`, zero-copy download: the
// host only has to synchronise the stream).  Both copies are written where the values are produced, so nothing is left to do
// here but make sure the block's writes are ordered before the kernel ends.
__device__ __forceinline__ void publish_body(const ReduceArgs &A)
{
    (void)A;
    __threadfence_system();
}

}  // namespace sb2
