// k_reduce_local / k_finish: per-locus and cross-locus sums.
//
//  * per-locus log-likelihood = sum over pattern tiles (TreeLikelihood.calcLogP: sum_p patternLogL[p]*weight[p]);
//  * `likelihood` and `speciescoalescent` CompoundDistribution sums over loci;
//  * analytic population-size integration: per species branch the sums over ALL loci of
//    (q = sum k_j, gamma = sum gamma_j/ploidy_j, logR = sum k_j log ploidy_j) -- MultispeciesCoalescent.java:207-225 --
//    go into the reduce vector, which the host all-reduces across GPUs (NCCL) when loci are sharded;
//    k_finish then applies the closed form of MultispeciesCoalescent.java:227-232 per branch.
//
// Compiled with -fmad=false.  exact=1 sums strictly sequentially in the reference's order (patterns ascending,
// loci in list order, branches by node number); exact=0 uses fixed-shape trees (deterministic, but not that order).
#include <math.h>

#include "sb2_device.cuh"

namespace sb2 {

constexpr int RED_THREADS = 256;

// deterministic block sum: fixed assignment of elements to threads, fixed-shape tree over threads
__device__ double block_sum(double v, double *scratch)
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

__device__ __forceinline__ void reduce_local_body(const ReduceArgs &A)
{
    __shared__ double scratch[32];
    const int tid = threadIdx.x, nt = blockDim.x, L = A.nLoci, S = A.S;
    double *perLocusCoal = A.out + 3, *perLocusLik = A.out + 3 + L;
    double likPart = 0.0, coalPart = 0.0, incPart = 0.0;
    for (int l = tid; l < L; l += nt) {
        double lik;
        if (A.useTiles && A.nOps[l] > 0) {
            const LocusDesc d = A.loci[l];
            lik = 0.0;
            if (A.exact) {
                for (int p = 0; p < d.nPat; p++) lik += A.patLogL[d.patBase + p] * A.weights[d.patBase + p];   // calcLogP order
            } else {
                for (int t = 0; t < d.nTiles; t++) lik += A.tileSum[d.tileBase + t];
            }
            A.cur.logL[l] = lik;
        } else {
            lik = A.cur.logL[l];   // clean locus: cached value (CompoundDistribution uses getCurrentLogP)
        }
        const double coal = A.cur.coal[l];
        perLocusLik[l] = lik;
        perLocusCoal[l] = coal;
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
    } else {
        likSum = block_sum(likPart, scratch);
        coalSum = block_sum(coalPart, scratch);
    }
    const double inc = block_sum(incPart, scratch);
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
    __shared__ double scratch[32];
    const int tid = threadIdx.x, nt = blockDim.x, L = A.nLoci, S = A.S;
    double *perBranch = A.out + 3 + 2 * L;
    double coal = A.red[0];
    const double lik = A.red[1];
    const bool incompatible = A.red[2] > 0.0;
    if (A.popKind == 2) {
        const double alpha = A.cur.hyper[0], beta = A.cur.hyper[1];
        for (int b = 0; b < S; b++) {
            const long long q = (long long)(A.red[3 + 3 * b] + 0.5);
            const double g = A.red[3 + 3 * b + 1], r = A.red[3 + 3 * b + 2];
            double part = 0.0;
            for (long long i = tid; i < q; i += nt) part += log(alpha + (double)i);   // logGammaRatio, :227-230
            const double lgr = block_sum(part, scratch);
            const double branchLogR = -r;
            const double logP = branchLogR + (alpha * log(beta)) - ((alpha + (double)q) * log(beta + g)) + lgr;   // :232
            if (tid == 0) perBranch[b] = logP;
            coal += logP;   // MultispeciesCoalescent.java:194: added one by one in node-number order
        }
    } else {
        for (int b = tid; b < S; b += nt) perBranch[b] = 0.0;
    }
    if (incompatible) coal = -INFINITY;   // MultispeciesCoalescent.java:153
    if (tid == 0) { A.out[0] = coal; A.out[1] = lik; A.out[2] = coal + lik; }
}

__global__ void __launch_bounds__(RED_THREADS) k_reduce_local(ReduceArgs A) { reduce_local_body(A); }
__global__ void __launch_bounds__(RED_THREADS) k_finish(ReduceArgs A) { finish_body(A); }
// single-GPU evaluation: no all-reduce between the local sums and the closed forms -> one launch
__global__ void __launch_bounds__(RED_THREADS) k_reduce_finish(ReduceArgs A)
{
    reduce_local_body(A);
    __threadfence_block();
    __syncthreads();
    finish_body(A);
}

void launch_reduce_finish(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    k_reduce_finish<<<1, RED_THREADS, 0, st>>>(a);
    ++*launches;
}

void launch_reduce_local(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    k_reduce_local<<<1, RED_THREADS, 0, st>>>(a);
    ++*launches;
}

void launch_finish(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    k_finish<<<1, RED_THREADS, 0, st>>>(a);
    ++*launches;
}

}  // namespace sb2
