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

#include "reduce_body.cuh"

namespace sb2 {

constexpr int RED_THREADS = 256;

__global__ void __launch_bounds__(RED_THREADS) k_reduce_local(ReduceArgs A) { cudaGridDependencySynchronize(); reduce_local_body(A); }
__global__ void __launch_bounds__(RED_THREADS) k_finish(ReduceArgs A) { finish_body(A); publish_body(A); }
// single-GPU evaluation: no all-reduce between the local sums and the closed forms -> one launch
__global__ void __launch_bounds__(RED_THREADS) k_reduce_finish(ReduceArgs A)
{
    int tb = -1, ntl = 0;
    if ((int)threadIdx.x < A.nLoci) { tb = A.loci[threadIdx.x].tileBase; ntl = A.loci[threadIdx.x].nTiles; }   // static: fetched while the pruning kernel still runs
    cudaGridDependencySynchronize();   // launched with programmatic serialization: wait for the pruning kernel's results
    reduce_local_body(A, tb, ntl);
    __threadfence_block();
    __syncthreads();
    finish_body(A);
    publish_body(A);
}

// k_reduce_xgpu: local sums + ALL-REDUCE OVER PEER MEMORY + closed forms, one kernel.
// Every rank stores its reduce vector straight into slot [parity][rank] of every peer's communication buffer over
// NVLink, each double as two 8-byte words that carry the evaluation's sequence tag in their upper half (flag-in-data, the
// idea of NCCL's LL protocol): the receiver polls the words themselves, so there is no fence and no separate flag on the
// critical path -- one one-way NVLink trip.  All ranks sum the slots in rank order and obtain bit-identical totals.
// Buffer halves alternate with the sequence parity: a rank can be at most one evaluation ahead of a peer (it needs the
// peer's words of evaluation s+1 before it can finish s+1 and start s+2), so a half is never overwritten while a slow peer
// still reads it.  A rank that never shows up cannot hang the GPU: the wait gives up after ~2 s and poisons the result
// with NaN.
__global__ void __launch_bounds__(RED_THREADS) k_reduce_xgpu(ReduceArgs A, CommArgs Cm)
{
    int tb = -1, ntl = 0;
    if ((int)threadIdx.x < A.nLoci) { tb = A.loci[threadIdx.x].tileBase; ntl = A.loci[threadIdx.x].nTiles; }
    cudaGridDependencySynchronize();
    reduce_local_body(A, tb, ntl);
    __threadfence_block();
    __syncthreads();
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
    for (int i = tid; i < n; i += nt) {
        double sum = 0.0;
        const long long t0 = clock64();
        for (int r = 0; r < nR; r++) {   // rank order: every rank adds the same numbers in the same order
            const volatile unsigned long long *src = Cm.peerSlots[Cm.rank] + (((size_t)half * COMM_MAX_RANKS + r) * n + i) * 2;
            unsigned long long w0, w1;
            for (;;) {
                w0 = src[0]; w1 = src[1];
                if ((w0 & 0xffffffff00000000ull) == tag && (w1 & 0xffffffff00000000ull) == tag) break;
                if (clock64() - t0 > 4000000000ll) { s_poison = 1; w0 = w1 = 0; break; }   // ~2 s at 2 GHz: a peer is gone
            }
            sum += __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        }
        A.red[i] = sum;
    }
    __threadfence_block();
    __syncthreads();
    if (tid == 0 && A.hostOut) {   // time spent waiting for the peers' words (rank skew + one NVLink trip), for sb2_exchange_wait_us
        unsigned long long waitT1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(waitT1));
        A.hostOut[3 + 2 * A.nLoci + A.S] = (double)(waitT1 - waitT0) * 1e-3;
    }
    if (s_poison) { for (int i = tid; i < n; i += nt) A.red[i] = NAN; __threadfence_block(); __syncthreads(); }
    finish_body(A);
    publish_body(A);
}

template <typename K>
static cudaError_t launch_pdl(K kernel, const ReduceArgs &a, cudaStream_t st)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(RED_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

cudaError_t launch_reduce_finish(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    ++*launches;
    return launch_pdl(k_reduce_finish, a, st);
}

cudaError_t launch_reduce_xgpu(const ReduceArgs &a, const CommArgs &c, cudaStream_t st, int64_t *launches)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(RED_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ++*launches;
    return cudaLaunchKernelEx(&cfg, k_reduce_xgpu, a, c);
}

cudaError_t launch_reduce_local(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    ++*launches;
    return launch_pdl(k_reduce_local, a, st);
}

cudaError_t launch_finish(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    k_finish<<<1, RED_THREADS, 0, st>>>(a);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace sb2
