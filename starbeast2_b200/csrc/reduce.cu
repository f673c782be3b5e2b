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
    if ((int)threadIdx.x < A.nLoci) { tb = A.loci[threadIdx.x].chunkBase; ntl = A.loci[threadIdx.x].nChunks; }   // static: fetched while the pruning kernel still runs
    cudaGridDependencySynchronize();   // launched with programmatic serialization: wait for the pruning kernel's results
    double sums3[3];
    reduce_local_body(A, tb, ntl, sums3);
    if (A.popKind == 2) { __threadfence_block(); __syncthreads(); }   // the per-branch sums travel through the reduce vector
    finish_body(A, sums3);
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
    if ((int)threadIdx.x < A.nLoci) { tb = A.loci[threadIdx.x].chunkBase; ntl = A.loci[threadIdx.x].nChunks; }
    cudaGridDependencySynchronize();
    reduce_local_body(A, tb, ntl);
    __threadfence_block();
    __syncthreads();
    xgpu_exchange_body(A, Cm);
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
