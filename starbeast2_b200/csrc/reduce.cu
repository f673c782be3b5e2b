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
    cudaGridDependencySynchronize();   // launched with programmatic serialization: wait for the pruning kernel's results
    reduce_local_body(A);
    __threadfence_block();
    __syncthreads();
    finish_body(A);
    publish_body(A);
}

template <typename K>
static void launch_pdl(K kernel, const ReduceArgs &a, cudaStream_t st)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(RED_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, a);
}

void launch_reduce_finish(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    launch_pdl(k_reduce_finish, a, st);
    ++*launches;
}

void launch_reduce_local(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    launch_pdl(k_reduce_local, a, st);
    ++*launches;
}

void launch_finish(const ReduceArgs &a, cudaStream_t st, int64_t *launches)
{
    k_finish<<<1, RED_THREADS, 0, st>>>(a);
    ++*launches;
}

}  // namespace sb2
