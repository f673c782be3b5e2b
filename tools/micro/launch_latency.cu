// Host-timed launch + completion latency of near-empty kernels on B200, varying one launch property at a time:
// cluster dimension, kernel-argument bytes, dynamic shared memory, completion by stream sync vs polled mapped word.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_latency launch_latency.cu
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

struct Small { unsigned long long *flag; unsigned long long seq; };
struct Big { unsigned long long *flag; unsigned long long seq; unsigned char pad[4400]; };

template <typename T>
__global__ void __launch_bounds__(256, 2) k_flag(const __grid_constant__ T a)
{
    extern __shared__ unsigned char sm[];
    if (threadIdx.x == 0 && blockIdx.x == 0) { *(volatile unsigned long long *)a.flag = a.seq; }
}

template <typename T>
double run(const char *name, int cluster, size_t smem, bool poll, int grid)
{
    unsigned long long *h, *d;
    cudaHostAlloc(&h, 64, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&d, h, 0);
    *h = 0;
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaFuncSetAttribute(k_flag<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cluster > 8) cudaFuncSetAttribute(k_flag<T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    T a;
    memset(&a, 0, sizeof a);
    a.flag = d;
    const int N = 3000, W = 300;
    double total = 0;
    for (int i = 0; i < N + W; i++) {
        a.seq = i + 1;
        auto t0 = std::chrono::steady_clock::now();
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = cluster > 0 ? 1 : 0;
        cudaLaunchKernelEx(&cfg, k_flag<T>, a);
        if (poll) { while (*(volatile unsigned long long *)h != a.seq) { } }
        else cudaStreamSynchronize(st);
        auto t1 = std::chrono::steady_clock::now();
        if (i >= W) total += std::chrono::duration<double, std::micro>(t1 - t0).count();
        if (poll && (i & 63) == 63) cudaStreamSynchronize(st);
    }
    cudaStreamSynchronize(st);
    printf("%-58s %6.2f us\n", name, total / N);
    cudaStreamDestroy(st);
    cudaFreeHost(h);
    return total / N;
}

int main()
{
    run<Small>("plain, 1 CTA, sync", 0, 0, false, 1);
    run<Small>("plain, 1 CTA, poll", 0, 0, true, 1);
    run<Small>("cluster 1, 1 CTA, poll", 1, 0, true, 1);
    run<Small>("cluster 3, 3 CTAs, poll", 3, 0, true, 3);
    run<Small>("cluster 3, 3 CTAs, 56 KB smem, poll", 3, 56 * 1024, true, 3);
    run<Small>("no cluster, 3 CTAs, 56 KB smem, poll", 0, 56 * 1024, true, 3);
    run<Big>("no cluster, 3 CTAs, 4.4 KB args, poll", 0, 0, true, 3);
    run<Big>("cluster 3, 3 CTAs, 56 KB smem, 4.4 KB args, poll", 3, 56 * 1024, true, 3);
    run<Big>("cluster 3, 3 CTAs, 56 KB smem, 4.4 KB args, sync", 3, 56 * 1024, false, 3);
    run<Big>("cluster 3, 150 CTAs, 56 KB smem, 4.4 KB args, poll", 3, 56 * 1024, true, 150);
    run<Small>("plain, 463 CTAs, poll", 0, 0, true, 463);
    return 0;
}
