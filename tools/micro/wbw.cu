// HBM micro-benchmark: pure-write, pure-read and copy bandwidth with 256-bit accesses (what k_treewalk's stores look like).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/wbw.bin tools/micro/wbw.cu && tools/micro/wbw.bin
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_fill(double *p, size_t n4, double v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.v4.f64 [%0], {%1,%1,%1,%1};" ::"l"(p + i * 4), "d"(v) : "memory");
}
__global__ void k_read(const double *p, size_t n4, double *out)
{
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        double a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p + i * 4));
        s += a + b + c + d;
    }
    if (s == 12345.678) *out = s;
}
__global__ void k_copy(const double *p, double *q, size_t n4)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        double a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p + i * 4));
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(q + i * 4), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    }
}

int main()
{
    const size_t bytes = 2ull << 30;   // 2 GiB per buffer
    double *a, *b;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
    const size_t n4 = bytes / 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        for (int mode = 0; mode < 4; mode++) {
            float best = 1e9f;
            for (int it = 0; it < 6; it++) {
                cudaEventRecord(e0);
                if (mode == 0) k_fill<<<blocks, 256>>>(a, n4, 1.0);
                else if (mode == 1) k_read<<<blocks, 256>>>(a, n4, b);
                else if (mode == 2) k_copy<<<blocks, 256>>>(a, b, n4);
                else cudaMemsetAsync(a, 0, bytes);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            const double moved = (mode == 2 ? 2.0 : 1.0) * bytes;
            printf("blocks %5d %-7s %8.3f ms  %8.1f GB/s\n", blocks, mode == 0 ? "fill" : mode == 1 ? "read" : mode == 2 ? "copy" : "memset", best, moved / best / 1e6);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
