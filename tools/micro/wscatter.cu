// HBM micro-benchmark: write bandwidth when every warp streams 256-bit stores into SCATTERED contiguous segments (what
// k_treewalk's output looks like: one (node, category) row of a 64-pattern tile = 2 KB per warp per op, rows far apart), with
// and without the int16 exponent rows next to them, aligned and misaligned.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/wscatter.bin tools/micro/wscatter.cu
#include <cstdio>
#include <cuda_runtime.h>

// each warp writes segsPerWarp segments of segBytes; segment index = (warpGlobal * A + i * B) % nSeg (scattered when B is large)
// shift: the segment starts `shift` doubles off its 128-byte line.  expMode 1: a 64-byte exponent row (int16 per lane) per 1 KB
// of partials, expShift int16 off its sector; expMode 2: one 128-byte exponent row (two int16 per lane) per 2 KB
__global__ void k_scatter(double *p, size_t nSeg, int segBytes, int segsPerWarp, size_t A, size_t B, int shift, int expMode, int expShift, short *e)
{
    const int lane = threadIdx.x & 31;
    const size_t wg = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    for (int i = 0; i < segsPerWarp; i++) {
        const size_t seg = (wg * A + (size_t)i * B) % nSeg;
        double *q = p + seg * (size_t)(segBytes / 8) + shift;
        for (int off = lane * 32; off < segBytes; off += 1024) {
            asm volatile("st.global.v4.f64 [%0], {%1,%1,%1,%1};" ::"l"(q + off / 8), "d"(1.0) : "memory");
            if (expMode == 1) e[seg * (size_t)(segBytes / 32) + off / 32 - lane + lane + expShift] = (short)i;
        }
        if (expMode == 2)
            for (int k = lane; k < segBytes / 64; k += 32) reinterpret_cast<int *>(e)[seg * (size_t)(segBytes / 64) + k] = i;
    }
}

int main()
{
    const size_t bytes = 2ull << 30;
    double *a;
    short *e;
    cudaMalloc(&a, bytes + 4096);
    cudaMalloc(&e, bytes / 16 + 4096);
    cudaMemset(a, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 128;
    const size_t warps = (size_t)blocks * threads / 32;
    struct Case { int seg, scattered, shift, expMode, expShift; const char *name; };
    const Case cases[] = {
        {1024, 0, 0, 0, 0, "sequential"}, {1024, 1, 0, 0, 0, "scattered"}, {1024, 1, 4, 0, 0, "scattered, rows 32 B off a line"},
        {1024, 1, 0, 1, 0, "scattered + 64 B exponent rows"}, {1024, 1, 4, 1, 1, "rows 32 B off + exponent rows 2 B off"},
        {2048, 0, 0, 0, 0, "sequential"}, {2048, 1, 0, 0, 0, "scattered"}, {2048, 1, 0, 1, 0, "scattered + 2 x 64 B exponent rows"},
        {2048, 1, 0, 2, 0, "scattered + 1 x 128 B exponent row"}, {2048, 1, 4, 1, 1, "rows 32 B off + exponent rows 2 B off"},
        {8192, 1, 0, 0, 0, "scattered"}, {8192, 1, 0, 2, 0, "scattered + 512 B exponent block"}, {32768, 1, 0, 0, 0, "scattered"},
    };
    for (const Case &c : cases) {
        const size_t nSeg = bytes / c.seg - 1;
        const int segsPerWarp = (int)(nSeg / warps);
        const size_t A = c.scattered ? 7919 : 1, B = c.scattered ? 1000003 : warps;
        float best = 1e9f;
        for (int it = 0; it < 5; it++) {
            cudaEventRecord(e0);
            k_scatter<<<blocks, threads>>>(a, nSeg, c.seg, segsPerWarp, A, B, c.shift, c.expMode, c.expShift, e);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        const double part = (double)segsPerWarp * warps * c.seg, all = part * (c.expMode ? 1.0625 : 1.0);
        printf("seg %6d B  %-42s %7.3f ms  partials %7.1f GB/s  with exponents %7.1f GB/s\n", c.seg, c.name, best, part / best / 1e6, all / best / 1e6);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
