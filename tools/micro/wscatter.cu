// HBM micro-benchmark: write bandwidth when every warp streams 256-bit stores into SCATTERED contiguous segments (what
// k_treewalk's output looks like: one (node, category) row of a 64-pattern tile = 2 KB per warp per op, rows far apart).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/wscatter.bin tools/micro/wscatter.cu
#include <cstdio>
#include <cuda_runtime.h>

// each warp writes `segs` segments of segBytes; segment index = (warpGlobal * A + i * B) % nSeg  (scattered when B is large)
__global__ void k_scatter(double *p, size_t nSeg, int segBytes, int segsPerWarp, size_t A, size_t B, int misalign, short *e)
{
    const int lane = threadIdx.x & 31;
    const size_t wg = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    for (int i = 0; i < segsPerWarp; i++) {
        const size_t seg = (wg * A + (size_t)i * B) % nSeg;
        double *q = p + seg * (size_t)(segBytes / 8) + misalign * 4;   // misalign: segments start 32 B off a 128-byte line
        if (e) e[seg * 32 + lane + (misalign ? 1 : 0)] = (short)i;      // the exponent row: 64 B per 1 KB of partials (2 B off a sector when misaligned)
        for (int off = lane * 32; off < segBytes; off += 1024)
            asm volatile("st.global.v4.f64 [%0], {%1,%1,%1,%1};" ::"l"(q + off / 8), "d"(1.0) : "memory");
    }
}

int main()
{
    const size_t bytes = 2ull << 30;
    double *a;
    cudaMalloc(&a, bytes);
    cudaMemset(a, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 128;
    const size_t warps = (size_t)blocks * threads / 32;
    short *e;
    cudaMalloc(&e, bytes / 16 + 1024);
    for (int segBytes : {1024, 2048, 4096, 8192, 32768, 262144}) {
        const size_t nSeg = bytes / segBytes - 1;
        const int segsPerWarp = (int)(nSeg / warps);
        for (int mode = 0; mode < (segBytes == 1024 ? 6 : 2); mode++) {
            // mode 0: sequential (warp w writes segments w, w + warps, ...: neighbours in time are neighbours in memory)
            // mode 1: scattered (each warp jumps by a large odd stride between its segments)
            const size_t A = mode == 0 ? 1 : 7919, B = mode == 0 ? warps : 1000003;
            const int mis = (mode == 2 || mode == 4) ? 1 : 0;      // modes 2..5 (1 KB segments only): misaligned; + exponent row; both; aligned + exponent row
            short *ex = mode >= 3 ? e : nullptr;
            float best = 1e9f;
            for (int it = 0; it < 5; it++) {
                cudaEventRecord(e0);
                k_scatter<<<blocks, threads>>>(a, nSeg, segBytes, segsPerWarp, A, B, mis, ex);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            const double moved = (double)segsPerWarp * warps * segBytes;
            printf("seg %7d B %-18s %8.3f ms  %8.1f GB/s\n", segBytes, mode == 0 ? "sequential" : mode == 1 ? "scattered" : mode == 2 ? "scat+mis32" : mode == 3 ? "scat+exp" : mode == 4 ? "scat+mis+exp(mis)" : "scat+exp(aligned)", best, moved / best / 1e6);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
