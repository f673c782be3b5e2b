// Alignment -> unique site patterns + weights on the device (BEAST.base Alignment.calcPatterns, the
// step that produces the tip states / pattern weights this engine is fed with; named by the reference's XML, e.g.
// tutorial/CanisPhylogeny/Canis-Phylogeny.xml:3-23 <data><sequence .../></data>).
//
// calcPatterns sorts the site columns lexicographically (taxon 0 is the most significant position), merges equal columns and
// counts them.  Here: (1) every column is packed into big-endian 64-bit words, 8 taxa per word, so that comparing words in
// order IS the lexicographic comparison; (2) one CTA bitonic-sorts the site indices in shared memory (ties broken by site
// number: deterministic), flags the first column of every run, scans the flags and writes patterns, weights and the
// site -> pattern map.  Byte / integer work: results are bit-identical to numpy.unique(axis=1) (tests/test_gpu_compress.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>
#include <algorithm>

#include "../../include/sb2gpu.h"

namespace {

constexpr int SORT_THREADS = 1024;
constexpr int MAX_SITES = 16384;   // two int arrays of the next power of two must fit 227 KB of shared memory

__global__ void k_pack_columns(const uint8_t *seq, int nTips, int nSites, int W, unsigned long long *keys)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (site, word) with site fastest: coalesced reads of seq rows
    if (i >= nSites * W) return;
    const int w = i / nSites, site = i - w * nSites;
    unsigned long long k = 0;
    for (int b = 0; b < 8; b++) {
        const int t = w * 8 + b;
        const unsigned long long v = t < nTips ? seq[(size_t)t * nSites + site] : 0;
        k = (k << 8) | v;
    }
    keys[(size_t)site * W + w] = k;
}

// -1 / 0 / +1; padding indices (>= nSites) sort last
__device__ __forceinline__ int compare_columns(const unsigned long long *keys, int W, int nSites, int a, int b)
{
    const bool pa = a >= nSites, pb = b >= nSites;
    if (pa || pb) return pa == pb ? (a < b ? -1 : (a > b ? 1 : 0)) : (pa ? 1 : -1);
    const unsigned long long *ka = keys + (size_t)a * W, *kb = keys + (size_t)b * W;
    for (int w = 0; w < W; w++) {
        const unsigned long long x = ka[w], y = kb[w];
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_unique(const uint8_t *seq, const unsigned long long *keys, int nTips, int nSites,
                                                                int W, int P2, uint8_t *patterns /*[tip][nSites]*/, int32_t *weights,
                                                                int32_t *siteToPattern, int32_t *nPatOut)
{
    extern __shared__ int sm[];
    int *idx = sm;            // [P2] site indices being sorted
    int *scan = sm + P2;      // [P2] run flags -> inclusive scan
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < P2; i += nt) idx[i] = i;
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P2; i += nt) {
                const int p = i ^ j;
                if (p > i) {
                    const int a = idx[i], b = idx[p];
                    int c = compare_columns(keys, W, nSites, a, b);
                    if (c == 0) c = a < b ? -1 : 1;                       // equal columns: by site number
                    const bool up = (i & k) == 0;
                    if ((c > 0) == up) { idx[i] = b; idx[p] = a; }
                }
            }
            __syncthreads();
        }
    }
    // first column of every run of equal columns
    for (int i = tid; i < P2; i += nt)
        scan[i] = (i < nSites) && (i == 0 || compare_columns(keys, W, nSites, idx[i - 1], idx[i]) != 0);
    __syncthreads();
    // inclusive scan (Hillis-Steele over shared memory; P2 <= 16384, log2 steps)
    for (int off = 1; off < P2; off <<= 1) {
        int v[MAX_SITES / SORT_THREADS];
        int n = 0;
        for (int i = tid; i < P2; i += nt) v[n++] = scan[i] + (i >= off ? scan[i - off] : 0);
        __syncthreads();
        n = 0;
        for (int i = tid; i < P2; i += nt) scan[i] = v[n++];
        __syncthreads();
    }
    const int nPat = scan[nSites - 1];
    if (tid == 0) *nPatOut = nPat;
    for (int i = tid; i < nSites; i += nt) {
        const int p = scan[i] - 1;
        siteToPattern[idx[i]] = p;
        const bool first = i == 0 || scan[i - 1] != scan[i];
        if (first) {
            int j = i + 1;                                   // run length = weight (runs are short on real data; bounded by nSites)
            while (j < nSites && scan[j] == scan[i]) j++;
            weights[p] = j - i;
            for (int t = 0; t < nTips; t++) patterns[(size_t)t * nSites + p] = seq[(size_t)t * nSites + idx[i]];
        }
    }
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char *sb2_compress_last_error(void) { return g_err.c_str(); }

// One block of at most MAX_SITES site columns [s0, s0 + nSites) of an alignment whose rows have nSitesTotal entries.
// patterns: compact rows [tip][*nPatterns] in a buffer whose rows hold patRowCap entries.
static int compress_block(int32_t device, int32_t nTips, int32_t nSitesTotal, int32_t s0, int32_t nSites, const uint8_t *sequences,
                          uint8_t *patterns, int32_t *weights, int32_t *siteToPattern, int32_t *nPatterns)
{
    auto fail = [&](int code, const std::string &m) { g_err = m; return code; };
    const int W = (nTips + 7) / 8;
    int P2 = 2;
    while (P2 < nSites) P2 <<= 1;
    const size_t seqBytes = (size_t)nTips * nSites;
    // one grow-only scratch arena per host thread and device: alignments are compressed one after the other at start-up,
    // and cudaMalloc / cudaFree per call would cost more than the kernels
    static thread_local struct { int device = -1; uint8_t *base = nullptr; size_t bytes = 0; } arena;
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t oSeq = 0, oPat = oSeq + up(seqBytes), oKeys = oPat + up(seqBytes), oW = oKeys + up(sizeof(unsigned long long) * (size_t)nSites * W),
                 oMap = oW + up(4 * (size_t)nSites), oN = oMap + up(4 * (size_t)nSites), need = oN + 256;
    cudaError_t st = cudaSuccess;
    if (arena.device != device || arena.bytes < need) {
        if (arena.base) { cudaSetDevice(arena.device); cudaFree(arena.base); cudaSetDevice(device); }
        arena.base = nullptr; arena.bytes = 0; arena.device = device;
        st = cudaMalloc((void **)&arena.base, need);
        if (st == cudaSuccess) arena.bytes = need;
    }
    if (st != cudaSuccess) return fail(SB2_ERR_NOMEM, cudaGetErrorString(st));
    uint8_t *dSeq = arena.base + oSeq, *dPat = arena.base + oPat;
    unsigned long long *dKeys = reinterpret_cast<unsigned long long *>(arena.base + oKeys);
    int32_t *dW = reinterpret_cast<int32_t *>(arena.base + oW), *dMap = reinterpret_cast<int32_t *>(arena.base + oMap),
            *dN = reinterpret_cast<int32_t *>(arena.base + oN);
    int rc = SB2_OK;
    // the block's columns of every row: host rows have nSitesTotal entries, device rows nSites
    if (st == cudaSuccess)
        st = cudaMemcpy2D(dSeq, (size_t)nSites, sequences + s0, (size_t)nSitesTotal, (size_t)nSites, (size_t)nTips, cudaMemcpyHostToDevice);
    if (st == cudaSuccess) {
        const int total = nSites * W;
        k_pack_columns<<<(total + 255) / 256, 256>>>(dSeq, nTips, nSites, W, dKeys);
        const size_t smem = 8 * (size_t)P2;
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_sort_unique, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_sort_unique<<<1, SORT_THREADS, smem>>>(dSeq, dKeys, nTips, nSites, W, P2, dPat, dW, dMap, dN);
        st = cudaGetLastError();
    }
    int32_t nPat = 0;
    if (st == cudaSuccess) st = cudaMemcpy(&nPat, dN, 4, cudaMemcpyDeviceToHost);
    if (st == cudaSuccess) {
        *nPatterns = nPat;
        st = cudaMemcpy(weights, dW, 4 * (size_t)nPat, cudaMemcpyDeviceToHost);
        if (st == cudaSuccess && siteToPattern) st = cudaMemcpy(siteToPattern, dMap, 4 * (size_t)nSites, cudaMemcpyDeviceToHost);
        // device rows have stride nSites; the caller's buffer is [tip][nPatterns], compact
        if (st == cudaSuccess)
            st = cudaMemcpy2D(patterns, (size_t)nPat, dPat, (size_t)nSites, (size_t)nPat, (size_t)nTips, cudaMemcpyDeviceToHost);
    }
    if (st != cudaSuccess) rc = fail(SB2_ERR_CUDA, cudaGetErrorString(st));
    return rc;
}

int sb2_compress_alignment(int32_t device, int32_t nTips, int32_t nSites, const uint8_t *sequences, uint8_t *patterns,
                           int32_t *weights, int32_t *siteToPattern, int32_t *nPatterns)
{
    auto fail = [&](int code, const std::string &m) { g_err = m; return code; };
    if (nTips < 1 || nSites < 1 || !sequences || !patterns || !weights || !nPatterns) return fail(SB2_ERR_ARG, "bad argument");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) return fail(SB2_ERR_CUDA, "no CUDA device; libsb2gpu has no CPU fallback");
    if (device < 0 || device >= nDev) return fail(SB2_ERR_ARG, "device out of range");
    cudaSetDevice(device);
    if (nSites <= MAX_SITES) return compress_block(device, nTips, nSites, 0, nSites, sequences, patterns, weights, siteToPattern, nPatterns);

    // Long alignments (phylogenomic concatenations): blocks of MAX_SITES columns are sorted and made unique on the device, the
    // blocks' sorted pattern lists are then merged here (the unique patterns of a block are few next to its sites, and this
    // is set-up time).  Same result as one pass: sorted unique columns, summed multiplicities, site -> pattern map.
    const int nBlocks = (nSites + MAX_SITES - 1) / MAX_SITES;
    std::vector<std::vector<uint8_t>> bPat(nBlocks);
    std::vector<std::vector<int32_t>> bW(nBlocks), bMap(nBlocks), bGlobal(nBlocks);
    std::vector<int32_t> bN(nBlocks, 0);
    for (int b = 0; b < nBlocks; b++) {
        const int s0 = b * MAX_SITES, n = std::min(MAX_SITES, nSites - s0);
        bPat[b].resize((size_t)nTips * n); bW[b].resize(n); bMap[b].resize(n);
        const int rc = compress_block(device, nTips, nSites, s0, n, sequences, bPat[b].data(), bW[b].data(), bMap[b].data(), &bN[b]);
        if (rc != SB2_OK) return rc;
        bGlobal[b].resize(bN[b]);
    }
    // column i of block b: bPat[b][t * bN[b] + i]; lexicographic order with taxon 0 most significant
    auto cmp = [&](int b1, int i1, int b2, int i2) {
        for (int t = 0; t < nTips; t++) {
            const uint8_t x = bPat[b1][(size_t)t * bN[b1] + i1], y = bPat[b2][(size_t)t * bN[b2] + i2];
            if (x != y) return x < y ? -1 : 1;
        }
        return 0;
    };
    std::vector<int32_t> head(nBlocks, 0);
    std::vector<std::pair<int, int>> order;   // (block, index) of the first occurrence of every global pattern
    std::vector<int32_t> gW;
    for (;;) {
        int best = -1;
        for (int b = 0; b < nBlocks; b++)
            if (head[b] < bN[b] && (best < 0 || cmp(b, head[b], best, head[best]) < 0)) best = b;
        if (best < 0) break;
        const int bi = head[best];
        const int32_t g = (int32_t)order.size();
        order.push_back({best, bi});
        int32_t w = 0;
        for (int b = 0; b < nBlocks; b++)
            if (head[b] < bN[b] && cmp(b, head[b], best, bi) == 0) { w += bW[b][head[b]]; bGlobal[b][head[b]] = g; if (b != best) head[b]++; }
        head[best]++;
        gW.push_back(w);
    }
    const int32_t nPat = (int32_t)order.size();
    *nPatterns = nPat;
    for (int32_t g = 0; g < nPat; g++) {
        weights[g] = gW[g];
        const int b = order[g].first, i = order[g].second;
        for (int t = 0; t < nTips; t++) patterns[(size_t)t * nPat + g] = bPat[b][(size_t)t * bN[b] + i];
    }
    if (siteToPattern)
        for (int b = 0; b < nBlocks; b++) {
            const int s0 = b * MAX_SITES, n = std::min(MAX_SITES, nSites - s0);
            for (int j = 0; j < n; j++) siteToPattern[s0 + j] = bGlobal[b][bMap[b][j]];
        }
    return SB2_OK;
}

}  // extern "C"
