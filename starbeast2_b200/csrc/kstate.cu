// k_kstate: TreeLikelihood.calculateLogP for the K-state, likelihood-only loci -- the morphological partitions that the FBD
// configurations score on the SPECIES tree (tutorial/CanisFBD/Canis-FBD.xml:964-1033: spec="TreeLikelihood"
// tree="@Tree.t:Species", StandardData nrOfStates 2..5, LewisMK, ascertained FilteredAlignment, strict clock).
//
// Replaces, for K states: BeerLikelihoodCore.calculate{StatesStates,StatesPartials,PartialsPartials}Pruning, scalePartials,
// integratePartials, calculateLogLikelihoods (BEAST.base, generic state count), GeneralSubstitutionModel.
// getTransitionProbabilities, Alignment.getAscertainmentCorrection and TreeLikelihood.calcLogP's ascertained branch.
//
// These data are tiny (tens of characters x tens of tips) and hang off the species tree, which changes in a minority of the
// MCMC steps; they are latency, not bandwidth.  One CTA per locus, everything from scratch whenever the locus is flagged:
// the tree and all transition matrices in shared memory, one thread per site pattern walking the whole tree in post-order
// for each rate category (partials in a small HBM scratch that only this thread touches), sums in the reference's order
// (this translation unit is compiled with -fmad=false), exact power-of-two rescaling at every node.  The result goes into the
// locus's cached log-likelihood in the state block, where the reductions, store() and restore() find it.
#include <math.h>

#include "sb2_device.cuh"

namespace sb2 {

constexpr int KS_THREADS = 256;   // 32 groups of 8 lanes: the usual tens of characters take one pass of the walk
constexpr int KS_SMEM_PATTERNS = 1024;   // pattern log-likelihoods kept in shared memory for the closing sum
constexpr int KS_SMEM_TIP_BYTES = 16 * 1024;   // tip states staged in shared memory up to this size (a tip op's state byte is on the walk's critical path)

constexpr size_t KS_SCRATCH_SMEM_MAX = 96 * 1024;   // per-locus partials + exponents up to this size live in shared memory

struct KsSmem {
    size_t gH, gP, gL, gR, order, stack, mats, scr, spl, tips, total;
};
// scratchBytes: size of the shared-memory partials scratch (the largest locus that fits KS_SCRATCH_SMEM_MAX, see
// kstate_scratch_bytes); the carve is the same for every K of one launch except for the two trailing regions
__host__ __device__ inline KsSmem ks_carve(int G, int C, int K, size_t scratchBytes)
{
    KsSmem m;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += (b + 15) & ~size_t(15); return r; };
    m.gH = take(8ull * G); m.gP = take(4ull * G); m.gL = take(4ull * G); m.gR = take(4ull * G);
    m.order = take(4ull * G); m.stack = take(4ull * G);
    m.spl = take(8ull * KS_SMEM_PATTERNS);   // pattern log-likelihoods for the closing sum
    m.tips = take(KS_SMEM_TIP_BYTES);        // the locus's tip states [tip][pattern] when they fit
    m.scr = take(scratchBytes);
    m.mats = take(8ull * G * C * K * K);
    m.total = o;
    return m;
}
// bytes of partials [internal][category][pattern][K] + int32 exponents [internal][category][pattern] of one locus
__host__ __device__ inline size_t ks_locus_scratch(int nInt, int C, int nPat, int K) { return (size_t)nInt * C * nPat * (8 * K + 4); }
size_t kstate_scratch_bytes(int nInt, int nCat, int nPat, int K) { return ks_locus_scratch(nInt, nCat, nPat, K); }
size_t kstate_smem_bytes(int maxNodes, int maxCat, int K, size_t scratchBytes) { return ks_carve(maxNodes, maxCat, K, scratchBytes).total; }

template <int K>
__device__ __forceinline__ void kstate_locus(const KStateArgs &A, int l, const LocusDesc &d, unsigned char *smem)
{
    constexpr int KK = K * K;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int G = d.nNodes, nT = d.nTips, nI = d.nInt, C = d.nCat, nPat = d.nPat, nb = d.nodeBase;
    const KsSmem m = ks_carve(A.maxNodes, A.maxCat, K, A.smemScratch);
    double *gH = reinterpret_cast<double *>(smem + m.gH);
    int *gP = reinterpret_cast<int *>(smem + m.gP), *gL = reinterpret_cast<int *>(smem + m.gL), *gR = reinterpret_cast<int *>(smem + m.gR);
    int *order = reinterpret_cast<int *>(smem + m.order), *stack = reinterpret_cast<int *>(smem + m.stack);
    double *mats = reinterpret_cast<double *>(smem + m.mats);   // [node][category][K*K]
    __shared__ LocusParams prm;
    __shared__ int s_root;
    // tip states: staged in shared memory when they fit (tens of tips x tens of characters) -- read from HBM inside the walk,
    // every tip op was a dependent L2 round trip (8 of the kernel's 38 us); fetched in the same wave of loads as the tree
    const uint8_t *gStates = A.tipStates + d.stateBase;
    uint8_t *sStates = smem + m.tips;
    const bool tipsInSmem = (size_t)nT * nPat <= KS_SMEM_TIP_BYTES;
    {
        if (tipsInSmem)
            for (int idx = tid; idx < nT * nPat; idx += nt) { const int t = idx / nPat, p = idx - t * nPat; sStates[idx] = gStates[(size_t)t * d.stateStride + p]; }
        constexpr int NP = (int)(sizeof(LocusParams) / 8);
        const double *src = reinterpret_cast<const double *>(&A.params[l]);
        double *dst = reinterpret_cast<double *>(&prm);
        for (int i = tid; i < NP; i += nt) dst[i] = src[i];
        for (int i = tid; i < G; i += nt) {   // the tree k_prepare has just made current
            gH[i] = A.cur.gHeight[nb + i]; gP[i] = A.cur.gParent[nb + i]; gL[i] = A.cur.gLeft[nb + i]; gR[i] = A.cur.gRight[nb + i];
        }
    }
    __syncthreads();
    for (int i = tid; i < G; i += nt) if (gP[i] < 0) s_root = i;
    __syncthreads();
    const int root = s_root;
    if (tid == 0) {   // post-order of the internal nodes (TreeLikelihood.traverse's order: left subtree, right subtree, node)
        int sp = 0, n = 0;
        // two-stack iterative post-order: visit node, push left then right -> reverse pre-order (node, right, left) reversed
        stack[sp++] = root;
        while (sp > 0) {
            const int v = stack[--sp];
            if (v < nT) continue;
            order[nI - 1 - n] = v; n++;
            stack[sp++] = gL[v]; stack[sp++] = gR[v];
        }
    }
    // GeneralSubstitutionModel.getTransitionProbabilities per (branch, category): |V exp(d L) V^-1|
    const double *Eval = prm.subst, *Evec = prm.subst + K, *Ievc = prm.subst + K + KK, *freqs = prm.subst + K + 2 * KK;
    // (warps 1.. : thread 0 is busy with the post-order meanwhile)
    for (int idx = tid - 32; idx < G * C; idx += nt - 32) {
        if (idx < 0) break;
        const int i = idx / C, c = idx - i * C;
        double *M = mats + (size_t)idx * KK;
        if (gP[i] < 0) { for (int q = 0; q < KK; q++) M[q] = 0.0; continue; }
        const double branchRate = (prm.clockKind == 2) ? A.explicitRates[nb + i] : prm.clockRate;   // StrictClockModel / explicit
        const double jointBranchRate = A.catRates[d.catBase + c] * branchRate;
        const double distance = (gH[gP[i]] - gH[i]) * jointBranchRate;
        double iexp[KK];
#pragma unroll
        for (int a = 0; a < K; a++) {
            const double temp = exp(distance * Eval[a]);
#pragma unroll
            for (int b = 0; b < K; b++) iexp[a * K + b] = Ievc[a * K + b] * temp;
        }
#pragma unroll
        for (int a = 0; a < K; a++)
#pragma unroll
            for (int b = 0; b < K; b++) {
                double temp = 0.0;
#pragma unroll
                for (int k = 0; k < K; k++) temp += Evec[a * K + k] * iexp[k * K + b];
                M[a * K + b] = fabs(temp);
            }
    }
    __syncthreads();

    // tip states: staged in shared memory when they fit (tens of tips x tens of characters) -- read from HBM inside the walk,
    // every tip op was a dependent L2 round trip
    // partials [internal][category][pattern][K] and exponents [internal][category][pattern]: in shared memory when the locus
    // fits (the usual case: tens of characters x tens of tips), else in the HBM scratch; only this thread touches its pattern's
    const size_t nElem = (size_t)nI * C * nPat;
    const bool inSmem = ks_locus_scratch(nI, C, nPat, K) <= A.smemScratch;
    const double *props = A.catProps + d.catBase;
    double *spl = reinterpret_cast<double *>(smem + m.spl);
    // Eight lanes per pattern, lane i < K computes output state i: an op's K x K contraction is K chains of K terms side by side
    // instead of one thread's K * K -- the walk is a chain of nInt dependent ops per pattern, and its latency, not its work, is
    // what this kernel costs.  Same association per output state as a one-thread walk: not a bit changes.
    auto walk = [&](double *scr, int32_t *scrE) {
        const int grp = tid >> 3, i = tid & 7, nGrp = nt >> 3;
        const bool lane_on = i < K;
        const int ii = lane_on ? i : 0;
        for (int p0 = 0; p0 < nPat; p0 += nGrp) {   // uniform trip count: every lane takes part in the shuffles
            const int p = p0 + grp;
            const bool on = p < nPat;
            const int pp = on ? p : 0;
            for (int c = 0; c < C; c++) {
                // one post-order pass for this (pattern, category)
                for (int oi = 0; oi < nI; oi++) {
                    const int v = order[oi], c1 = gL[v], c2 = gR[v];
                    double out = 1.0;
                    int cum = 0;
#pragma unroll
                    for (int side = 0; side < 2; side++) {
                        const int ch = side ? c2 : c1;
                        const double *M = mats + ((size_t)ch * C + c) * KK + ii * K;   // row ii of the child's matrix
                        if (ch < nT) {
                            const int st = tipsInSmem ? sStates[ch * nPat + pp] : gStates[(size_t)ch * d.stateStride + pp];
                            const double f = (st < K) ? M[st] : 1.0;    // gap / '?': factor 1
                            out = side ? out * f : f;
                        } else {
                            const size_t e = ((size_t)(ch - nT) * C + c) * nPat + pp;
                            cum += scrE[e];
                            double sum = 0.0;
#pragma unroll
                            for (int j = 0; j < K; j++) sum += M[j] * scr[e * K + j];
                            out = side ? out * sum : sum;
                        }
                    }
                    // exact power-of-two rescale (partials are non-negative: the largest has the largest high word)
                    int hmax = lane_on ? __double2hiint(out) : 0;
#pragma unroll
                    for (int off = 4; off > 0; off >>= 1) hmax = max(hmax, __shfl_xor_sync(0xffffffffu, hmax, off, 8));
                    int field = hmax >> 20;
                    if (field <= 0 || field > 1023) field = 1023;
                    const double scale = __hiloint2double((2046 - field) << 20, 0);
                    const size_t e = ((size_t)(v - nT) * C + c) * nPat + pp;
                    if (on && lane_on) scr[e * K + i] = out * scale;
                    if (on && i == 0) scrE[e] = cum + field - 1023;
                    __syncwarp();   // the group's K values and the exponent are visible to its lanes before the parent reads them
                }
            }
            // integratePartials over the categories at a common exponent, frequencies, log
            const size_t r0 = (size_t)(root - nT) * C * nPat + pp;
            int E = -2147483647;
            for (int c = 0; c < C; c++) E = max(E, scrE[r0 + (size_t)c * nPat]);
            double acc = 0.0;
            for (int c = 0; c < C; c++) {
                const size_t e = r0 + (size_t)c * nPat;
                const int diff = scrE[e] - E;
                const double f = (diff < -1022) ? 0.0 : __hiloint2double((1023 + diff) << 20, 0);
                const double v = scr[e * K + ii] * f * props[c];
                acc = (c == 0) ? v : acc + v;
            }
            const double term = freqs[ii] * acc;
            double sum = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < K; s2++) sum += __shfl_sync(0xffffffffu, term, (tid & 24) + s2);   // states in order, as a single thread adds them
            if (on && i == 0) {
                const double lk = (E > -1000) ? log(ldexp(sum, E)) : (log(sum) + E * 0.6931471805599453);
                A.patLogL[d.patBase + p] = lk;
                if (p < KS_SMEM_PATTERNS) spl[p] = lk;   // the closing sum reads it back from shared memory, not through L2
            }
        }
    };
    // two instantiations of the walk: the scratch in shared memory (LDS / STS) or in HBM -- one generic pointer would make every access generic
    if (inSmem) walk(reinterpret_cast<double *>(smem + m.scr), reinterpret_cast<int32_t *>(smem + m.scr + 8 * K * nElem));
    else walk(A.scratch + d.kBase, A.scratchExp + d.kBase / K);
    __syncthreads();
    if (tid == 0) {   // TreeLikelihood.calcLogP, with Alignment.getAscertainmentCorrection when patterns are excluded
        const double *plG = A.patLogL + d.patBase;
        auto pl = [&](int k) { return k < KS_SMEM_PATTERNS ? spl[k] : plG[k]; };
        const int32_t *w = A.weights + d.patBase;
        double logP = 0.0;
        if (d.nExcl > 0) {
            double excludeProb = 0.0, returnProb = 1.0;
            for (int i = 0; i < d.nExcl; i++) excludeProb += exp(pl(A.excluded[d.exclBase + i]));
            returnProb -= excludeProb;
            const double corr = log(returnProb);
            for (int k = 0; k < nPat; k++) logP += (pl(k) - corr) * w[k];
        } else {
            for (int k = 0; k < nPat; k++) logP += pl(k) * w[k];
        }
        A.cur.logL[l] = logP;
    }
}

__global__ void __launch_bounds__(KS_THREADS) k_kstate(KStateArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int l = A.kLoci[blockIdx.x];
    const uint8_t fl = A.uniformFlags >= 0 ? (uint8_t)A.uniformFlags
                     : (A.singleLocus >= 0 ? (uint8_t)(l == A.singleLocus ? A.singleFlags : 0) : A.flags[l]);
    if (!(fl & LF_RUN)) return;   // the cached log-likelihood stands
    const LocusDesc d = A.loci[l];
    switch (d.kStates) {
    case 2: kstate_locus<2>(A, l, d, smem_raw); break;
    case 3: kstate_locus<3>(A, l, d, smem_raw); break;
    case 4: kstate_locus<4>(A, l, d, smem_raw); break;
    case 5: kstate_locus<5>(A, l, d, smem_raw); break;
    default: break;
    }
}

cudaError_t launch_kstate(const KStateArgs &a, int nK, int maxK, cudaStream_t st)
{
    const size_t smem = kstate_smem_bytes(a.maxNodes, a.maxCat, maxK, a.smemScratch);
    static size_t configured[MAX_DEVICES + 1] = {};
    cudaError_t rc = ensure_dynamic_smem(k_kstate, smem, configured);
    if (rc != cudaSuccess) return rc;
    k_kstate<<<nK, KS_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sb2
