// k_treewalk: the Felsenstein-pruning kernel (the roofline kernel of the path).
//
// Replaces BeerLikelihoodCore4.calculateStatesStatesPruning / calculateStatesPartialsPruning /
// calculatePartialsPartialsPruning, scalePartials, integratePartials, calculateLogLikelihoods and
// TreeLikelihood.calcLogP (BEAST.base >= 2.7.3, named by /root/reference/version.xml:2 and
// tutorial/CanisPhylogeny/Canis-Phylogeny.xml:684; restated from SURVEY.md Appendix B).
//
// B200-first decomposition: the pruning recursion is independent across site patterns, so the grid is
// (locus x tile of patterns) and ONE CTA walks the locus's whole dirty subtree for its tile, in the post-order
// schedule k_prepare emitted.  Results that a later op of the same walk consumes never leave the SM: they sit in a
// shared-memory stack of floor(log2(nInt))+1 tiles (heavy-child-first order bounds the depth).  Every result is
// also streamed to HBM once (256-bit stores) because later incremental (dirty-node-only) evaluations read clean
// siblings from there (256-bit non-allocating loads).  HBM traffic of a full evaluation is therefore the
// partials written once + tip states + exponents, instead of write + two re-reads per node.
//
// Thread mapping: one thread per (pattern, rate category); a warp is 32 consecutive patterns of one category, so
// global and shared accesses are fully coalesced / conflict free.  The two 4x4 matrices of an op are staged in
// shared memory once per CTA and read as broadcasts.
//
// Rescaling: every node, every pattern, by an exact power of two (the exponent of the largest of the pattern's
// C x 4 values); scaled values keep their mantissas, so rescaling never perturbs a result and evaluation is
// independent of the update history.  The cumulative exponent of a subtree is stored per (node, pattern) as int16.
#include <math.h>

#include "sb2_device.cuh"

namespace sb2 {

__device__ __forceinline__ void ld256_nc(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <bool EXACT>
__device__ __forceinline__ double dot4(const double *m, double a0, double a1, double a2, double a3)
{
    if (EXACT)   // the reference's order, no fused multiply-add: ((m0 a0 + m1 a1) + m2 a2) + m3 a3
        return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m[0], a0), __dmul_rn(m[1], a1)), __dmul_rn(m[2], a2)), __dmul_rn(m[3], a3));
    return fma(m[3], a3, fma(m[2], a2, fma(m[1], a1, m[0] * a0)));
}

size_t treewalk_smem_bytes(int nCat, int chunksPerTile, int stackDepth)
{
    const size_t warps = (size_t)nCat * chunksPerTile;
    size_t b = (size_t)stackDepth * warps * 32 * 4 * 8;   // partials stack
    b += (size_t)stackDepth * warps * 32 * 4;             // exponent stack
    b += 2ull * 2 * nCat * 16 * 8;                        // matrices, double buffered
    b += 2ull * warps * 32 * 4;                           // exponent exchange, double buffered
    b += warps * 8 + 64;
    return b;
}

template <bool EXACT>
__global__ void __launch_bounds__(1024) k_treewalk(WalkArgs A)
{
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int tile = blockIdx.x;
    const int l = A.tileLocus[tile];
    const int nOps = A.nOps[l];
    if (nOps == 0) return;   // locus clean: cached logL stands
    const LocusDesc d = A.loci[l];
    const int C = A.nCat, chunks = A.chunksPerTile, depth = A.stackDepth;
    const int warps = C * chunks;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const int c = w % C, chunk = w / C;
    const int nPat = d.nPat, nI = d.nInt;
    const int p = (tile - d.tileBase) * (chunks * 32) + chunk * 32 + lane;
    const bool valid = p < nPat;

    double *stack = reinterpret_cast<double *>(smem_raw);                          // [depth][warps][32][4]
    int *sexp = reinterpret_cast<int *>(stack + (size_t)depth * warps * 128);      // [depth][warps][32]
    double *smat = reinterpret_cast<double *>(sexp + (size_t)depth * warps * 32);  // [2][2][C][16]
    int *smax = reinterpret_cast<int *>(smat + 2 * 2 * C * 16);                    // [2][warps][32]
    double *sred = reinterpret_cast<double *>(smax + 2 * warps * 32);              // [warps]

    const Op *ops = A.ops + d.opBase;
    const double *mats = A.mats + d.matBase;
    const uint8_t *states = A.tipStates + d.stateBase;
    const size_t nodeStride = (size_t)C * nPat * 4;
    double *partials = A.partials + d.partBase;
    int16_t *cumExp = A.cumExp + d.expBase;
    const int myStack = (w * 32 + lane);
    const int nMatThreads = 32 * C;   // 2 matrices x C x 16 doubles

    for (int k = 0; k < nOps; k++) {
        const Op o = ops[k];
        const int par = k & 1;
        // stage the op's two branch matrices
        double *sm = smat + par * (2 * C * 16);
        for (int t = tid; t < nMatThreads; t += blockDim.x) {
            const int which = t >= 16 * C;
            const int r = t - which * 16 * C;
            sm[t] = mats[(size_t)(which ? o.mB : o.mA) * C * 16 + r];
        }
        __syncthreads();

        double out0 = 0, out1 = 0, out2 = 0, out3 = 0;
        int cum = 0;
        if (valid) {
            const double *M1 = sm + c * 16, *M2 = sm + (C + c) * 16;
            double sa0, sa1, sa2, sa3, sb0, sb1, sb2_, sb3;
            // operand a
            if (o.aKind == OPK_TIP) {
                const int s = states[(size_t)o.aOff * nPat + p];
                if (s < 4) { sa0 = M1[s]; sa1 = M1[4 + s]; sa2 = M1[8 + s]; sa3 = M1[12 + s]; }
                else { sa0 = sa1 = sa2 = sa3 = 1.0; }
            } else {
                double a0, a1, a2, a3;
                if (o.aKind == OPK_STACK) {
                    const double4 *sp = reinterpret_cast<const double4 *>(stack + ((size_t)o.aSlot * warps * 32 + myStack) * 4);
                    const double4 v = *sp;
                    a0 = v.x; a1 = v.y; a2 = v.z; a3 = v.w;
                    cum += sexp[o.aSlot * warps * 32 + myStack];
                } else {
                    ld256_nc(partials + (size_t)o.aOff * nodeStride + ((size_t)c * nPat + p) * 4, a0, a1, a2, a3);
                    cum += cumExp[(size_t)o.aOff * nPat + p];
                }
                sa0 = dot4<EXACT>(M1, a0, a1, a2, a3); sa1 = dot4<EXACT>(M1 + 4, a0, a1, a2, a3);
                sa2 = dot4<EXACT>(M1 + 8, a0, a1, a2, a3); sa3 = dot4<EXACT>(M1 + 12, a0, a1, a2, a3);
            }
            // operand b
            if (o.bKind == OPK_TIP) {
                const int s = states[(size_t)o.bOff * nPat + p];
                if (s < 4) { sb0 = M2[s]; sb1 = M2[4 + s]; sb2_ = M2[8 + s]; sb3 = M2[12 + s]; }
                else { sb0 = sb1 = sb2_ = sb3 = 1.0; }
            } else {
                double a0, a1, a2, a3;
                if (o.bKind == OPK_STACK) {
                    const double4 *sp = reinterpret_cast<const double4 *>(stack + ((size_t)o.bSlot * warps * 32 + myStack) * 4);
                    const double4 v = *sp;
                    a0 = v.x; a1 = v.y; a2 = v.z; a3 = v.w;
                    cum += sexp[o.bSlot * warps * 32 + myStack];
                } else {
                    ld256_nc(partials + (size_t)o.bOff * nodeStride + ((size_t)c * nPat + p) * 4, a0, a1, a2, a3);
                    cum += cumExp[(size_t)o.bOff * nPat + p];
                }
                sb0 = dot4<EXACT>(M2, a0, a1, a2, a3); sb1 = dot4<EXACT>(M2 + 4, a0, a1, a2, a3);
                sb2_ = dot4<EXACT>(M2 + 8, a0, a1, a2, a3); sb3 = dot4<EXACT>(M2 + 12, a0, a1, a2, a3);
            }
            out0 = sa0 * sb0; out1 = sa1 * sb1; out2 = sa2 * sb2_; out3 = sa3 * sb3;
        }
        // exponent of the pattern's largest value over categories x states
        const double m = fmax(fmax(out0, out1), fmax(out2, out3));
        int field = (m > 0.0) ? ((__double2hiint(m) >> 20) & 0x7ff) : 1023;
        if (C > 1) {
            int *sx = smax + par * warps * 32;
            sx[w * 32 + lane] = (m > 0.0) ? field : 0;
            __syncthreads();
            int best = 0;
            for (int cc = 0; cc < C; cc++) best = max(best, sx[(chunk * C + cc) * 32 + lane]);
            field = best > 0 ? best : 1023;
        }
        if (field > 1023) field = 1023;   // never scale down (inf/nan guard)
        const int e = field - 1023;       // <= 0
        const double scale = __hiloint2double((1023 - e) << 20, 0);   // 2^-e, exact
        out0 *= scale; out1 *= scale; out2 *= scale; out3 *= scale;
        cum += e;
        if (valid) {
            *reinterpret_cast<double4 *>(stack + ((size_t)o.outSlot * warps * 32 + myStack) * 4) = make_double4(out0, out1, out2, out3);
            sexp[o.outSlot * warps * 32 + myStack] = cum;
            st256(partials + (size_t)o.outOff * nodeStride + ((size_t)c * nPat + p) * 4, out0, out1, out2, out3);
            if (c == 0) cumExp[(size_t)o.outOff * nPat + p] = (int16_t)cum;
        }
    }
    __syncthreads();

    // root: integratePartials over categories, (+pInv on constant patterns), frequencies, log, pattern weight
    double contrib = 0.0;
    if (c == 0 && valid) {
        const LocusParams *prm = &A.params[l];
        const double *freqs = (prm->substKind == 0) ? prm->subst + 1 : prm->subst + 36;
        const double *props = A.catProps + d.catBase;
        const int cum = sexp[chunk * C * 32 + lane];   // slot 0
        double acc[4];
        {
            const double4 v = *reinterpret_cast<const double4 *>(stack + ((size_t)(chunk * C) * 32 + lane) * 4);
            const double pr = props[0];
            acc[0] = v.x * pr; acc[1] = v.y * pr; acc[2] = v.z * pr; acc[3] = v.w * pr;
        }
        for (int cc = 1; cc < C; cc++) {
            const double4 v = *reinterpret_cast<const double4 *>(stack + ((size_t)(chunk * C + cc) * 32 + lane) * 4);
            const double pr = props[cc];
            if (EXACT) {
                acc[0] = __dadd_rn(acc[0], __dmul_rn(v.x, pr)); acc[1] = __dadd_rn(acc[1], __dmul_rn(v.y, pr));
                acc[2] = __dadd_rn(acc[2], __dmul_rn(v.z, pr)); acc[3] = __dadd_rn(acc[3], __dmul_rn(v.w, pr));
            } else {
                acc[0] = fma(v.x, pr, acc[0]); acc[1] = fma(v.y, pr, acc[1]); acc[2] = fma(v.z, pr, acc[2]); acc[3] = fma(v.w, pr, acc[3]);
            }
        }
        const double pInv = prm->pInv;
        const int mask = (pInv > 0.0) ? A.constMask[d.patBase + p] : 0;
        double lk;
        if (mask) {
            // unscale first (underflow here is harmless: pInv dominates), then add pInv on the constant states
            double sum = 0.0;
            for (int s = 0; s < 4; s++) {
                double v = ldexp(acc[s], cum);
                if (mask & (1 << s)) v = __dadd_rn(v, pInv);
                sum = __dadd_rn(sum, __dmul_rn(freqs[s], v));
            }
            lk = log(sum);
        } else {
            double sum = 0.0;
            for (int s = 0; s < 4; s++) sum = __dadd_rn(sum, __dmul_rn(freqs[s], acc[s]));
            // 2^cum is exact, so when the unscaled value is representable this is the reference's log(sum) bit for bit
            lk = (cum > -1000) ? log(ldexp(sum, cum)) : (log(sum) + cum * 0.6931471805599453);
        }
        A.patLogL[d.patBase + p] = lk;
        contrib = __dmul_rn(lk, (double)A.weights[d.patBase + p]);
    }
    // fixed-shape tile reduction (deterministic)
    for (int off = 16; off > 0; off >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, off);
    if (lane == 0) sred[w] = contrib;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int ch = 0; ch < chunks; ch++) t += sred[ch * C];
        A.tileSum[tile] = t;
    }
}

void launch_treewalk(const WalkArgs &a, int nTiles, bool exact, cudaStream_t st)
{
    const size_t smem = treewalk_smem_bytes(a.nCat, a.chunksPerTile, a.stackDepth);
    static size_t configured[2] = {0, 0};
    if (smem > 48 * 1024 && smem > configured[exact]) {
        if (exact) cudaFuncSetAttribute(k_treewalk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        else cudaFuncSetAttribute(k_treewalk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[exact] = smem;
    }
    const int threads = 32 * a.nCat * a.chunksPerTile;
    if (exact) k_treewalk<true><<<nTiles, threads, smem, st>>>(a);
    else k_treewalk<false><<<nTiles, threads, smem, st>>>(a);
}

}  // namespace sb2
