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

constexpr int OPS_BATCH = 64;   // ops staged in shared memory per batch (two halves)

size_t treewalk_smem_bytes(int nCat, int chunksPerTile, int stackDepth)
{
    const size_t warps = (size_t)nCat * chunksPerTile;
    size_t b = (size_t)stackDepth * warps * 32 * 4 * 8;   // partials stack
    b += (size_t)stackDepth * warps * 32 * 4;             // exponent stack
    b += 2ull * 2 * nCat * 16 * 8;                        // matrices, double buffered
    b += 2ull * warps * 32 * 4;                           // exponent exchange, double buffered
    b += warps * 8 + 64;
    b = (b + 15) & ~size_t(15);
    b += 2ull * OPS_BATCH * sizeof(Op);                   // op schedule, two half batches
    return b;
}

// raw operand data of one op for this thread's (pattern, category), loaded one op ahead of its use
struct Operand {
    double x0, x1, x2, x3;
    int aux;   // TIP: state; GLOBAL: cumulative exponent
};

__device__ __forceinline__ void fetch_operand(Operand &r, uint8_t kind, int off, const uint8_t *states, const double *partials,
                                              const int16_t *cumExp, size_t nodeStride, int nPat, int c, int p)
{
    if (kind == OPK_TIP) {
        r.aux = states[(size_t)off * nPat + p];
    } else if (kind == OPK_GLOBAL) {
        ld256_nc(partials + (size_t)off * nodeStride + ((size_t)c * nPat + p) * 4, r.x0, r.x1, r.x2, r.x3);
        r.aux = cumExp[(size_t)off * nPat + p];
    }
}

template <bool EXACT>
__device__ __forceinline__ void apply_operand(const Operand &r, uint8_t kind, int slot, const double *M, const double *stack,
                                              const int *sexp, int warps, int myStack, double &s0, double &s1, double &s2,
                                              double &s3, int &cum)
{
    if (kind == OPK_TIP) {
        const int s = r.aux;
        if (s < 4) { s0 = M[s]; s1 = M[4 + s]; s2 = M[8 + s]; s3 = M[12 + s]; }
        else { s0 = s1 = s2 = s3 = 1.0; }
        return;
    }
    double a0, a1, a2, a3;
    if (kind == OPK_STACK) {
        const double4 v = *reinterpret_cast<const double4 *>(stack + ((size_t)slot * warps * 32 + myStack) * 4);
        a0 = v.x; a1 = v.y; a2 = v.z; a3 = v.w;
        cum += sexp[slot * warps * 32 + myStack];
    } else {
        a0 = r.x0; a1 = r.x1; a2 = r.x2; a3 = r.x3;
        cum += r.aux;
    }
    s0 = dot4<EXACT>(M, a0, a1, a2, a3); s1 = dot4<EXACT>(M + 4, a0, a1, a2, a3);
    s2 = dot4<EXACT>(M + 8, a0, a1, a2, a3); s3 = dot4<EXACT>(M + 12, a0, a1, a2, a3);
}

template <bool EXACT, int MAXT>
__global__ void __launch_bounds__(MAXT, (MAXT <= 256 ? 3 : 1)) k_treewalk(WalkArgs A)
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
    const int nPat = d.nPat;
    const int p = (tile - d.tileBase) * (chunks * 32) + chunk * 32 + lane;
    const bool valid = p < nPat;

    double *stack = reinterpret_cast<double *>(smem_raw);                          // [depth][warps][32][4]
    int *sexp = reinterpret_cast<int *>(stack + (size_t)depth * warps * 128);      // [depth][warps][32]
    double *smat = reinterpret_cast<double *>(sexp + (size_t)depth * warps * 32);  // [2][2][C][16]
    int *smax = reinterpret_cast<int *>(smat + 2 * 2 * C * 16);                    // [2][warps][32]
    double *sred = reinterpret_cast<double *>(smax + 2 * warps * 32);              // [warps]
    Op *sops = reinterpret_cast<Op *>((reinterpret_cast<uintptr_t>(sred + warps) + 8 + 15) & ~uintptr_t(15));   // [2][OPS_BATCH]

    const Op *ops = A.ops + d.opBase;
    const double *mats = A.mats + d.matBase;
    const uint8_t *states = A.tipStates + d.stateBase;
    const size_t nodeStride = (size_t)C * nPat * 4;
    double *partials = A.partials + d.partBase;
    int16_t *cumExp = A.cumExp + d.expBase;
    const int myStack = (w * 32 + lane);
    const int nMat = 32 * C;   // 2 matrices x C x 16 doubles, one per staging thread
    const bool stager = tid < nMat;
    const int mWhich = tid >= 16 * C, mR = tid - mWhich * 16 * C;

    auto load_batch = [&](int b) {
        const int first = b * OPS_BATCH, cnt = min(OPS_BATCH, nOps - first);
        const int4 *src = reinterpret_cast<const int4 *>(ops + first);
        int4 *dst = reinterpret_cast<int4 *>(sops + (b & 1) * OPS_BATCH);
        for (int t = tid; t < cnt * 2; t += blockDim.x) dst[t] = src[t];
    };
    load_batch(0);
    __syncthreads();

    // software pipeline: while op k is computed, op k+1's matrices, tip states and clean-child partials are in flight
    Op o = sops[0];
    double mreg = 0.0;
    Operand pa, pb;
    pa.x0 = pa.x1 = pa.x2 = pa.x3 = 0.0; pa.aux = 0; pb = pa;
    if (stager) mreg = mats[(size_t)(mWhich ? o.mB : o.mA) * C * 16 + mR];
    if (valid) {
        fetch_operand(pa, o.aKind, o.aOff, states, partials, cumExp, nodeStride, nPat, c, p);
        fetch_operand(pb, o.bKind, o.bOff, states, partials, cumExp, nodeStride, nPat, c, p);
    }
    for (int k = 0; k < nOps; k++) {
        const int par = k & 1;
        double *sm = smat + par * (2 * C * 16);
        if (stager) sm[tid] = mreg;
        if (k + 1 < nOps && (k + 1) % OPS_BATCH == 0) load_batch((k + 1) / OPS_BATCH);   // other half: its ops are all consumed
        __syncthreads();

        Op on = o;
        Operand na = pa, nb = pb;
        if (k + 1 < nOps) {
            on = sops[(((k + 1) / OPS_BATCH) & 1) * OPS_BATCH + (k + 1) % OPS_BATCH];
            if (stager) mreg = mats[(size_t)(mWhich ? on.mB : on.mA) * C * 16 + mR];
            if (valid) {
                fetch_operand(na, on.aKind, on.aOff, states, partials, cumExp, nodeStride, nPat, c, p);
                fetch_operand(nb, on.bKind, on.bOff, states, partials, cumExp, nodeStride, nPat, c, p);
            }
        }

        double out0 = 0, out1 = 0, out2 = 0, out3 = 0;
        int cum = 0;
        if (valid) {
            const double *M1 = sm + c * 16, *M2 = sm + (C + c) * 16;
            double sa0, sa1, sa2, sa3, sb0, sb1, sb2_, sb3;
            apply_operand<EXACT>(pa, o.aKind, o.aSlot, M1, stack, sexp, warps, myStack, sa0, sa1, sa2, sa3, cum);
            apply_operand<EXACT>(pb, o.bKind, o.bSlot, M2, stack, sexp, warps, myStack, sb0, sb1, sb2_, sb3, cum);
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
        o = on; pa = na; pb = nb;
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

template <bool EXACT, int MAXT>
static void launch_one(const WalkArgs &a, int nTiles, size_t smem, int threads, cudaStream_t st)
{
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(k_treewalk<EXACT, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    k_treewalk<EXACT, MAXT><<<nTiles, threads, smem, st>>>(a);
}

void launch_treewalk(const WalkArgs &a, int nTiles, bool exact, cudaStream_t st)
{
    const size_t smem = treewalk_smem_bytes(a.nCat, a.chunksPerTile, a.stackDepth);
    const int threads = 32 * a.nCat * a.chunksPerTile;
    if (threads <= 256) {
        if (exact) launch_one<true, 256>(a, nTiles, smem, threads, st);
        else launch_one<false, 256>(a, nTiles, smem, threads, st);
    } else {   // more than 8 rate categories: register-capped variant
        if (exact) launch_one<true, 1024>(a, nTiles, smem, threads, st);
        else launch_one<false, 1024>(a, nTiles, smem, threads, st);
    }
}

}  // namespace sb2
