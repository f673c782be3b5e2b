// Pieces of the Felsenstein-pruning walk shared by k_treewalk (treewalk.cu, the streaming kernel) and k_step (step.cu, the
// fused one-launch step): the two must produce the same bits for the same op list, so the arithmetic lives here once.
// All floating-point contractions are explicit (fma / __dmul_rn / __dadd_rn): the result does not depend on -fmad.
#pragma once
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
__device__ __forceinline__ void ld256_coherent(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// s[r][i] = sum_j M[i][j] a_r[j] for this thread's R patterns, folded into out (first operand: out = s, second: out *= s)
template <bool EXACT, int R, bool FIRST>
__device__ __forceinline__ void rows_times(const double *M, const double (&a)[R][4], double (&out)[R][4])
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double2 m01 = *reinterpret_cast<const double2 *>(M + i * 4);
        const double2 m23 = *reinterpret_cast<const double2 *>(M + i * 4 + 2);
#pragma unroll
        for (int r = 0; r < R; r++) {
            double s;
            if (EXACT)   // the reference's order, no fused multiply-add: ((m0 a0 + m1 a1) + m2 a2) + m3 a3
                s = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(m01.x, a[r][0]), __dmul_rn(m01.y, a[r][1])), __dmul_rn(m23.x, a[r][2])), __dmul_rn(m23.y, a[r][3]));
            else
                s = fma(m23.y, a[r][3], fma(m23.x, a[r][2], fma(m01.y, a[r][1], __dmul_rn(m01.x, a[r][0]))));
            out[r][i] = FIRST ? s : __dmul_rn(out[r][i], s);
        }
    }
}

// Exact power-of-two rescale of one (pattern, category) result: partials are non-negative, so the largest value has the
// largest high word and its exponent field is the scale.  Returns the exponent added to the subtree's cumulative exponent.
__device__ __forceinline__ int rescale4(const double (&v)[4], double &o0, double &o1, double &o2, double &o3)
{
    const int hmax = max(max(__double2hiint(v[0]), __double2hiint(v[1])), max(__double2hiint(v[2]), __double2hiint(v[3])));
    int field = hmax >> 20;                                             // 0 for zero / denormal, >= 0x7ff for inf / nan
    if (field <= 0 || field > 1023) field = 1023;                       // no scaling: nothing to scale / never scale down
    const double scale = __hiloint2double((2046 - field) << 20, 0);     // 2^-(field-1023), exact
    o0 = __dmul_rn(v[0], scale); o1 = __dmul_rn(v[1], scale); o2 = __dmul_rn(v[2], scale); o3 = __dmul_rn(v[3], scale);
    return field - 1023;
}

// Root of one pattern: integratePartials over the categories (each carries its own exponent: brought to the largest one,
// exactly), + pInv on constant patterns, frequencies, log.  The root's scaled partials of category cc sit in stack slot 0 at
// index si0 + cc * catStride of both planes (st2 / st2 + planeStride), exponents in sexp.
template <bool EXACT>
__device__ __forceinline__ double root_pattern_logl(const double2 *st2, int planeStride, const int16_t *sexp, int si0, int catStride, int C,
                                                    const double *props, const double *freqs, double pInv, int mask)
{
    int E = -32768;
    for (int cc = 0; cc < C; cc++) E = max(E, (int)sexp[si0 + cc * catStride]);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int cc = 0; cc < C; cc++) {
        const int si = si0 + cc * catStride;
        const double2 lo = st2[si], hi = st2[planeStride + si];
        const int diff = (int)sexp[si] - E;                          // <= 0
        const double f = (diff < -1022) ? 0.0 : __hiloint2double((1023 + diff) << 20, 0);
        const double pr = props[cc];
        const double v0 = __dmul_rn(lo.x, f), v1 = __dmul_rn(lo.y, f), v2 = __dmul_rn(hi.x, f), v3 = __dmul_rn(hi.y, f);
        if (cc == 0) {
            acc[0] = __dmul_rn(v0, pr); acc[1] = __dmul_rn(v1, pr); acc[2] = __dmul_rn(v2, pr); acc[3] = __dmul_rn(v3, pr);
        } else if (EXACT) {
            acc[0] = __dadd_rn(acc[0], __dmul_rn(v0, pr)); acc[1] = __dadd_rn(acc[1], __dmul_rn(v1, pr));
            acc[2] = __dadd_rn(acc[2], __dmul_rn(v2, pr)); acc[3] = __dadd_rn(acc[3], __dmul_rn(v3, pr));
        } else {
            acc[0] = fma(v0, pr, acc[0]); acc[1] = fma(v1, pr, acc[1]); acc[2] = fma(v2, pr, acc[2]); acc[3] = fma(v3, pr, acc[3]);
        }
    }
    if (mask) {
        // unscale first (underflow here is harmless: pInv dominates), then add pInv on the constant states
        double sum = 0.0;
        for (int s = 0; s < 4; s++) {
            double v = ldexp(acc[s], E);
            if (mask & (1 << s)) v = __dadd_rn(v, pInv);
            sum = __dadd_rn(sum, __dmul_rn(freqs[s], v));
        }
        return log(sum);
    }
    double sum = 0.0;
    for (int s = 0; s < 4; s++) sum = __dadd_rn(sum, __dmul_rn(freqs[s], acc[s]));
    // 2^E is exact, so when the unscaled value is representable this is the reference's log(sum) bit for bit
    return (E > -1000) ? log(ldexp(sum, E)) : fma((double)E, 0.6931471805599453, log(sum));
}

// per-pattern contribution lk * weight, summed over the 32 patterns of a chunk by a fixed-shape shuffle tree
__device__ __forceinline__ double chunk_sum(double contrib)
{
    for (int off = 16; off > 0; off >>= 1) contrib = __dadd_rn(contrib, __shfl_xor_sync(0xffffffffu, contrib, off));
    return contrib;
}

}  // namespace sb2
