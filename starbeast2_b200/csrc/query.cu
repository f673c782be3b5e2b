// Operator-side walks over every gene tree, answered from the trees the device already holds.
//
// The reference answers two questions per species-tree proposal by a recursive walk of EVERY gene tree on one Java thread,
// with HashSet<String> taxon lookups at the leaves:
//   * CoordinatedUniform / CoordinatedExponential.getConnectingNodes  (CoordinatedUniform.java:94-177,
//     CoordinatedExponential.java:92-160): which gene nodes move together with species node s, and how far they may move;
//   * NodeReheight2.recalculateMaxHeight (NodeReheight2.java:150-203): the lowest gene coalescence that joins lineages from
//     the two sides of the chosen species node.
// Both recursions only ask, per gene node, WHICH CLASSES OF TIPS sit below it.  With class bits {1: under the left child /
// left of the centre, 2: right, 4: elsewhere} and state(g) = OR of the classes of g's tips:
//   descendsThrough(g) = LEFT_ONLY (1) | RIGHT_ONLY (2) | BOTH (3) | NEITHER (>= 4)                 (:119-177)
//   connecting(g)      = g internal and state(g) == 3                                                 (every add() in :133-175)
//   tipwardFreedom     = min over connecting g, children c with state(c) != 3, of h(g) - h(c)         (:152,:162,:171-172)
//   rootwardFreedom    = min over connecting g whose parent p has state(p) >= 4, of h(p) - h(g)       (:146,:157)
//   maxHeight          = min over g whose two children have states {1, 2}, of h(g)                    (NodeReheight2.java:195-200)
// min() over the same set of doubles is order-independent, so the results are bit-identical to the Java.
//
// One CTA per locus: tips OR their class rootward through a shared-memory copy of the parent array (a lineage stops at the
// first ancestor that already has its bit), then one thread per internal node classifies and the CTA min-reduces.
#include "sb2_device.cuh"

namespace sb2 {

__device__ __forceinline__ unsigned long long enc_min(double d)   // order-preserving map double -> u64
{
    const long long b = __double_as_longlong(d);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}

__device__ __forceinline__ double warp_min(double v)
{
    for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

__global__ void __launch_bounds__(QUERY_THREADS) k_tip_class_query(QueryArgs A)
{
    extern __shared__ int sm[];
    const LocusDesc &d = A.loci[blockIdx.x];
    const int G = d.nNodes, nTips = d.nTips, nb = d.nodeBase, tid = threadIdx.x;
    if (d.kStates) {   // a likelihood-only locus is not a gene tree: the co-ordinated operators do not see it
        if (A.mask) for (int i = tid; i < G; i += QUERY_THREADS) A.mask[nb + i] = 0;
        return;
    }
    int *par = sm;
    unsigned *state = reinterpret_cast<unsigned *>(sm + A.maxNodes);
    for (int i = tid; i < G; i += QUERY_THREADS) { par[i] = A.gParent[nb + i]; state[i] = 0u; }
    __syncthreads();
    for (int t = tid; t < nTips; t += QUERY_THREADS) {
        const unsigned cls = A.leafClass[A.tipSpecies[d.tipBase + t]];
        for (int g = t; g >= 0; g = par[g])
            if (atomicOr(&state[g], cls) & cls) break;   // an earlier lineage already carried this class to the root
    }
    __syncthreads();
    double tipward = INFINITY, rootward = INFINITY, maxH = INFINITY;
    int count = 0;
    for (int g = tid; g < G; g += QUERY_THREADS) {
        if (g < nTips) { if (A.mask) A.mask[nb + g] = 0; continue; }
        const int l = A.gLeft[nb + g], r = A.gRight[nb + g];
        const unsigned s = state[g], sl = state[l], sr = state[r];
        const double h = A.gHeight[nb + g];
        if (A.mode == QUERY_CONNECTING) {
            const bool conn = s == 3u;
            if (A.mask) A.mask[nb + g] = conn;
            if (conn) {
                count++;
                if (sl != 3u) tipward = fmin(tipward, h - A.gHeight[nb + l]);
                if (sr != 3u) tipward = fmin(tipward, h - A.gHeight[nb + r]);
                const int p = par[g];
                if (p >= 0 && state[p] >= 4u) rootward = fmin(rootward, A.gHeight[nb + p] - h);
            }
        } else {
            if ((sl == 1u && sr == 2u) || (sl == 2u && sr == 1u)) maxH = fmin(maxH, h);
        }
    }
    tipward = warp_min(tipward); rootward = warp_min(rootward); maxH = warp_min(maxH);
    for (int off = 16; off > 0; off >>= 1) count += __shfl_xor_sync(0xffffffffu, count, off);
    if ((tid & 31) == 0) {
        if (tipward < INFINITY) atomicMin(&A.result[0], enc_min(tipward));
        if (rootward < INFINITY) atomicMin(&A.result[1], enc_min(rootward));
        if (maxH < INFINITY) atomicMin(&A.result[2], enc_min(maxH));
        if (count) atomicAdd(reinterpret_cast<unsigned long long *>(&A.result[3]), (unsigned long long)count);
    }
}

cudaError_t launch_tip_class_query(const QueryArgs &a, int nLoci, cudaStream_t st)
{
    const size_t smem = 8 * (size_t)a.maxNodes;
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static size_t configured[MAX_DEVICES + 1] = {};
    cudaError_t rc = ensure_dynamic_smem(k_tip_class_query, smem, configured);
    if (rc != cudaSuccess) return rc;
    k_tip_class_query<<<nLoci, QUERY_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sb2
