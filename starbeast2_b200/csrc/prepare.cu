// k_prepare: one CTA per locus, the whole locus in shared memory (body: prepare_body.cuh, shared with the fused one-launch
// step of step.cu).  First of the three kernels of the streaming path  k_prepare -> k_treewalk -> k_reduce_*.
//
// Compiled with -fmad=false (see prepare_body.cuh).
#include "prepare_body.cuh"

namespace sb2 {

size_t prepare_smem_bytes(int maxNodes, int S) { return prep_carve(nullptr, nullptr, maxNodes, S); }

__global__ void __launch_bounds__(PREP_THREADS) k_prepare(PrepareArgs A, int maxNodes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int l = blockIdx.x, tid = threadIdx.x;
    // let the pruning kernel's CTAs launch now: they stage their static inputs and then wait for this grid to complete
    cudaTriggerProgrammaticLaunchCompletion();
    if (A.nSync) {
        bool mine = false;
        for (int j = 0; j < A.nSync; j++) mine |= A.syncLoci[j] == l;
        if (mine) {
            sync_stored_slices(A, l);
            __syncthreads();   // the stored-state reads below (buffer indices) must see these writes
        }
    }
    const uint8_t fl = A.uniformFlags >= 0 ? (uint8_t)A.uniformFlags
                     : (A.singleLocus >= 0 ? (uint8_t)(l == A.singleLocus ? A.singleFlags : 0) : A.flags[l]);
    if (!(fl & LF_RUN)) {
        if (tid == 0) A.nOps[l] = 0;
        return;
    }
    if (A.loci[l].kStates) {   // K-state likelihood-only locus: no embedding, no schedule -- only its tree becomes current (k_kstate reads it)
        const LocusDesc d = A.loci[l];
        if (fl & LF_TREE_IN) {
            const double *inH = reinterpret_cast<const double *>(A.inTree + d.treeOff);
            const int *inP = reinterpret_cast<const int *>(inH + d.nNodes), *inL = inP + d.nNodes, *inR = inL + d.nNodes;
            for (int i = tid; i < d.nNodes; i += blockDim.x) {
                A.cur.gHeight[d.nodeBase + i] = inH[i]; A.cur.gParent[d.nodeBase + i] = inP[i];
                A.cur.gLeft[d.nodeBase + i] = inL[i]; A.cur.gRight[d.nodeBase + i] = inR[i];
            }
        }
        if (tid == 0) A.nOps[l] = 0;
        return;
    }
    __shared__ LocusParams sprm;
    prepare_locus<false>(A, maxNodes, l, fl, smem_raw, sprm, PrepFused{});
}

cudaError_t launch_prepare(const PrepareArgs &a, int maxNodes, cudaStream_t st)
{
    const size_t smem = prepare_smem_bytes(maxNodes, a.S);
    static size_t configured[MAX_DEVICES + 1] = {};
    cudaError_t rc = ensure_dynamic_smem(k_prepare, smem, configured);
    if (rc != cudaSuccess) return rc;
    k_prepare<<<a.nLoci, PREP_THREADS, smem, st>>>(a, maxNodes);
    return cudaGetLastError();
}

}  // namespace sb2
