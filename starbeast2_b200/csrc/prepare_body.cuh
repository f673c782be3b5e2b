// prepare_locus: the body of k_prepare (prepare.cu: one CTA per locus, the whole locus in shared memory) -- and, with FUSED,
// of k_step (step.cu), where every CTA of a locus's cluster repeats it for itself and only the cluster's leader writes state.
//
// Does, for every locus flagged LF_RUN, everything of the path that is tree-shaped and latency-bound:
//   1. takes the incoming gene tree (if any), diffs its topology against the device-resident one;
//   2. embeds the gene tree in the species tree -- GeneTree.update()/collateCoalescenceEvents
//      (/root/reference/src/starbeast2/GeneTree.java:202-318, 332-379), one thread per gene leaf, lineages
//      race rootward and claim nodes with atomicCAS (the reference walks them one after another);
//      while climbing, each gene edge accumulates its StarBeastClock rate (StarBeastClock.java:54-75)
//      directly from the species path it crosses -- the G x S speciesOccupancy matrix is never materialised;
//   3. sorts coalescent times per species branch (GeneTree.getCoalescentTimes, GeneTree.java:387-405) and
//      evaluates the per-branch population-model terms: ConstantPopulations.constantLogP
//      (ConstantPopulations.java:95-118), LinearWithConstantRoot.branchLogP/linearLogP
//      (LinearWithConstantRoot.java:50-70, 155-201) or the analytic-integration summands
//      (MultispeciesCoalescent.java:207-225);
//   4. decides which branch matrices and node partials are dirty the way TreeLikelihood.traverse does
//      (length*rate changed / model dirty / child updated), flips the double-buffer indices
//      (BeerLikelihoodCore.setNodeMatrixForUpdate / setNodePartialsForUpdate) and emits a post-order,
//      heavy-child-first pruning schedule whose live results fit a stack of floor(log2(nInt))+1 tiles;
//   5. computes P = exp(Q d) for every dirty branch x category (HKY closed form or eigen system).
//
// Every translation unit that includes this file is compiled with -fmad=false: sums and products associate exactly as in the
// reference's Java (no fused multiply-add), so these terms differ from it only through log/exp.
#pragma once
#include <math.h>

#include "sb2_device.cuh"

namespace sb2 {

constexpr int PREP_THREADS = 256;   // the matrix phase (G x C items, ~400 instructions each) and the rank sort scale with it

struct PrepSmem {
    double *gH, *rate, *times, *sH, *sRate, *btOld, *popA, *popB;
    int *gP, *gL, *gR, *assign, *cnt, *sP, *sL, *sR, *n, *k, *kstart, *misc;
    uint8_t *matDirty, *dirty, *topo, *stM, *stP, *curP, *curM;
    int *first;   // per internal node with two dirty children: the child evaluated first; else -1
    // multi-worker walks: per side-subtree root, the helper worker that evaluates it (0 = stays with worker 0), its mailbox
    // slot and the offset of its ops inside that worker's list; `path` is scratch for the main path
    uint8_t *wk, *mbx;
    int *woff, *path;
    int *pos;     // schedule position of each dirty internal node
};

__host__ __device__ inline size_t prep_carve(PrepSmem *s, unsigned char *base, int G, int S)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 7) & ~size_t(7); return r; };
    size_t gH = take(8ull * G), rate = take(8ull * G), times = take(8ull * G), sH = take(8ull * S), sRate = take(8ull * S);
    size_t btOld = take(8ull * G), popA = take(8ull * S), popB = take(8ull * S);
    size_t gP = take(4ull * G), gL = take(4ull * G), gR = take(4ull * G), assign = take(4ull * G), cnt = take(4ull * G);
    size_t sP = take(4ull * S), sL = take(4ull * S), sR = take(4ull * S), n = take(4ull * S), k = take(4ull * S), ks = take(4ull * (S + 1));
    size_t misc = take(4ull * 8), first = take(4ull * G);
    size_t md = take(G), di = take(G), to = take(G), stM = take(G), stP = take(G), cuP = take(G), cuM = take(G);
    size_t wk = take(G), mbx = take(G), woff = take(4ull * G), path = take(4ull * G), pos = take(4ull * G);
    if (s) {
        s->pos = (int *)(base + pos);
        s->wk = base + wk; s->mbx = base + mbx; s->woff = (int *)(base + woff); s->path = (int *)(base + path);
        s->gH = (double *)(base + gH); s->rate = (double *)(base + rate); s->times = (double *)(base + times);
        s->sH = (double *)(base + sH); s->sRate = (double *)(base + sRate);
        s->gP = (int *)(base + gP); s->gL = (int *)(base + gL); s->gR = (int *)(base + gR);
        s->assign = (int *)(base + assign); s->cnt = (int *)(base + cnt);
        s->sP = (int *)(base + sP); s->sL = (int *)(base + sL); s->sR = (int *)(base + sR);
        s->n = (int *)(base + n); s->k = (int *)(base + k); s->kstart = (int *)(base + ks); s->misc = (int *)(base + misc);
        s->matDirty = base + md; s->dirty = base + di; s->topo = base + to; s->first = (int *)(base + first);
        s->stM = base + stM; s->stP = base + stP; s->curP = base + cuP; s->curM = base + cuM;
        s->btOld = (double *)(base + btOld); s->popA = (double *)(base + popA); s->popB = (double *)(base + popB);
    }
    return o;
}


__device__ __forceinline__ void prep_st256(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// One gene edge: climb from species branch b (edge starts at height lastH) to the branch that holds height hg.
// Lineage counts: GeneTree.java:346.  Rate: StarBeastClock.java:60-70 sums r_j*occ_j over species node numbers j
// ascending (zeros are no-ops), occ_j from GeneTree.java:345,354; reproduced by selecting the path's nodes in
// ascending order.
__device__ __forceinline__ int climb_edge(int b, double lastH, double hg, const PrepSmem &s, bool wantRate,
                                          double clock, double *rateOut)
{
    const int b0 = b;
    int end = b, nseg = 1;
    while (s.sP[end] >= 0 && hg >= s.sH[s.sP[end]]) {   // note >= (GeneTree.java:337)
        end = s.sP[end];
        atomicAdd(&s.n[end], 1);
        nseg++;
    }
    if (wantRate) {
        double wsum = 0.0, len = 0.0;
        if (nseg == 1) {
            const double occ = hg - lastH;
            wsum = wsum + s.sRate[end] * occ;
            len = len + occ;
        } else {
            int prev = -1;
            for (int it = 0; it < nseg; it++) {
                int best = 0x7fffffff;
                double occBest = 0.0;
                int j = b0;
                double lo = lastH;
                for (;;) {
                    const double hi = (j == end) ? hg : s.sH[s.sP[j]];
                    if (j > prev && j < best) { best = j; occBest = hi - lo; }
                    if (j == end) break;
                    lo = s.sH[s.sP[j]];
                    j = s.sP[j];
                }
                wsum = wsum + s.sRate[best] * occBest;
                len = len + occBest;
                prev = best;
            }
        }
        *rateOut = clock * wsum / len;   // StarBeastClock.java:70
    }
    return end;
}

struct HkyTabs {
    double fA, fC, fG, fT, tab1A, tab2A, tab3A, tab1C, tab2C, tab3C, tab1G, tab2G, tab3G, tab1T, tab2T, tab3T, beta, A_R, A_Y;
};

// HKY.setupMatrix (BEAST.base; restated from upstream)
__device__ __forceinline__ void hky_setup(const double *sp, HkyTabs &t)
{
    const double kappa = sp[0];
    t.fA = sp[1]; t.fC = sp[2]; t.fG = sp[3]; t.fT = sp[4];
    const double freqR = t.fA + t.fG, freqY = t.fC + t.fT;
    const double r1 = (1 / freqR) - 1;
    t.tab1A = t.fA * r1; t.tab3A = t.fA / freqR; t.tab2A = 1 - t.tab3A;
    const double r2 = 1 / r1;
    t.tab1C = t.fC * r2; t.tab3C = t.fC / freqY; t.tab2C = 1 - t.tab3C;
    t.tab1G = t.fG * r1; t.tab3G = t.tab2A; t.tab2G = t.tab3A;
    t.tab1T = t.fT * r2; t.tab3T = t.tab2C; t.tab2T = t.tab3C;
    t.beta = -1.0 / (2.0 * (freqR * freqY + kappa * (t.fA * t.fG + t.fC * t.fT)));
    t.A_R = 1.0 + freqR * (kappa - 1);
    t.A_Y = 1.0 + freqY * (kappa - 1);
}

// HKY.getTransitionProbabilities
__device__ __forceinline__ void hky_matrix(const HkyTabs &t, double distance, double *m)
{
    const double xx = t.beta * distance;
    const double bbR = exp(xx * t.A_R), bbY = exp(xx * t.A_Y), aa = exp(xx);
    const double oneminusa = 1 - aa;
    const double t1Aaa = t.tab1A * aa, t1Gaa = t.tab1G * aa, t1Caa = t.tab1C * aa, t1Taa = t.tab1T * aa;
    m[0] = t.fA + t1Aaa + (t.tab2A * bbR);
    m[1] = t.fC * oneminusa;
    m[2] = t.fG + t1Gaa - (t.tab3G * bbR);
    m[3] = t.fT * oneminusa;
    m[4] = t.fA * oneminusa;
    m[5] = t.fC + t1Caa + (t.tab2C * bbY);
    m[6] = t.fG * oneminusa;
    m[7] = t.fT + t1Taa - (t.tab3T * bbY);
    m[8] = t.fA + t1Aaa - (t.tab3A * bbR);
    m[9] = m[1];
    m[10] = t.fG + t1Gaa + (t.tab2G * bbR);
    m[11] = m[3];
    m[12] = m[4];
    m[13] = t.fC + t1Caa - (t.tab3C * bbY);
    m[14] = m[6];
    m[15] = t.fT + t1Taa + (t.tab2T * bbY);
}

// GeneralSubstitutionModel.getTransitionProbabilities: |V exp(d L) V^-1|
__device__ __forceinline__ void eigen_matrix(const double *sp, double distance, double *m)
{
    const double *Eval = sp, *Evec = sp + 4, *Ievc = sp + 20;
    double iexp[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double temp = exp(distance * Eval[i]);
#pragma unroll
        for (int j = 0; j < 4; j++) iexp[i * 4 + j] = Ievc[i * 4 + j] * temp;
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double temp = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) temp += Evec[i * 4 + k] * iexp[k * 4 + j];
            m[i * 4 + j] = fabs(temp);
        }
}

// lazy store(): locus l changed in the previous step -> stored := cur for its slices, before anything moves
__device__ __forceinline__ void sync_stored_slices(const PrepareArgs &A, int l)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    {
        {
            const LocusDesc ds = A.loci[l];
            const int nbs = ds.nodeBase, Gs = ds.nNodes;
#pragma unroll 1
            for (int i = tid; i < Gs; i += nt) {
                const int k = nbs + i;
                const double h = A.cur.gHeight[k], br = A.cur.bRate[k], bt = A.cur.bTime[k];
                const int p = A.cur.gParent[k], L = A.cur.gLeft[k], R = A.cur.gRight[k], as = A.cur.assign[k];
                const uint8_t cp = A.cur.curP[k], cm = A.cur.curM[k];
                A.stored.gHeight[k] = h; A.stored.bRate[k] = br; A.stored.bTime[k] = bt;
                A.stored.gParent[k] = p; A.stored.gLeft[k] = L; A.stored.gRight[k] = R; A.stored.assign[k] = as;
                A.stored.curP[k] = cp; A.stored.curM[k] = cm;
            }
#pragma unroll 1
            for (int b = tid; b < A.S; b += nt) {
                const size_t o = (size_t)l * A.S + b;
                const double x = A.cur.brLogP[o], g = A.cur.anG[o], r = A.cur.anR[o];
                const int n = A.cur.brN[o], k = A.cur.brK[o];
                A.stored.brLogP[o] = x; A.stored.anG[o] = g; A.stored.anR[o] = r; A.stored.brN[o] = n; A.stored.brK[o] = k;
            }
            if (tid == 0) { A.stored.logL[l] = A.cur.logL[l]; A.stored.coal[l] = A.cur.coal[l]; }
        }
    }
}

// FUSED (k_step): the CTA is one of a thread-block cluster that serves locus l.  Every CTA runs the whole body into its OWN
// shared memory (schedule -> F.sops, op-ordered matrices -> F.smats, op count -> *F.nOpsOut), but only the leader writes the
// locus's state in HBM, and only after every CTA of the cluster has read what it needs of the old state (cluster barrier:
// all threads arrive after the load wave; the leader waits there, followers wait later -- see step.cu).
struct PrepFused {
    bool leader;       // this CTA owns the locus's global writes
    bool storedIsCur;  // the lazy store() copy of this locus is happening concurrently (by the leader): its stored buffer
                       // indices equal the current ones, read those instead
    Op *sops;          // [nInt]
    double *smats;     // [nInt][category][side][16]
    int *nOpsOut;      // shared-memory int
    unsigned long long *dbg;   // profile mode, first cluster's leader only: %globaltimer stamps 8.. of the phases below; else nullptr
    const uint8_t *treeSrc;    // the incoming gene tree when it travels in the kernel arguments (single-locus steps), else nullptr
    bool splitReturn;          // see prepare_locus: the halves of the CTA return separately
};
// Return value of prepare_locus<true>: PREP_ALL = every thread returned together after a __syncthreads(); PREP_HALVES = the
// likelihood half (threads >= PREP_THREADS/2) returned as soon as ITS work -- schedule and matrices -- was complete and has
// signalled named barrier 4 (bar.arrive); the coalescent half returned after its own work and must bar.sync 4 before it reads
// the schedule or the matrices.
enum { PREP_ALL = 0, PREP_HALVES = 1 };
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Must be called by all PREP_THREADS threads of the CTA, for a locus whose flags `fl` have LF_RUN set.  !FUSED: threads may
// return early (group A when it has nothing left to do); FUSED: every thread returns, converged, after a __syncthreads().
template <bool FUSED>
__device__ __forceinline__ int prepare_locus(const PrepareArgs &A, int maxNodes, int l, uint8_t fl, unsigned char *smem_raw,
                                              LocusParams &sprm, const PrepFused &F)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const bool owner = !FUSED || F.leader;   // writes the locus's state in HBM
    auto stamp = [&](int i, int who) {
        if (FUSED && F.dbg && tid == who) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); F.dbg[i] = t; }
    };
    const LocusDesc d = A.loci[l];
    const LocusParams *prm = &sprm;
    const int G = d.nNodes, nT = d.nTips, nI = d.nInt, S = A.S, nb = d.nodeBase, C = d.nCat;
    PrepSmem s;
    prep_carve(&s, smem_raw, maxNodes, S);
    const bool forceAll = fl & LF_FORCE_ALL;
    const bool forceMats = forceAll || (fl & LF_FORCE_MATS);

    // ---- 1. everything any phase needs from HBM, fetched in ONE wave of independent loads -------------------------------
    // (parameters, species tree, gene tree, cached per-node state).  With a cold L2 every dependent round trip costs about a
    // microsecond on this latency path, so each thread first issues all its loads (indices clamped instead of predicated,
    // so that nothing stands between them) and only then stores to shared memory.
    {
        constexpr int NP = (int)(sizeof(LocusParams) / 8);
        const double *prmSrc = reinterpret_cast<const double *>(&A.params[l]);
        double *prmDst = reinterpret_cast<double *>(&sprm);
        const bool treeIn = fl & LF_TREE_IN;
        const double *inH = reinterpret_cast<const double *>((FUSED && F.treeSrc) ? F.treeSrc : A.inTree + d.treeOff);
        const int *inP = reinterpret_cast<const int *>(inH + G), *inL = inP + G, *inR = inL + G;
        const int span = max(max(G, S), NP);
        for (int base = 0; base < span; base += nt) {
            const int i = base + tid;
            const int ip = min(i, NP - 1), is = min(i, S - 1), ig = min(i, G - 1), it = min(i, nT - 1);
            const int sp = A.cur.sParent[is], sl = A.cur.sLeft[is], sr = A.cur.sRight[is];
            const double sh = A.cur.sHeight[is], srt = A.cur.sRates[is], pa = A.cur.popA[is], pb = A.cur.popB[is];
            const int curL = A.cur.gLeft[nb + ig], curR = A.cur.gRight[nb + ig];
            double h;
            int p, L, R;
            if (treeIn) { h = inH[ig]; p = inP[ig]; L = inL[ig]; R = inR[ig]; }
            else { h = A.cur.gHeight[nb + ig]; p = A.cur.gParent[nb + ig]; L = curL; R = curR; }
            const int tipSp = A.tipSpecies[d.tipBase + it];
            const double bt = A.cur.bTime[nb + ig];
            const uint8_t cuP = A.cur.curP[nb + ig], cuM = A.cur.curM[nb + ig];
            uint8_t stM = cuM, stP = cuP;
            if (!FUSED || !F.storedIsCur) { stM = A.stored.curM[nb + ig]; stP = A.stored.curP[nb + ig]; }
            const double pv = prmSrc[ip];
            if (i < S) {
                s.sP[i] = sp; s.sL[i] = sl; s.sR[i] = sr; s.sH[i] = sh; s.sRate[i] = srt; s.popA[i] = pa; s.popB[i] = pb;
                s.n[i] = 0; s.k[i] = 0;
            }
            if (i < G) {
                s.gH[i] = h; s.gP[i] = p; s.gL[i] = L; s.gR[i] = R;
                s.topo[i] = treeIn && ((L != curL) || (R != curR));
                if (!FUSED && treeIn) { A.cur.gHeight[nb + i] = h; A.cur.gParent[nb + i] = p; A.cur.gLeft[nb + i] = L; A.cur.gRight[nb + i] = R; }
                s.assign[i] = (i < nT) ? tipSp : -1;   // GeneTree.java:243,247
                s.cnt[i] = 0;
                s.btOld[i] = bt; s.stM[i] = stM; s.stP[i] = stP; s.curP[i] = cuP; s.curM[i] = cuM;
            }
            if (i < NP) prmDst[i] = pv;
        }
        if (tid == 0) { s.misc[0] = 1; /* compatible */ s.misc[1] = G - 1; /* root */ }
    }
    if (!FUSED) {   // warm L2 with the pruning kernel's static inputs of this locus (fire and forget)
        auto warm = [&](const void *p, size_t bytes) {
            const char *c = reinterpret_cast<const char *>(p);
            for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)nt * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
        };
        warm(A.pfTipStates + d.stateBase, (size_t)nT * d.stateStride);
        warm(A.pfWeights + d.patBase, 4ull * d.nPat);
        warm(A.pfTileLocus + d.tileBase, 4ull * d.nTiles);
        warm(A.pfCatProps + d.catBase, 8ull * C);
        warm(A.pfConstMask + d.patBase, (size_t)d.nPat);
    }
    __syncthreads();
    stamp(8, 0);
    if (FUSED) {
        // every thread of the cluster has consumed the old state: from here on the leader may overwrite it
        cluster_arrive_release();
        if (F.leader) {
            cluster_wait_acquire();
            if (fl & LF_TREE_IN)
                for (int i = tid; i < G; i += nt) {
                    A.cur.gHeight[nb + i] = s.gH[i]; A.cur.gParent[nb + i] = s.gP[i]; A.cur.gLeft[nb + i] = s.gL[i]; A.cur.gRight[nb + i] = s.gR[i];
                }
        }
    }
    for (int i = tid; i < G; i += nt) if (s.gP[i] < 0) s.misc[1] = i;
    for (int i = tid; i < nT; i += nt) atomicAdd(&s.n[s.assign[i]], 1);   // leafCoalescentLineageCounts, GeneTree.java:222
    if (FUSED) asm volatile("cp.async.wait_all;" ::: "memory");   // the caller's asynchronous tile staging (static data, long landed)
    __syncthreads();
    const int root = s.misc[1];

    // ---- warp specialisation -----------------------------------------------------------------------------------------------
    // The multispecies-coalescent half (2. embedding, 3. sorted times + population terms) and the likelihood half (4. dirty
    // detection + pruning schedule, 5. transition matrices) only meet in the branch rates, so on this latency path they run
    // CONCURRENTLY: group A = threads 0..127, group B = threads 128..255, each with its own named barrier.  With a species-tree
    // clock B waits (barrier 3) until A's embedding walks have produced the rates; with a strict / explicit clock it starts at once.
    constexpr int gn = PREP_THREADS / 2;
    const int grpId = tid / gn, gt = tid - grpId * gn;
    auto gsync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grpId), "n"(gn) : "memory"); };
    const int clockKind = prm->clockKind;
    const double clockRate = prm->clockRate;
    const bool wantRate = clockKind == 1;
    __shared__ int s_lane[WALK_MAX_WORKERS + 1], s_hStart[WALK_MAX_WORKERS * WALK_MAILBOX], s_hCnt[WALK_MAX_WORKERS * WALK_MAILBOX], s_nH;
    if (grpId == 0) {
    // a follower CTA of a fused step needs this half only for the StarBeastClock rates (the embedding walks), never for the
    // coalescent terms
    if (owner || wantRate) {
    // ---- 2. embedding walks (one lineage per thread) ---------------------------------------------------
    for (int leaf = gt; leaf < nT; leaf += gn) {
        int last = leaf, g = s.gP[leaf], b = s.assign[leaf];
        double lastH = 0.0;   // GeneTree.java:260 (gene tips are taken to sit at height 0, quirk Q6)
        double topH = (s.sP[b] >= 0) ? s.sH[s.sP[b]] : INFINITY;   // height at which the lineage leaves its current species branch
        while (g >= 0) {
            const double hg = s.gH[g];
            const int gnext = s.gP[g];   // depends only on g: fetched together with the height
            if (wantRate || hg >= topH) {   // the lineage climbs, or every edge needs its StarBeastClock rate
                double r;
                b = climb_edge(b, lastH, hg, s, wantRate, clockRate, &r);
                if (wantRate) s.rate[last] = r;
                topH = (s.sP[b] >= 0) ? s.sH[s.sP[b]] : INFINITY;
            }
            const int old = atomicCAS(&s.assign[g], -1, b);   // first lineage to arrive claims the node (:356-359)
            if (old == -1) {
                atomicAdd(&s.k[b], 1);
                last = g; lastH = hg; g = gnext;
            } else {
                if (old != b) s.misc[0] = 0;                 // GeneTree.java:375-377 incompatible
                break;
            }
        }
    }
    gsync();
    stamp(9, 0);
    if (wantRate) {   // the rates are complete: release group B (it waits on barrier 3 for them)
        if (gt == 0) s.rate[root] = clockRate;   // StarBeastClock.java:72
        __threadfence_block();
        asm volatile("bar.arrive 3, %0;" ::"n"(PREP_THREADS) : "memory");
    }
    }
    if (owner) {
    const bool compat = s.misc[0] != 0;
    for (int i = gt; i < G; i += gn) A.cur.assign[nb + i] = s.assign[i];

    // ---- 3. per-branch sorted times + population-model terms ------------------------------------------------
    // rank sort of the internal nodes by (species branch, height): branch b's times are contiguous and ascending.
    // O(nInt^2) comparisons, spread over all threads: `split` threads share one node (each scans a strided part of
    // the list, 4 comparisons in flight) and combine their partial ranks with shuffles.
    {
        int split = 1;
        while (split < 32 && split * 2 * nI <= gn) split *= 2;   // power of two, split * nI <= gn
        for (int base = 0; base < nI; base += gn / split) {
            const int ii = base + gt / split, part = gt % split;
            const bool act = ii < nI;
            const int i = nT + (act ? ii : 0);
            const int bi = s.assign[i];
            const double hi = s.gH[i];
            int rank = 0;
            if (act) {
#pragma unroll 4
                for (int j = nT + part; j < G; j += split) {
                    const int bj = s.assign[j];
                    const double hj = s.gH[j];
                    rank += (bj < bi) || (bj == bi && (hj < hi || (hj == hi && j < i)));
                }
            }
            for (int off = split >> 1; off > 0; off >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, off);
            if (act && part == 0) s.times[rank] = hi;
        }
    }
    gsync();
    const double ploidy = d.ploidy;
    const double logPloidy = log(ploidy);
    for (int b = gt; b < S; b += gn) {
        const int n = s.n[b], k = s.k[b];
        int kstart = 0;                       // branch b's block of the rank-sorted times starts after all lower-numbered branches'
        for (int bb = 0; bb < b; bb++) kstart += s.k[bb];
        const double *ts = s.times + kstart;
        const double t0 = s.sH[b];
        const double tEnd = (s.sP[b] < 0) ? INFINITY : s.sH[s.sP[b]];   // GeneTree.java:394
        auto T = [&](int i) { return i == 0 ? t0 : (i <= k ? ts[i - 1] : tEnd); };
        double term = 0.0, g = 0.0;
        if (compat) {
            const bool isRoot = s.sP[b] < 0;
            bool linear = false;
            double tipSize = 0.0, topSize = 0.0;
            if (A.popKind == 1) {   // LinearWithConstantRoot.branchLogP, :50-70
                tipSize = (s.sL[b] < 0) ? s.popA[b] : (s.popB[s.sL[b]] + s.popB[s.sR[b]]);
                if (!isRoot) { linear = true; topSize = s.popB[b]; }
            }
            if (!linear) {
                // shared gamma loop: ConstantPopulations.java:96-103 == MultispeciesCoalescent.java:215-222
                for (int i = 0; i < k; i++) g += (T(i + 1) - T(i)) * (n - i) * (n - (i + 1.0)) / 2.0;
                if (n - k > 1) g += (T(k + 1) - T(k)) * (n - k) * (n - (k + 1.0)) / 2.0;
                if (A.popKind == 2) {
                    term = 0.0;   // analytic: summands only (GeneTree.java:182)
                } else {
                    const double popSize = (A.popKind == 0) ? s.popA[b] : tipSize;
                    const double branchGamma = g / ploidy;
                    const double branchLogR = -k * logPloidy;
                    term = branchLogR - (k * log(popSize)) - (branchGamma / popSize);   // ConstantPopulations.java:107
                }
            } else {   // LinearWithConstantRoot.linearLogP, :155-201
                const double fPopSizeTop = topSize * ploidy, fPopSizeBottom = tipSize * ploidy;
                const double d5 = fPopSizeTop - fPopSizeBottom;
                const double fTime0 = T(0);
                const double a = d5 / (T(k + 1) - fTime0);
                const double bb = fPopSizeBottom;
                double logP = 0.0;
                if (fabs(d5) < 1e-10) {
                    for (int i = 0; i <= k; i++) {
                        const double fTimeip1 = T(i + 1);
                        const double fPopSize = a * (fTimeip1 - fTime0) + bb;
                        if (i < k) logP += -log(fPopSize);
                        const int i1 = n - i;
                        logP -= (i1 * (i1 - 1.0) / 2.0) * (fTimeip1 - T(i)) / fPopSize;
                    }
                } else {
                    const double vv = bb - a * fTime0;
                    for (int i = 0; i <= k; i++) {
                        const double fPopSize = a * T(i + 1) + vv;
                        if (i < k) logP += -log(fPopSize);
                        const double f = fPopSize / (a * T(i) + vv);
                        const int i1 = n - i;
                        logP += -(i1 * (i1 - 1.0) / 2.0) * log(f) / a;
                    }
                }
                term = logP;
            }
        }
        const size_t o = (size_t)l * S + b;
        A.cur.brLogP[o] = term;
        A.cur.brN[o] = n;
        A.cur.brK[o] = k;
        A.cur.anG[o] = g / ploidy;          // MultispeciesCoalescent.java:224
        A.cur.anR[o] = k * logPloidy;       // :211 (negated when pooled)
        s.sRate[b] = term;                  // species rates are no longer needed: reuse as the term buffer
    }
    gsync();
    if (gt == 0) {
        double logP = 0.0;
        if (!compat) logP = -INFINITY;                       // GeneTree.java:175-178
        else for (int b = 0; b < S; b++) logP += s.sRate[b];  // GeneTree.java:185-196, node-number order
        A.cur.coal[l] = logP;
    }
    stamp(10, 0);
    }

    } else {
    if (wantRate) {
        asm volatile("bar.sync 3, %0;" ::"n"(PREP_THREADS) : "memory");
    } else {
        for (int i = gt; i < G; i += gn)
            s.rate[i] = (clockKind == 2) ? A.explicitRates[nb + i] : clockRate;   // StrictClockModel / explicit
        gsync();
    }
    // ---- 4. dirtiness (TreeLikelihood.traverse) and the pruning schedule ----------------------------------------
    for (int i = gt; i < G; i += gn) {
        uint8_t md = 0;
        if (s.gP[i] >= 0) {
            const double bt = (s.gH[s.gP[i]] - s.gH[i]) * s.rate[i];   // node.getLength() * branchRate
            md = forceMats || !(bt == s.btOld[i]);
            if (md) {
                s.curM[i] = 1 - s.stM[i];                              // setNodeMatrixForUpdate
                if (owner) { A.cur.bTime[nb + i] = bt; A.cur.curM[nb + i] = s.curM[i]; }
            }
        }
        s.matDirty[i] = md;
        s.dirty[i] = 0;
        s.wk[i] = 0;
        if (owner) A.cur.bRate[nb + i] = s.rate[i];
    }
    gsync();
    // a node is recomputed when its own inputs changed or anything below it was recomputed: mark it and its ancestors
    // (the dirty set is upward closed).  Racing markers are benign: whoever marked a node first keeps climbing.
    for (int i = nT + gt; i < G; i += gn)
        if (forceAll || s.topo[i] || s.matDirty[s.gL[i]] || s.matDirty[s.gR[i]]) {
            s.dirty[i] = 1;
            int a = s.gP[i];
            while (a >= 0 && s.dirty[a] == 0) { s.dirty[a] = 2; a = s.gP[a]; }
        }
    gsync();
    // cnt[i] = number of dirty internal nodes in the subtree of i
    for (int i = nT + gt; i < G; i += gn)
        if (s.dirty[i]) {
            int a = i;
            while (a >= 0) { atomicAdd(&s.cnt[a], 1); a = s.gP[a]; }
        }
    gsync();
    stamp(11, gn);
    if (gt == 0) {
        const int n = s.dirty[root] ? s.cnt[root] : 0;
        if (owner) A.nOps[l] = n;
        if (FUSED) *F.nOpsOut = n;
    }
    // at a node whose two children are both dirty the heavier child (more dirty nodes, ties -> left) is evaluated first
    // and parks its result; everywhere else nothing is parked
    for (int i = nT + gt; i < G; i += gn) {
        const int Lc = s.gL[i], Rc = s.gR[i];
        const bool both = Lc >= nT && Rc >= nT && s.dirty[Lc] && s.dirty[Rc];
        s.first[i] = both ? ((s.cnt[Rc] > s.cnt[Lc]) ? Rc : Lc) : -1;
    }
    gsync();
    // Multi-worker walks (latency-bound shapes): worker 0 keeps the main path -- from the root always into the first-evaluated
    // (or only) dirty child -- and the small side subtrees; each larger side subtree goes, whole, to the least-loaded helper
    // that still has a free mailbox.  Helpers meet their subtrees in the order the main path needs them (deepest first).
    const int W = FUSED ? 1 : A.workers;
    if (W > 1) {
        if (gt == 0) {
            int loads[WALK_MAX_WORKERS] = {0, 0, 0, 0}, boxes[WALK_MAX_WORKERS] = {0, 0, 0, 0};
            int np = 0, nH = 0;
            for (int p = root; p >= nT && s.dirty[p];) {
                s.path[np++] = p;
                const int Lc = s.gL[p], Rc = s.gR[p];
                const bool dl = Lc >= nT && s.dirty[Lc], dr = Rc >= nT && s.dirty[Rc];
                p = (dl && dr) ? s.first[p] : (dl ? Lc : (dr ? Rc : -1));
            }
            for (int j = np - 1; j >= 0; j--) {
                const int p = s.path[j], f = s.first[p];
                if (f < 0) continue;
                const int c2 = (s.gL[p] == f) ? s.gR[p] : s.gL[p], sz = s.cnt[c2];
                if (sz < 2) continue;                       // a single op is cheaper done in place than waited for
                int best = 0;
                for (int w = 1; w < W; w++)
                    if (boxes[w] < WALK_MAILBOX && (best == 0 || loads[w] < loads[best])) best = w;
                if (best == 0) continue;
                s.wk[c2] = (uint8_t)best; s.mbx[c2] = (uint8_t)boxes[best]++; s.woff[c2] = loads[best];
                loads[best] += sz;
                s_hStart[nH] = s.cnt[f]; s_hCnt[nH] = sz; nH++;
            }
            s_nH = nH;
            int helped = 0;
            for (int w = 1; w < W; w++) helped += loads[w];
            s_lane[0] = 0; s_lane[1] = (s.dirty[root] ? s.cnt[root] : 0) - helped;
            for (int w = 1; w < W; w++) s_lane[w + 1] = s_lane[w] + loads[w];
            for (int w = W; w < WALK_MAX_WORKERS; w++) s_lane[w + 1] = s_lane[W];
            for (int w = 0; w <= WALK_MAX_WORKERS; w++) A.laneOps[l * (WALK_MAX_WORKERS + 1) + w] = s_lane[w];
        }
        gsync();
    }
    for (int i = nT + gt; i < G; i += gn) {
        if (!s.dirty[i]) continue;
        // position in the post-order and stack slot: walk to the root; every ancestor at which this path comes in through
        // the second-evaluated child adds the first child's subtree to the position and one parked result to the depth
        int start = 0, base = 0, a = i, sideRoot = -1, sideStart = 0;
        for (int p = s.gP[a]; p >= 0; a = p, p = s.gP[p]) {
            const int f = s.first[p];
            if (f >= 0 && f != a) { start += s.cnt[f]; base += 1; sideRoot = a; sideStart = s.cnt[f]; }   // the last one found hangs off the main path
        }
        int pos = start + s.cnt[i] - 1;

        if (W > 1) {
            if (sideRoot >= 0 && s.wk[sideRoot]) {   // inside a helper's subtree: contiguous in the post-order, re-based to the helper's list and stack

                pos = s_lane[s.wk[sideRoot]] + s.woff[sideRoot] + (pos - sideStart);
                base -= 1;
            } else {                                 // worker 0: the global order with the helpers' subtrees taken out
                int sub = 0;
                for (int h = 0; h < s_nH; h++) if (s_hStart[h] < pos) sub += s_hCnt[h];
                pos -= sub;
            }
        }
        s.pos[i] = pos;
        const int newP = 1 - s.stP[i];                                   // setNodePartialsForUpdate
        if (owner) A.cur.curP[nb + i] = (uint8_t)newP;
        const int Lc = s.gL[i], Rc = s.gR[i];
        const bool dl = Lc >= nT && s.dirty[Lc], dr = Rc >= nT && s.dirty[Rc];
        const int first = (dl && dr) ? s.first[i] : (dl ? Lc : Rc);
        const unsigned cp = (unsigned)C * d.patStride;   // (pattern, category) elements per node buffer (rows padded to 64 patterns)
        // the child evaluated last (the second of two dirty children, or the only dirty one) is consumed straight from
        // registers; only a first-evaluated child is parked: in stack slot `base`, or, beyond the shared-memory stack, in
        // its own global buffer
        // a second-evaluated child that a helper worker evaluates: it arrives through the helper's mailbox, and the
        // first-evaluated child is then the op right before this one in worker 0's list (register forwarding)
        const int second = (dl && dr) ? ((first == Lc) ? Rc : Lc) : -1;
        const bool remote = W > 1 && second >= 0 && s.wk[second] != 0;
        auto operand = [&](int c, bool cd, uint32_t &off, unsigned &kind, unsigned &slot) {
            slot = 0; off = 0;
            if (c < nT) { kind = OPK_TIP; off = c; }
            else if (cd) {
                if (remote && c == second) { kind = OPK_REMOTE; slot = A.stackDepth + s.mbx[c]; off = s.wk[c]; }
                else if (remote || !(dl && dr && c == first)) { kind = OPK_PREV; }
                else if (base < A.stackDepth) { kind = OPK_STACK; slot = base; }
                else { kind = OPK_SPILL; off = ((1 - s.stP[c]) * nI + (c - nT)) * cp; }   // the buffer the child is being written to
            }
            else { kind = OPK_GLOBAL; off = (s.curP[c] * nI + (c - nT)) * cp; }
        };
        Op o;
        unsigned aKind, bKind, aSlot, bSlot;
        o.outOff = (newP * nI + (i - nT)) * cp;
        operand(Lc, dl, o.aOff, aKind, aSlot);
        operand(Rc, dr, o.bOff, bKind, bSlot);
        // is this node itself parked?  yes iff it is the first-evaluated child of a parent with two dirty children
        // (then its slot is the parent's base == its own base); the root parks in slot 0 for the category combine
        unsigned outSlot = OP_NO_SLOT;
        const int par = s.gP[i];
        unsigned signal = 0;
        if (par < 0) outSlot = 0;
        else if (W > 1 && s.wk[i]) { outSlot = A.stackDepth + s.mbx[i]; signal = OP_SIGNAL; }   // a helper's finished subtree
        else if (s.first[par] == i && base < A.stackDepth) {
            const int sib = (s.gL[par] == i) ? s.gR[par] : s.gL[par];
            if (!(W > 1 && s.wk[sib])) outSlot = base;     // sibling evaluated elsewhere: the parent takes this result from registers
        }
        o.ctl = op_ctl(aKind, bKind, aSlot, bSlot, outSlot) | signal;

        if (FUSED) F.sops[pos] = o;
        else A.ops[d.opBase + pos] = o;
    }
    gsync();
    stamp(12, gn);

    }
    // ---- 5. transition matrices: recompute the dirty branches; stream every scheduled op's pair in op order ----------------
    // few matrices: group B computes them alone, without waiting for group A; many: all threads, after a full barrier
    const bool matsByB = (!FUSED || F.splitReturn) && G * C <= gn;
    if (matsByB) { if (grpId == 0) return PREP_HALVES; }
    else __syncthreads();
    const int mt = matsByB ? gt : tid, mn = matsByB ? gn : nt;
    HkyTabs tabs;
    const int substKind = prm->substKind;
    if (substKind == 0) hky_setup(prm->subst, tabs);
    for (int idx = mt; idx < G * C; idx += mn) {
        const int i = idx / C, c = idx - i * C;
        const int par = s.gP[i];
        if (par < 0) continue;
        const bool md = s.matDirty[i], used = s.dirty[par];   // used: the parent's op reads this matrix
        if (!md && !used) continue;
        double *store = A.mats + d.matBase + ((size_t)(s.curM[i] * G + i) * C + c) * 16;
        double m[16];
        if (md) {
            const double jointBranchRate = A.catRates[d.catBase + c] * s.rate[i];   // siteModel.getRateForCategory * branchRate
            const double distance = (s.gH[par] - s.gH[i]) * jointBranchRate;
            if (substKind == 0) hky_matrix(tabs, distance, m);
            else eigen_matrix(prm->subst, distance, m);
            if (owner) {   // 256-bit stores: whole 32-byte sectors per lane (half as many store instructions as with 128-bit ones)
#pragma unroll
                for (int q = 0; q < 16; q += 4) prep_st256(store + q, m[q], m[q + 1], m[q + 2], m[q + 3]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; q += 2) { const double2 v = *reinterpret_cast<const double2 *>(store + q); m[q] = v.x; m[q + 1] = v.y; }
        }
        if (used) {   // [op][category][side][16]: the pruning kernel's warp (one category) reads 256 contiguous bytes per op
            const int side = (s.gR[par] == i) ? 1 : 0;
            double *dst = FUSED ? F.smats + (((size_t)s.pos[par] * C + c) * 2 + side) * 16
                                : A.opMats + (((size_t)(d.opBase + s.pos[par]) * C + c) * 2 + side) * 16;
            if (FUSED) {   // shared memory
#pragma unroll
                for (int q = 0; q < 16; q += 2) *reinterpret_cast<double2 *>(dst + q) = make_double2(m[q], m[q + 1]);
            } else {
#pragma unroll
                for (int q = 0; q < 16; q += 4) prep_st256(dst + q, m[q], m[q + 1], m[q + 2], m[q + 3]);
            }
        }
    }
    if (FUSED && matsByB) {   // this half is done: let the coalescent half know the schedule and the matrices are complete
        gsync();
        stamp(13, gn);
        asm volatile("bar.arrive 4, %0;" ::"n"(PREP_THREADS) : "memory");
        return PREP_HALVES;
    }
    if (FUSED) __syncthreads();   // schedule and matrices are complete in this CTA's shared memory
    stamp(13, 0);
    return PREP_ALL;
}

}  // namespace sb2
