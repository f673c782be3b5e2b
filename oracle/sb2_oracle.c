/*
 * sb2_oracle.c -- CPU restatement of the StarBEAST2 hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / reported CPU baseline.
 *
 * Two halves, pinned differently:
 *
 *  (1) Reference-owned half (genomescale/starbeast2, src/starbeast2/*.java).  Every
 *      function cites the Java file:line it follows loop-for-loop.  PINNED by the
 *      reference's own known-answer tests (src/sb2tests/*Test.java), reproduced in
 *      tests/test_oracle_golden.py.
 *
 *  (2) BEAST.base half (TreeLikelihood, BeerLikelihoodCore4, HKY, GeneralSubstitutionModel,
 *      SiteModel; CompEvol/beast2 >= 2.7.3 per /root/reference/version.xml:2).  That source
 *      is NOT vendored under /root/reference and the reference has no test that
 *      constructs a TreeLikelihood, so this half restates the published algorithm
 *       and is "PARITY UNPINNED": it is checked against
 *      brute-force state enumeration, scipy.linalg.expm and analytic JC69 limits instead.
 *
 * Numerics: plain IEEE fp64, compile with -O2 -ffp-contract=off so that, like the JVM,
 * no multiply-add is fused and sums associate exactly as written.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SB2O_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* (1a) GeneTree.update() + collateCoalescenceEvents()   GeneTree.java:202-318, 332-379        */
/* ------------------------------------------------------------------------------------------ */

/* GeneTree.java:332-379 -- walk one lineage rootward until it meets an already visited node. */
static int collate_coalescence_events(
    int lastGeneNode, double lastHeight, int geneNode, int speciesNode,
    const int32_t *gParent, const double *gHeight,
    const int32_t *sParent, const double *sHeight, int S, int blocksize,
    int32_t *lineageCounts, int32_t *coalCounts, double *coalTimes,
    int32_t *assignment, double *occupancy)
{
    for (;;) {
        const double geneNodeHeight = gHeight[geneNode];
        /* :337  climb while the coalescence is at or above the parent species node (note >=) */
        while (sParent[speciesNode] >= 0 && geneNodeHeight >= sHeight[sParent[speciesNode]]) {
            const int sp = sParent[speciesNode];
            const double spHeight = sHeight[sp];
            if (occupancy) occupancy[(size_t)lastGeneNode * S + speciesNode] = spHeight - lastHeight; /* :345 */
            lineageCounts[sp]++;                                                                      /* :346 */
            speciesNode = sp;
            lastHeight = spHeight;
        }
        if (occupancy) occupancy[(size_t)lastGeneNode * S + speciesNode] = geneNodeHeight - lastHeight; /* :354 */
        const int existing = assignment[geneNode];
        if (existing == -1) {                                                                         /* :356 */
            assignment[geneNode] = speciesNode;
            coalTimes[(size_t)speciesNode * blocksize + coalCounts[speciesNode]++] = geneNodeHeight;  /* :359 */
            const int next = gParent[geneNode];
            if (next < 0) return 1;                                                                   /* :362 root */
            lastGeneNode = geneNode;
            lastHeight = geneNodeHeight;
            geneNode = next;
        } else if (existing == speciesNode) {
            return 1;                                                                                 /* :373 */
        } else {
            return 0;                                                                                 /* :375 incompatible */
        }
    }
}

/*
 * GeneTree.update() -- GeneTree.java:202-318.  The Java grows `blocksize` on demand (Q7 in
 * SURVEY.md); the caller here passes a capacity >= nTips-1 so no regrowth is needed.  The
 * G x S occupancy matrix is zero-filled every call exactly like :250 (it is part of the
 * cost the reference pays per step).  Returns 1 if the gene tree fits the species tree.
 */
SB2O_API int sb2o_gene_tree_update(
    int G, int nTips, const int32_t *gParent, const double *gHeight, const int32_t *tipSpecies,
    int S, int nSpeciesLeaves, const int32_t *sParent, const double *sHeight, int blocksize,
    int32_t *lineageCounts, int32_t *coalCounts, double *coalTimes, int32_t *assignment, double *occupancy)
{
    /* :243-252 reset */
    for (int i = 0; i < S; i++) { lineageCounts[i] = 0; coalCounts[i] = 0; }
    for (int i = 0; i < nTips; i++) { assignment[i] = tipSpecies[i]; lineageCounts[tipSpecies[i]]++; } /* :214-223,243-244 */
    for (int i = nTips; i < G; i++) assignment[i] = -1;                                               /* :247 */
    (void)nSpeciesLeaves;
    if (occupancy) memset(occupancy, 0, sizeof(double) * (size_t)G * S);                              /* :250 */

    for (int leaf = 0; leaf < nTips; leaf++) {                                                        /* :254 */
        const int speciesLeaf = tipSpecies[leaf];
        const int first = gParent[leaf];
        if (!collate_coalescence_events(leaf, 0.0 /* :260 */, first, speciesLeaf, gParent, gHeight,
                                        sParent, sHeight, S, blocksize, lineageCounts, coalCounts,
                                        coalTimes, assignment, occupancy))
            return 0;                                                                                 /* :265-268 */
    }
    return 1;
}

static int cmp_double(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

/* GeneTree.getCoalescentTimes(nodeI) -- GeneTree.java:387-405.  out has k+2 entries. */
SB2O_API void sb2o_coalescent_times(
    int nodeI, const int32_t *sParent, const double *sHeight, const int32_t *coalCounts,
    const double *coalTimes, int blocksize, double *out)
{
    const int k = coalCounts[nodeI];
    out[0] = sHeight[nodeI];
    out[k + 1] = (sParent[nodeI] < 0) ? INFINITY : sHeight[sParent[nodeI]];                           /* :394 */
    memcpy(out + 1, coalTimes + (size_t)nodeI * blocksize, sizeof(double) * (size_t)k);              /* :401 */
    qsort(out, (size_t)k + 2, sizeof(double), cmp_double);                                            /* :402 */
}

/* ------------------------------------------------------------------------------------------ */
/* (1b) population models                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* shared gamma loop: ConstantPopulations.java:96-103 == MultispeciesCoalescent.java:215-222 */
static double partial_gamma(const double *t, int n, int k)
{
    double g = 0.0;
    for (int i = 0; i < k; i++)
        g += (t[i + 1] - t[i]) * (n - i) * (n - (i + 1.0)) / 2.0;
    if (n - k > 1)
        g += (t[k + 1] - t[k]) * (n - k) * (n - (k + 1.0)) / 2.0;
    return g;
}

/* ConstantPopulations.constantLogP -- ConstantPopulations.java:95-118 */
SB2O_API double sb2o_constant_logp(double popSize, double ploidy, const double *t, int n, int k)
{
    const double partialGamma = partial_gamma(t, n, k);
    const double branchGamma = partialGamma / ploidy;
    const double branchLogR = -k * log(ploidy);
    return branchLogR - (k * log(popSize)) - (branchGamma / popSize);                                 /* :107 */
}

/* LinearWithConstantRoot.linearLogP -- LinearWithConstantRoot.java:155-201 */
SB2O_API double sb2o_linear_logp(double topPopSize, double tipPopSize, double ploidy, const double *t, int n, int k)
{
    const double fPopSizeTop = topPopSize * ploidy;
    const double fPopSizeBottom = tipPopSize * ploidy;
    const double d5 = fPopSizeTop - fPopSizeBottom;
    const double fTime0 = t[0];
    const double a = d5 / (t[k + 1] - fTime0);
    const double b = fPopSizeBottom;
    double logP = 0.0;
    if (fabs(d5) < 1e-10) {                                                                           /* :165 */
        for (int i = 0; i <= k; i++) {
            const double fTimeip1 = t[i + 1];
            const double fPopSize = a * (fTimeip1 - fTime0) + b;
            if (i < k) logP += -log(fPopSize);
            const int i1 = n - i;
            logP -= (i1 * (i1 - 1.0) / 2.0) * (fTimeip1 - t[i]) / fPopSize;
        }
    } else {                                                                                          /* :178 */
        const double vv = b - a * fTime0;
        for (int i = 0; i <= k; i++) {
            const double fPopSize = a * t[i + 1] + vv;
            if (i < k) logP += -log(fPopSize);
            const double f = fPopSize / (a * t[i] + vv);
            const int i1 = n - i;
            logP += -(i1 * (i1 - 1.0) / 2.0) * log(f) / a;
        }
    }
    return logP;
}

/*
 * LinearWithConstantRoot.branchLogP -- LinearWithConstantRoot.java:50-70.
 * tipPop[leaf], topPop[node<root] indexed by species node number (root excluded, Q8).
 */
SB2O_API double sb2o_lwcr_branch_logp(
    int nodeI, int nSpeciesLeaves, const int32_t *sParent, const int32_t *sLeft, const int32_t *sRight,
    const double *tipPop, const double *topPop, double ploidy, const double *t, int n, int k)
{
    double branchTip;
    if (nodeI < nSpeciesLeaves && sLeft[nodeI] < 0) branchTip = tipPop[nodeI];                        /* :55 */
    else branchTip = topPop[sLeft[nodeI]] + topPop[sRight[nodeI]];                                    /* :58-60 */
    if (sParent[nodeI] < 0) return sb2o_constant_logp(branchTip, ploidy, t, n, k);                    /* :63-64 */
    return sb2o_linear_logp(topPop[nodeI], branchTip, ploidy, t, n, k);                               /* :66-68 */
}

/* MultispeciesCoalescent.analyticalLogP -- MultispeciesCoalescent.java:200-235.
 * times: concatenated per-gene arrays, timesOff[j] = start of gene j's k_j+2 entries. */
SB2O_API double sb2o_analytical_logp(
    double alpha, double beta, int nGenes, const double *perGenePloidy,
    const double *times, const int64_t *timesOff, const int32_t *n, const int32_t *k)
{
    int branchQ = 0;
    double branchLogR = 0.0, branchGamma = 0.0;
    for (int j = 0; j < nGenes; j++) {
        const double pl = perGenePloidy[j];
        branchLogR -= k[j] * log(pl);
        branchQ += k[j];
        branchGamma += partial_gamma(times + timesOff[j], n[j], k[j]) / pl;
    }
    double logGammaRatio = 0.0;
    for (int i = 0; i < branchQ; i++) logGammaRatio += log(alpha + i);
    return branchLogR + (alpha * log(beta)) - ((alpha + branchQ) * log(beta + branchGamma)) + logGammaRatio;
}

/* ------------------------------------------------------------------------------------------ */
/* (1c) clocks                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* StarBeastClock.update -- StarBeastClock.java:54-75 */
SB2O_API void sb2o_starbeast_clock(int G, int S, double geneTreeRate, const double *speciesRates,
                                   const double *occupancy, double *branchRates)
{
    for (int i = 0; i < G - 1; i++) {
        double weightedSum = 0.0, branchLength = 0.0;
        for (int j = 0; j < S; j++) {
            weightedSum += speciesRates[j] * occupancy[(size_t)i * S + j];
            branchLength += occupancy[(size_t)i * S + j];
        }
        branchRates[i] = geneTreeRate * weightedSum / branchLength;
    }
    branchRates[G - 1] = geneTreeRate;                                                                /* :72 */
}

/*
 * Standard-normal quantile.  The reference calls commons-math 2.x
 * NormalDistributionImpl.inverseCumulativeProbability (UncorrelatedRates.java:136-141), an
 * iterative bracketing solver on erf with ~1e-9 absolute accuracy that is not part of
 * /root/reference (quirk Q5).  The oracle uses the exact quantile instead (Acklam's rational
 * start + two Halley steps on erfc), which agrees with the reference's golden vectors
 * (UncorrelatedRatesTest.java:107-117, StarbeastClockTest.java:106-110) to their 7 printed digits.
 */
static double std_normal_quantile(double p)
{
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    double x;
    if (p < 0.02425) {
        const double q = sqrt(-2 * log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (p <= 1 - 0.02425) {
        const double q = p - 0.5, r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        const double q = sqrt(-2 * log(1 - p));
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    for (int it = 0; it < 2; it++) {
        const double e = 0.5 * erfc(-x / sqrt(2.0)) - p;
        const double u = e * sqrt(2.0 * M_PI) * exp(x * x / 2.0);
        x = x - u / (1.0 + x * u / 2.0);
    }
    return x;
}

/* UncorrelatedRates.update, log-normal bins -- UncorrelatedRates.java:132-145 */
SB2O_API void sb2o_ucln_bin_rates(int nBins, double stdev, double *binRates)
{
    const double newMean = -(0.5 * stdev * stdev);
    for (int i = 0; i < nBins; i++)
        binRates[i] = exp(newMean + stdev * std_normal_quantile((i + 0.5) / nBins));
}

/* UncorrelatedRates init, exponential bins (UCED) -- UncorrelatedRates.java:110-117 */
SB2O_API void sb2o_uced_bin_rates(int nBins, double *binRates)
{
    for (int i = 0; i < nBins; i++) binRates[i] = -log(1.0 - (i + 0.5) / nBins);
}

/* UncorrelatedRates.update, rate assignment -- UncorrelatedRates.java:147-161 */
SB2O_API void sb2o_uncorrelated_rates(int S, int nEstimated, const int32_t *binIndex, const double *binRates,
                                      double estimatedMean, int estimateRoot, int rootNode, double *ratesArray)
{
    for (int i = 0; i < nEstimated; i++) ratesArray[i] = binRates[binIndex[i]] * estimatedMean;
    if (!estimateRoot) ratesArray[rootNode] = estimatedMean;
    (void)S;
}

/* RandomLocalRates.recurseBranchRates -- RandomLocalRates.java:72-90 */
static void rlr_recurse(int node, double parentHeight, double rate, int rootNodeNumber, const int32_t *left, const int32_t *right,
                        const double *height, const uint8_t *indicators, const double *branchRates, double *ratesArray,
                        double *strictTotal, double *relaxedTotal)
{
    const double nodeHeight = height[node];
    if (node < rootNodeNumber && indicators[node]) rate = branchRates[node];
    const double branchLength = parentHeight - nodeHeight;
    *strictTotal += branchLength;
    *relaxedTotal += branchLength * rate;
    ratesArray[node] = rate;
    if (left[node] >= 0) {
        rlr_recurse(left[node], nodeHeight, rate, rootNodeNumber, left, right, height, indicators, branchRates, ratesArray, strictTotal, relaxedTotal);
        rlr_recurse(right[node], nodeHeight, rate, rootNodeNumber, left, right, height, indicators, branchRates, ratesArray, strictTotal, relaxedTotal);
    }
}

/* RandomLocalRates.update -- RandomLocalRates.java:92-117 (no reference test: PARITY UNPINNED).  root = node S-1 (:54). */
SB2O_API void sb2o_random_local_rates(int S, int root, const int32_t *left, const int32_t *right, const double *height,
                                      const uint8_t *indicators, const double *branchRates, double estimatedMean, double *ratesArray)
{
    const int rootNodeNumber = S - 1;
    double strictTotal = 0.0, relaxedTotal = 0.0;
    ratesArray[rootNodeNumber] = estimatedMean;
    rlr_recurse(root, height[root], 1.0, rootNodeNumber, left, right, height, indicators, branchRates, ratesArray, &strictTotal, &relaxedTotal);
    const double scaleFactor = estimatedMean * strictTotal / relaxedTotal;
    for (int i = 0; i < rootNodeNumber; i++) ratesArray[i] = ratesArray[i] * scaleFactor;
}

/* ------------------------------------------------------------------------------------------ */
/* (2) BEAST.base half -- PARITY UNPINNED (see header)                                         */
/* ------------------------------------------------------------------------------------------ */

/* HKY.getTransitionProbabilities + setupMatrix (closed form; state order A,C,G,T; [from][to]). */
SB2O_API void sb2o_hky_matrix(double kappa, const double *pi, double distance, double *matrix)
{
    const double freqA = pi[0], freqC = pi[1], freqG = pi[2], freqT = pi[3];
    const double freqR = freqA + freqG, freqY = freqC + freqT;
    const double r1 = (1 / freqR) - 1;
    const double tab1A = freqA * r1;
    const double tab3A = freqA / freqR;
    const double tab2A = 1 - tab3A;
    const double r2 = 1 / r1;
    const double tab1C = freqC * r2;
    const double tab3C = freqC / freqY;
    const double tab2C = 1 - tab3C;
    const double tab1G = freqG * r1;
    const double tab3G = tab2A;
    const double tab2G = tab3A;
    const double tab1T = freqT * r2;
    const double tab3T = tab2C;
    const double tab2T = tab3C;
    const double beta = -1.0 / (2.0 * (freqR * freqY + kappa * (freqA * freqG + freqC * freqT)));
    const double A_R = 1.0 + freqR * (kappa - 1);
    const double A_Y = 1.0 + freqY * (kappa - 1);

    const double xx = beta * distance;
    const double bbR = exp(xx * A_R);
    const double bbY = exp(xx * A_Y);
    const double aa = exp(xx);
    const double oneminusa = 1 - aa;
    const double t1Aaa = tab1A * aa;
    matrix[0] = freqA + t1Aaa + (tab2A * bbR);
    matrix[1] = freqC * oneminusa;
    const double t1Gaa = tab1G * aa;
    matrix[2] = freqG + t1Gaa - (tab3G * bbR);
    matrix[3] = freqT * oneminusa;
    matrix[4] = freqA * oneminusa;
    const double t1Caa = tab1C * aa;
    matrix[5] = freqC + t1Caa + (tab2C * bbY);
    matrix[6] = freqG * oneminusa;
    const double t1Taa = tab1T * aa;
    matrix[7] = freqT + t1Taa - (tab3T * bbY);
    matrix[8] = freqA + t1Aaa - (tab3A * bbR);
    matrix[9] = matrix[1];
    matrix[10] = freqG + t1Gaa + (tab2G * bbR);
    matrix[11] = matrix[3];
    matrix[12] = matrix[4];
    matrix[13] = freqC + t1Caa - (tab3C * bbY);
    matrix[14] = matrix[6];
    matrix[15] = freqT + t1Taa + (tab2T * bbY);
}

/* GeneralSubstitutionModel.getTransitionProbabilities: P = |V exp(d L) V^-1| (note the abs). */
SB2O_API void sb2o_eigen_matrix(const double *Eval, const double *Evec, const double *Ievc, double distance, double *matrix)
{
    double iexp[16];
    for (int i = 0; i < 4; i++) {
        const double temp = exp(distance * Eval[i]);
        for (int j = 0; j < 4; j++) iexp[i * 4 + j] = Ievc[i * 4 + j] * temp;
    }
    int u = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double temp = 0.0;
            for (int k = 0; k < 4; k++) temp += Evec[i * 4 + k] * iexp[k * 4 + j];
            matrix[u++] = fabs(temp);
        }
}

enum { SB2O_SUBST_HKY = 0, SB2O_SUBST_EIGEN = 1 };
/* substParams: HKY {kappa, pi[4]}; EIGEN {eval[4], evec[16], ievc[16], pi[4]} */
#define SB2O_SUBST_STRIDE 40

static void transition_matrix(int kind, const double *sp, double distance, double *m)
{
    if (kind == SB2O_SUBST_HKY) sb2o_hky_matrix(sp[0], sp + 1, distance, m);
    else sb2o_eigen_matrix(sp, sp + 4, sp + 20, distance, m);
}
static const double *subst_freqs(int kind, const double *sp) { return kind == SB2O_SUBST_HKY ? sp + 1 : sp + 36; }

/* BeerLikelihoodCore4.calculateStatesStatesPruning */
static void states_states(const uint8_t *s1, const double *m1, const uint8_t *s2, const double *m2,
                          double *p3, int nPat, int nCat)
{
    int v = 0;
    for (int l = 0; l < nCat; l++)
        for (int k = 0; k < nPat; k++) {
            const int a = s1[k], b = s2[k];
            int w = l * 16;
            if (a < 4 && b < 4) {
                for (int i = 0; i < 4; i++, w += 4) p3[v++] = m1[w + a] * m2[w + b];
            } else if (a < 4) {
                for (int i = 0; i < 4; i++, w += 4) p3[v++] = m1[w + a];
            } else if (b < 4) {
                for (int i = 0; i < 4; i++, w += 4) p3[v++] = m2[w + b];
            } else {
                for (int i = 0; i < 4; i++) p3[v++] = 1.0;
            }
        }
}

/* BeerLikelihoodCore4.calculateStatesPartialsPruning */
static void states_partials(const uint8_t *s1, const double *m1, const double *p2, const double *m2,
                            double *p3, int nPat, int nCat)
{
    int u = 0, v = 0;
    for (int l = 0; l < nCat; l++)
        for (int k = 0; k < nPat; k++) {
            const int a = s1[k];
            int w = l * 16;
            for (int i = 0; i < 4; i++, w += 4) {
                const double sum = m2[w] * p2[v] + m2[w + 1] * p2[v + 1] + m2[w + 2] * p2[v + 2] + m2[w + 3] * p2[v + 3];
                p3[u++] = (a < 4) ? m1[w + a] * sum : sum;
            }
            v += 4;
        }
}

/* BeerLikelihoodCore4.calculatePartialsPartialsPruning */
static void partials_partials(const double *p1, const double *m1, const double *p2, const double *m2,
                              double *p3, int nPat, int nCat)
{
    int u = 0, v = 0;
    for (int l = 0; l < nCat; l++)
        for (int k = 0; k < nPat; k++) {
            int w = l * 16;
            for (int i = 0; i < 4; i++, w += 4) {
                const double sum1 = m1[w] * p1[v] + m1[w + 1] * p1[v + 1] + m1[w + 2] * p1[v + 2] + m1[w + 3] * p1[v + 3];
                const double sum2 = m2[w] * p2[v] + m2[w + 1] * p2[v + 1] + m2[w + 2] * p2[v + 2] + m2[w + 3] * p2[v + 3];
                p3[u++] = sum1 * sum2;
            }
            v += 4;
        }
}

/* BeerLikelihoodCore.scalePartials (threshold 1e-100; log of the per-pattern max) */
static void scale_partials(double *p, double *scal, int nPat, int nCat)
{
    for (int i = 0; i < nPat; i++) {
        double scaleFactor = 0.0;
        for (int k = 0; k < nCat; k++)
            for (int j = 0; j < 4; j++) {
                const double x = p[((size_t)k * nPat + i) * 4 + j];
                if (x > scaleFactor) scaleFactor = x;
            }
        if (scaleFactor < 1.0E-100) {
            for (int k = 0; k < nCat; k++)
                for (int j = 0; j < 4; j++) p[((size_t)k * nPat + i) * 4 + j] /= scaleFactor;
            scal[i] = log(scaleFactor);
        } else {
            scal[i] = 0.0;
        }
    }
}

typedef struct {
    int G, nTips, nPat, nCat;
    const int32_t *left, *right, *parent;
    const double *height;
    const uint8_t *tipStates; /* [tip][pattern] */
    int substKind;
    const double *substParams;
    const double *catRates; /* already x mutationRate (SiteModel.getRateForCategory) */
    const double *branchRates;
    double *partials; /* [G - nTips][nCat][nPat][4] */
    double *matrices; /* [G][nCat][16] */
    double *scaling;  /* [G - nTips][nPat] or NULL */
} tl_ctx;

/* TreeLikelihood.traverse (everything dirty): matrices for the branch above `node`, children, then node. */
static void tl_traverse(const tl_ctx *c, int node)
{
    const int par = c->parent[node];
    if (par >= 0) {
        const double branchRate = c->branchRates[node];
        for (int i = 0; i < c->nCat; i++) {
            const double jointBranchRate = c->catRates[i] * branchRate;
            const double distance = (c->height[par] - c->height[node]) * jointBranchRate;
            transition_matrix(c->substKind, c->substParams, distance, c->matrices + ((size_t)node * c->nCat + i) * 16);
        }
    }
    if (node >= c->nTips) {
        const int c1 = c->left[node], c2 = c->right[node];
        tl_traverse(c, c1);
        tl_traverse(c, c2);
        const size_t psz = (size_t)c->nCat * c->nPat * 4;
        double *p3 = c->partials + (size_t)(node - c->nTips) * psz;
        const double *m1 = c->matrices + (size_t)c1 * c->nCat * 16, *m2 = c->matrices + (size_t)c2 * c->nCat * 16;
        /* BeerLikelihoodCore.calculatePartials dispatch */
        if (c1 < c->nTips) {
            if (c2 < c->nTips)
                states_states(c->tipStates + (size_t)c1 * c->nPat, m1, c->tipStates + (size_t)c2 * c->nPat, m2, p3, c->nPat, c->nCat);
            else
                states_partials(c->tipStates + (size_t)c1 * c->nPat, m1, c->partials + (size_t)(c2 - c->nTips) * psz, m2, p3, c->nPat, c->nCat);
        } else {
            if (c2 < c->nTips)
                states_partials(c->tipStates + (size_t)c2 * c->nPat, m2, c->partials + (size_t)(c1 - c->nTips) * psz, m1, p3, c->nPat, c->nCat);
            else
                partials_partials(c->partials + (size_t)(c1 - c->nTips) * psz, m1, c->partials + (size_t)(c2 - c->nTips) * psz, m2, p3, c->nPat, c->nCat);
        }
        if (c->scaling) scale_partials(p3, c->scaling + (size_t)(node - c->nTips) * c->nPat, c->nPat, c->nCat);
    }
}

/*
 * TreeLikelihood.calculateLogP, full evaluation.
 * useScaling: 0 never, 1 always (threshold scaling as BeerLikelihoodCore.scalePartials),
 *             2 BEAST's lazy rule: evaluate unscaled, re-evaluate scaled if the result is -inf.
 * constMask[p]: bit s set when pattern p is constant in state s (TreeLikelihood.calcConstantPatternIndices);
 *             only used when pInv > 0.
 * Optional outputs (may be NULL): patternLogL[nPat], partialsOut[(G-nTips)*nCat*nPat*4],
 *             matricesOut[G*nCat*16], scalingOut[(G-nTips)*nPat] (log scale factors; zeros if unscaled).
 */
SB2O_API double sb2o_tree_likelihood(
    int G, int nTips, int nPat, int nCat,
    const int32_t *parent, const int32_t *left, const int32_t *right, const double *height,
    const uint8_t *tipStates, const int32_t *weights,
    int substKind, const double *substParams,
    const double *catRates, const double *catProps, double pInv, const uint8_t *constMask,
    const double *branchRates, int useScaling,
    double *patternLogL, double *partialsOut, double *matricesOut, double *scalingOut)
{
    const int nInt = G - nTips;
    const size_t psz = (size_t)nCat * nPat * 4;
    double *partials = partialsOut ? partialsOut : (double *)malloc(sizeof(double) * psz * nInt);
    double *matrices = matricesOut ? matricesOut : (double *)malloc(sizeof(double) * (size_t)G * nCat * 16);
    double *scaling = (double *)calloc((size_t)nInt * nPat, sizeof(double));
    double *rootPartials = (double *)malloc(sizeof(double) * (size_t)nPat * 4);
    int root = G - 1;
    for (int i = 0; i < G; i++) if (parent[i] < 0) root = i;
    const double *freqs = subst_freqs(substKind, substParams);
    double logP = 0.0;

    for (int pass = 0; pass < 2; pass++) {
        const int scaled = (useScaling == 1) || (useScaling == 2 && pass == 1);
        tl_ctx c = {G, nTips, nPat, nCat, left, right, parent, height, tipStates, substKind, substParams,
                    catRates, branchRates, partials, matrices, scaled ? scaling : NULL};
        if (root < nTips) { logP = NAN; break; }
        for (int i = 0; i < nCat * 16; i++) matrices[(size_t)root * nCat * 16 + i] = 0.0;
        tl_traverse(&c, root);

        /* BeerLikelihoodCore.integratePartials */
        const double *rp = partials + (size_t)(root - nTips) * psz;
        for (int k = 0; k < nPat; k++)
            for (int i = 0; i < 4; i++) rootPartials[k * 4 + i] = rp[k * 4 + i] * catProps[0];
        for (int l = 1; l < nCat; l++)
            for (int k = 0; k < nPat; k++)
                for (int i = 0; i < 4; i++) rootPartials[k * 4 + i] += rp[((size_t)l * nPat + k) * 4 + i] * catProps[l];
        /* TreeLikelihood: proportion invariant on constant patterns */
        if (pInv > 0.0 && constMask)
            for (int k = 0; k < nPat; k++)
                for (int i = 0; i < 4; i++)
                    if (constMask[k] & (1 << i)) rootPartials[k * 4 + i] += pInv;
        /* calculateLogLikelihoods + getLogScalingFactor + calcLogP */
        logP = 0.0;
        for (int k = 0; k < nPat; k++) {
            double sum = 0.0;
            for (int i = 0; i < 4; i++) sum += freqs[i] * rootPartials[k * 4 + i];
            double logScale = 0.0;
            if (scaled)
                for (int nd = 0; nd < nInt; nd++) logScale += scaling[(size_t)nd * nPat + k];
            const double lk = log(sum) + logScale;
            if (patternLogL) patternLogL[k] = lk;
            logP += lk * weights[k];
        }
        if (useScaling != 2 || pass == 1 || !(logP == -INFINITY)) break;
    }
    if (scalingOut) memcpy(scalingOut, scaling, sizeof(double) * (size_t)nInt * nPat);
    if (!partialsOut) free(partials);
    if (!matricesOut) free(matrices);
    free(scaling);
    free(rootPartials);
    return logP;
}

/* ------------------------------------------------------------------------------------------ */
/* (1c') K-state likelihoods: the morphological TreeLikelihoods of the FBD    */
/*      configs -- /root/reference/tutorial/CanisFBD/Canis-FBD.xml:964-1033: TreeLikelihood on    */
/*      tree="@Tree.t:Species", StandardData with nrOfStates 2..5, LewisMK, ascertained           */
/*      FilteredAlignment.  The classes are BEAST.base's generic BeerLikelihoodCore /             */
/*      GeneralSubstitutionModel / Alignment.getAscertainmentCorrection and the morph-models       */
/*      package's LewisMK (neither under /root/reference): restated from the published            */
/*      algorithm, the loops of the 4-state versions above with nrOfStates in place of 4.          */
/*      PARITY UNPINNED like the rest of the BEAST.base half; pinned by brute force over internal  */
/*      states and by the Mk closed form (tests/test_oracle_kstate.py).                            */
/* ------------------------------------------------------------------------------------------ */

/* GeneralSubstitutionModel.getTransitionProbabilities for K states: P = |V exp(d L) V^-1| */
SB2O_API void sb2o_eigen_matrix_k(int K, const double *Eval, const double *Evec, const double *Ievc, double distance, double *matrix)
{
    double iexp[64];
    for (int i = 0; i < K; i++) {
        const double temp = exp(distance * Eval[i]);
        for (int j = 0; j < K; j++) iexp[i * K + j] = Ievc[i * K + j] * temp;
    }
    int u = 0;
    for (int i = 0; i < K; i++)
        for (int j = 0; j < K; j++) {
            double temp = 0.0;
            for (int k = 0; k < K; k++) temp += Evec[i * K + k] * iexp[k * K + j];
            matrix[u++] = fabs(temp);
        }
}

typedef struct {
    int G, nTips, nPat, nCat, K;
    const int32_t *left, *right, *parent;
    const double *height;
    const uint8_t *tipStates;
    const double *Eval, *Evec, *Ievc, *catRates, *branchRates;
    double *partials; /* [G - nTips][nCat][nPat][K] */
    double *matrices; /* [G][nCat][K*K] */
    double *scaling;  /* [G - nTips][nPat] or NULL */
} tlk_ctx;

/* BeerLikelihoodCore.calculate{StatesStates,StatesPartials,PartialsPartials}Pruning + scalePartials, generic state count;
 * a tip state >= K (gap / '?' / ambiguity with useAmbiguities=false) contributes the factor 1 */
static void tlk_traverse(const tlk_ctx *c, int node)
{
    const int K = c->K, KK = K * K;
    const int par = c->parent[node];
    if (par >= 0)
        for (int i = 0; i < c->nCat; i++) {
            const double jointBranchRate = c->catRates[i] * c->branchRates[node];
            sb2o_eigen_matrix_k(K, c->Eval, c->Evec, c->Ievc, (c->height[par] - c->height[node]) * jointBranchRate,
                                c->matrices + ((size_t)node * c->nCat + i) * KK);
        }
    if (node < c->nTips) return;
    const int c1 = c->left[node], c2 = c->right[node];
    tlk_traverse(c, c1);
    tlk_traverse(c, c2);
    const size_t psz = (size_t)c->nCat * c->nPat * K;
    double *p3 = c->partials + (size_t)(node - c->nTips) * psz;
    const double *m1 = c->matrices + (size_t)c1 * c->nCat * KK, *m2 = c->matrices + (size_t)c2 * c->nCat * KK;
    const uint8_t *s1 = c1 < c->nTips ? c->tipStates + (size_t)c1 * c->nPat : NULL;
    const uint8_t *s2 = c2 < c->nTips ? c->tipStates + (size_t)c2 * c->nPat : NULL;
    const double *p1 = s1 ? NULL : c->partials + (size_t)(c1 - c->nTips) * psz;
    const double *p2 = s2 ? NULL : c->partials + (size_t)(c2 - c->nTips) * psz;
    size_t u = 0, v = 0;
    for (int l = 0; l < c->nCat; l++)
        for (int k = 0; k < c->nPat; k++) {
            int w = l * KK;
            for (int i = 0; i < K; i++, w += K) {
                double f1, f2;
                if (s1) f1 = s1[k] < K ? m1[w + s1[k]] : 1.0;
                else { f1 = 0.0; for (int j = 0; j < K; j++) f1 += m1[w + j] * p1[v + j]; }
                if (s2) f2 = s2[k] < K ? m2[w + s2[k]] : 1.0;
                else { f2 = 0.0; for (int j = 0; j < K; j++) f2 += m2[w + j] * p2[v + j]; }
                p3[u++] = f1 * f2;
            }
            v += K;
        }
    if (c->scaling) {
        double *scal = c->scaling + (size_t)(node - c->nTips) * c->nPat;
        for (int i = 0; i < c->nPat; i++) {
            double scaleFactor = 0.0;
            for (int k = 0; k < c->nCat; k++)
                for (int j = 0; j < K; j++) { const double x = p3[((size_t)k * c->nPat + i) * K + j]; if (x > scaleFactor) scaleFactor = x; }
            if (scaleFactor < 1.0E-100) {
                for (int k = 0; k < c->nCat; k++)
                    for (int j = 0; j < K; j++) p3[((size_t)k * c->nPat + i) * K + j] /= scaleFactor;
                scal[i] = log(scaleFactor);
            } else scal[i] = 0.0;
        }
    }
}

/*
 * TreeLikelihood.calculateLogP for K states (2..8), everything dirty, optionally with ascertainment correction:
 * excluded[nExcluded] = the pattern indices Alignment puts in excludedPatterns (their weight is 0); then
 *   correction = log(1 - sum_excluded exp(patternLogL))            (Alignment.getAscertainmentCorrection)
 *   logP = sum_p (patternLogL[p] - correction) * weight[p]         (TreeLikelihood.calcLogP, useAscertainedSitePatterns)
 * useScaling as in sb2o_tree_likelihood.  patternLogL (optional) receives the uncorrected values.
 */
SB2O_API double sb2o_tree_likelihood_k(
    int G, int nTips, int nPat, int nCat, int K,
    const int32_t *parent, const int32_t *left, const int32_t *right, const double *height,
    const uint8_t *tipStates, const int32_t *weights,
    const double *Eval, const double *Evec, const double *Ievc, const double *freqs,
    const double *catRates, const double *catProps, const double *branchRates, int useScaling,
    int nExcluded, const int32_t *excluded, double *patternLogL)
{
    const int nInt = G - nTips;
    const size_t psz = (size_t)nCat * nPat * K;
    double *partials = (double *)malloc(sizeof(double) * psz * nInt);
    double *matrices = (double *)calloc((size_t)G * nCat * K * K, sizeof(double));
    double *scaling = (double *)calloc((size_t)nInt * nPat, sizeof(double));
    double *rootPartials = (double *)malloc(sizeof(double) * (size_t)nPat * K);
    double *plog = (double *)malloc(sizeof(double) * (size_t)nPat);
    int root = G - 1;
    for (int i = 0; i < G; i++) if (parent[i] < 0) root = i;
    double logP = 0.0;
    for (int pass = 0; pass < 2; pass++) {
        const int scaled = (useScaling == 1) || (useScaling == 2 && pass == 1);
        tlk_ctx c = {G, nTips, nPat, nCat, K, left, right, parent, height, tipStates, Eval, Evec, Ievc, catRates, branchRates,
                     partials, matrices, scaled ? scaling : NULL};
        if (root < nTips) { logP = NAN; break; }
        tlk_traverse(&c, root);
        const double *rp = partials + (size_t)(root - nTips) * psz;
        for (int k = 0; k < nPat; k++)
            for (int i = 0; i < K; i++) rootPartials[k * K + i] = rp[k * K + i] * catProps[0];
        for (int l = 1; l < nCat; l++)
            for (int k = 0; k < nPat; k++)
                for (int i = 0; i < K; i++) rootPartials[k * K + i] += rp[((size_t)l * nPat + k) * K + i] * catProps[l];
        for (int k = 0; k < nPat; k++) {
            double sum = 0.0;
            for (int i = 0; i < K; i++) sum += freqs[i] * rootPartials[k * K + i];
            double logScale = 0.0;
            if (scaled) for (int nd = 0; nd < nInt; nd++) logScale += scaling[(size_t)nd * nPat + k];
            plog[k] = log(sum) + logScale;
        }
        logP = 0.0;
        if (nExcluded > 0) {
            double excludeProb = 0.0, returnProb = 1.0;
            for (int i = 0; i < nExcluded; i++) excludeProb += exp(plog[excluded[i]]);
            returnProb -= excludeProb;
            const double corr = log(returnProb);
            for (int k = 0; k < nPat; k++) logP += (plog[k] - corr) * weights[k];
        } else {
            for (int k = 0; k < nPat; k++) logP += plog[k] * weights[k];
        }
        if (useScaling != 2 || pass == 1 || !(logP == -INFINITY)) break;
    }
    if (patternLogL) memcpy(patternLogL, plog, sizeof(double) * (size_t)nPat);
    free(partials); free(matrices); free(scaling); free(rootPartials); free(plog);
    return logP;
}

/* ------------------------------------------------------------------------------------------ */
/* (1d) operator-side walks over every gene tree -- PARITY UNPINNED: the    */
/*      reference has no test of CoordinatedUniform / CoordinatedExponential / NodeReheight2,   */
/*      so these follow the Java recursions literally and are checked by properties only.       */
/* ------------------------------------------------------------------------------------------ */

enum { DT_LEFT_ONLY = 0, DT_RIGHT_ONLY = 1, DT_BOTH = 2, DT_NEITHER = 3 };

static void min_set(double *stored, double v) { if (v < *stored) *stored = v; }   /* MinimumDouble.java:13-17 */

/* CoordinatedOperator.findDescendants -- CoordinatedOperator.java:27-43, as a membership flag per species leaf */
static void find_descendants(const int32_t *sLeft, const int32_t *sRight, int node, uint8_t *member)
{
    if (sLeft[node] < 0) { member[node] = 1; return; }
    find_descendants(sLeft, sRight, sLeft[node], member);
    find_descendants(sLeft, sRight, sRight[node], member);
}

/* CoordinatedUniform.findConnectingNodes -- CoordinatedUniform.java:118-177 (rootward != NULL) and
 * CoordinatedExponential.findConnectingNodes -- CoordinatedExponential.java:115-160 (rootward == NULL) */
static int find_connecting(int node, const int32_t *gLeft, const int32_t *gRight, const double *gHeight, const int32_t *tipSpecies,
                           const uint8_t *inLeft, const uint8_t *inRight, uint8_t *connecting, double *tipward, double *rootward)
{
    if (gLeft[node] < 0) {
        if (inLeft[tipSpecies[node]]) return DT_LEFT_ONLY;
        else if (inRight[tipSpecies[node]]) return DT_RIGHT_ONLY;
        else return DT_NEITHER;
    }
    const int l = gLeft[node], r = gRight[node];
    const int ld = find_connecting(l, gLeft, gRight, gHeight, tipSpecies, inLeft, inRight, connecting, tipward, rootward);
    const int rd = find_connecting(r, gLeft, gRight, gHeight, tipSpecies, inLeft, inRight, connecting, tipward, rootward);
    if (ld == rd) {
        if (ld == DT_BOTH) connecting[node] = 1;
        return ld;
    }
    const double h = gHeight[node];
    if (ld == DT_BOTH) {
        if (rd == DT_NEITHER) {
            if (rootward) min_set(rootward, h - gHeight[l]);
            return DT_NEITHER;
        } else {
            min_set(tipward, h - gHeight[r]);
            connecting[node] = 1;
            return DT_BOTH;
        }
    } else if (rd == DT_BOTH) {
        if (ld == DT_NEITHER) {
            if (rootward) min_set(rootward, h - gHeight[r]);
            return DT_NEITHER;
        } else {
            min_set(tipward, h - gHeight[l]);
            connecting[node] = 1;
            return DT_BOTH;
        }
    } else if (ld == DT_NEITHER || rd == DT_NEITHER) {
        return DT_NEITHER;
    } else {
        min_set(tipward, h - gHeight[l]);
        min_set(tipward, h - gHeight[r]);
        connecting[node] = 1;
        return DT_BOTH;
    }
}

/* getConnectingNodes over all gene trees -- CoordinatedUniform.java:94-116 / CoordinatedExponential.java:92-113.
 * Trees are concatenated (nodeOff/tipOff prefix sums); connecting[sum G] receives 0/1; freedoms = {tipward, rootward}
 * start at +inf (MinimumDouble.java:9-11); exponential != 0 leaves rootward untouched.  Returns the number of connecting nodes. */
SB2O_API int sb2o_connecting_nodes(int S, const int32_t *sLeft, const int32_t *sRight, int speciesNode, int nTrees, const int64_t *nodeOff,
                          const int64_t *tipOff, const int32_t *gParent, const int32_t *gLeft, const int32_t *gRight, const double *gHeight,
                          const int32_t *tipSpecies, int exponential, uint8_t *connecting, double *freedoms)
{
    uint8_t *inLeft = (uint8_t *)calloc((size_t)S, 1), *inRight = (uint8_t *)calloc((size_t)S, 1);
    find_descendants(sLeft, sRight, sLeft[speciesNode], inLeft);
    find_descendants(sLeft, sRight, sRight[speciesNode], inRight);
    freedoms[0] = INFINITY; freedoms[1] = INFINITY;
    int count = 0;
    for (int j = 0; j < nTrees; j++) {
        const int64_t o = nodeOff[j];
        const int G = (int)(nodeOff[j + 1] - o);
        int root = 0;
        while (gParent[o + root] >= 0) root++;
        memset(connecting + o, 0, (size_t)G);
        find_connecting(root, gLeft + o, gRight + o, gHeight + o, tipSpecies + tipOff[j], inLeft, inRight, connecting + o,
                        &freedoms[0], exponential ? NULL : &freedoms[1]);
        for (int i = 0; i < G; i++) count += connecting[o + i];
    }
    free(inLeft); free(inRight);
    return count;
}

enum { RP_LEFT = 0, RP_RIGHT = 1, RP_BOTH = 2 };

/* NodeReheight2.recurseMaxHeight -- NodeReheight2.java:175-203 */
static int recurse_max_height(int node, const int32_t *gLeft, const int32_t *gRight, const double *gHeight, const uint8_t *leafPositions, double *maxHeight)
{
    const int l = gLeft[node], r = gRight[node];
    const int lp = gLeft[l] < 0 ? leafPositions[l] : recurse_max_height(l, gLeft, gRight, gHeight, leafPositions, maxHeight);
    const int rp = gLeft[r] < 0 ? leafPositions[r] : recurse_max_height(r, gLeft, gRight, gHeight, leafPositions, maxHeight);
    if (lp == rp) return lp;
    if (lp != RP_BOTH && rp != RP_BOTH) *maxHeight = fmin(*maxHeight, gHeight[node]);
    return RP_BOTH;
}

/* NodeReheight2.recalculateMaxHeight -- NodeReheight2.java:150-173.  canonicalMap[species node] = index in the canonical
 * order (:228,:238,:249); a leaf is LEFT when that index < centerIndex.  `origin` = +inf when no origin is specified. */
SB2O_API double sb2o_max_height(const int32_t *canonicalMap, int centerIndex, double origin, int nTrees, const int64_t *nodeOff, const int64_t *tipOff,
                       const int32_t *gParent, const int32_t *gLeft, const int32_t *gRight, const double *gHeight, const int32_t *tipSpecies)
{
    double maxHeight = origin;
    for (int i = 0; i < nTrees; i++) {
        const int64_t o = nodeOff[i];
        const int nTips = (int)(tipOff[i + 1] - tipOff[i]);
        uint8_t *leafPositions = (uint8_t *)malloc((size_t)nTips);
        for (int j = 0; j < nTips; j++) leafPositions[j] = canonicalMap[tipSpecies[tipOff[i] + j]] < centerIndex ? RP_LEFT : RP_RIGHT;
        int root = 0;
        while (gParent[o + root] >= 0) root++;
        recurse_max_height(root, gLeft + o, gRight + o, gHeight + o, leafPositions, &maxHeight);
        free(leafPositions);
    }
    return maxHeight;
}

/*
 * TreeLikelihood.traverse with BEAST's dirtiness logic, for the regime-B CPU baseline: the matrix of a
 * branch is recomputed when branchChanged[node] is set (node dirty, or length x rate differs from the cached value), and the
 * partials of an internal node when either child reported an update.  Unscaled cache (the lazy rule never switched
 * scaling on); returns the update flag like the Java.
 */
static int tl_traverse_update(const tl_ctx *c, int node, const uint8_t *branchChanged)
{
    int update = 0;
    const int par = c->parent[node];
    if (par >= 0 && branchChanged[node]) {
        const double branchRate = c->branchRates[node];
        for (int i = 0; i < c->nCat; i++) {
            const double distance = (c->height[par] - c->height[node]) * (c->catRates[i] * branchRate);
            transition_matrix(c->substKind, c->substParams, distance, c->matrices + ((size_t)node * c->nCat + i) * 16);
        }
        update = 1;
    }
    if (node >= c->nTips) {
        const int c1 = c->left[node], c2 = c->right[node];
        const int u1 = tl_traverse_update(c, c1, branchChanged);
        const int u2 = tl_traverse_update(c, c2, branchChanged);
        if (u1 || u2) {
            const size_t psz = (size_t)c->nCat * c->nPat * 4;
            double *p3 = c->partials + (size_t)(node - c->nTips) * psz;
            const double *m1 = c->matrices + (size_t)c1 * c->nCat * 16, *m2 = c->matrices + (size_t)c2 * c->nCat * 16;
            if (c1 < c->nTips) {
                if (c2 < c->nTips)
                    states_states(c->tipStates + (size_t)c1 * c->nPat, m1, c->tipStates + (size_t)c2 * c->nPat, m2, p3, c->nPat, c->nCat);
                else
                    states_partials(c->tipStates + (size_t)c1 * c->nPat, m1, c->partials + (size_t)(c2 - c->nTips) * psz, m2, p3, c->nPat, c->nCat);
            } else {
                if (c2 < c->nTips)
                    states_partials(c->tipStates + (size_t)c2 * c->nPat, m2, c->partials + (size_t)(c1 - c->nTips) * psz, m1, p3, c->nPat, c->nCat);
                else
                    partials_partials(c->partials + (size_t)(c1 - c->nTips) * psz, m1, c->partials + (size_t)(c2 - c->nTips) * psz, m2, p3, c->nPat, c->nCat);
            }
            update = 1;
        }
    }
    return update;
}

/* Incremental TreeLikelihood.calculateLogP on cached (unscaled) partials / matrices from sb2o_tree_likelihood(useScaling=0).
 * Returns the log-likelihood; *nodesUpdated receives the number of internal nodes whose partials were recomputed. */
SB2O_API double sb2o_tree_likelihood_update(
    int G, int nTips, int nPat, int nCat,
    const int32_t *parent, const int32_t *left, const int32_t *right, const double *height,
    const uint8_t *tipStates, const int32_t *weights,
    int substKind, const double *substParams,
    const double *catRates, const double *catProps, double pInv, const uint8_t *constMask,
    const double *branchRates, const uint8_t *branchChanged,
    double *partials, double *matrices, double *rootPartials /* scratch [nPat*4] */)
{
    const size_t psz = (size_t)nCat * nPat * 4;
    int root = G - 1;
    for (int i = 0; i < G; i++) if (parent[i] < 0) root = i;
    const double *freqs = subst_freqs(substKind, substParams);
    tl_ctx c = {G, nTips, nPat, nCat, left, right, parent, height, tipStates, substKind, substParams,
                catRates, branchRates, partials, matrices, NULL};
    tl_traverse_update(&c, root, branchChanged);
    const double *rp = partials + (size_t)(root - nTips) * psz;
    for (int k = 0; k < nPat; k++)
        for (int i = 0; i < 4; i++) rootPartials[k * 4 + i] = rp[k * 4 + i] * catProps[0];
    for (int l = 1; l < nCat; l++)
        for (int k = 0; k < nPat; k++)
            for (int i = 0; i < 4; i++) rootPartials[k * 4 + i] += rp[((size_t)l * nPat + k) * 4 + i] * catProps[l];
    if (pInv > 0.0 && constMask)
        for (int k = 0; k < nPat; k++)
            for (int i = 0; i < 4; i++)
                if (constMask[k] & (1 << i)) rootPartials[k * 4 + i] += pInv;
    double logP = 0.0;
    for (int k = 0; k < nPat; k++) {
        double sum = 0.0;
        for (int i = 0; i < 4; i++) sum += freqs[i] * rootPartials[k * 4 + i];
        logP += log(sum) * weights[k];
    }
    return logP;
}

/* ------------------------------------------------------------------------------------------ */
/* (3) whole-posterior driver: the timed CPU baseline (bench.py cpu_baseline / --impl reference) */
/* ------------------------------------------------------------------------------------------ */

enum { SB2O_CLOCK_STRICT = 0, SB2O_CLOCK_SPECIES_RATES = 1, SB2O_CLOCK_EXPLICIT = 2 };
enum { SB2O_POP_CONSTANT = 0, SB2O_POP_LWCR = 1, SB2O_POP_ANALYTIC = 2 };

typedef struct {
    int32_t nLoci, S, nSpeciesLeaves, popKind;
    const int32_t *sParent, *sLeft, *sRight;
    const double *sHeight;
    const double *speciesRates;   /* [S] SpeciesTreeRates.getRatesArray() */
    const double *popA, *popB;    /* CONSTANT: popA=N[S]; LWCR: popA=tip[leaves], popB=top[S-1] */
    double alpha, beta;           /* ANALYTIC */
    const int32_t *nTips, *nPat, *nCat, *substKind, *clockKind;
    const int64_t *nodeOff, *tipOff, *patOff, *stateOff, *catOff; /* exclusive prefix sums, nLoci+1 entries */
    const int32_t *gParent, *gLeft, *gRight;
    const double *gHeight;
    const int32_t *tipSpecies;
    const uint8_t *tipStates;
    const int32_t *weights;
    const uint8_t *constMask;
    const double *ploidy, *substParams /* [nLoci][40] */, *catRates, *catProps, *pInv, *clockRate;
    const double *explicitRates;  /* by nodeOff, used for CLOCK_EXPLICIT */
} sb2o_problem;

/*
 * One full posterior evaluation the way the reference computes it with every node dirty:
 * per locus GeneTree.update (with the G x S occupancy fill), per-branch pop-model terms or the
 * analytic pooled term, StarBeastClock rates, then TreeLikelihood.  Loci loop is OpenMP-parallel
 * when nThreads > 1 (upper bound of what CompoundDistribution useThreads could reach).
 * Outputs: perLocusCoal[nLoci] (GeneTree.calculateLogP), perLocusLik[nLoci], perBranch[S]
 * (analytic mode), total[3] = {coalescent (MultispeciesCoalescent.calculateLogP), likelihood, sum}.
 */
SB2O_API int sb2o_posterior(const sb2o_problem *pb, int nThreads, int useScaling,
                            double *perLocusCoal, double *perLocusLik, double *perBranch, double *total)
{
    const int S = pb->S, L = pb->nLoci;
    int maxG = 0;
    for (int l = 0; l < L; l++) { const int G = 2 * pb->nTips[l] - 1; if (G > maxG) maxG = G; }
    /* analytic mode needs (n, k, sorted times) of every (locus, branch) after the loci loop */
    int32_t *allN = NULL, *allK = NULL; double *allT = NULL; const int tcap = maxG + 2;
    if (pb->popKind == SB2O_POP_ANALYTIC) {
        allN = (int32_t *)malloc(sizeof(int32_t) * (size_t)L * S);
        allK = (int32_t *)malloc(sizeof(int32_t) * (size_t)L * S);
        allT = (double *)malloc(sizeof(double) * (size_t)L * S * tcap);
    }
#ifdef _OPENMP
    if (nThreads < 1) nThreads = 1;
    omp_set_num_threads(nThreads);
#else
    (void)nThreads;
#endif
#pragma omp parallel
    {
        int32_t *n = (int32_t *)malloc(sizeof(int32_t) * S), *k = (int32_t *)malloc(sizeof(int32_t) * S);
        int32_t *assign = (int32_t *)malloc(sizeof(int32_t) * maxG);
        double *times = (double *)malloc(sizeof(double) * (size_t)S * maxG);
        double *occ = (double *)malloc(sizeof(double) * (size_t)maxG * S);
        double *bt = (double *)malloc(sizeof(double) * (maxG + 2));
        double *rates = (double *)malloc(sizeof(double) * maxG);
#pragma omp for schedule(dynamic, 1)
        for (int l = 0; l < L; l++) {
            const int nTips = pb->nTips[l], G = 2 * nTips - 1, blk = nTips;
            const int64_t no = pb->nodeOff[l];
            const int ok = sb2o_gene_tree_update(G, nTips, pb->gParent + no, pb->gHeight + no, pb->tipSpecies + pb->tipOff[l],
                                                 S, pb->nSpeciesLeaves, pb->sParent, pb->sHeight, blk, n, k, times, assign, occ);
            double coal = 0.0;
            if (!ok) coal = -INFINITY;                                   /* GeneTree.java:175-178 */
            else if (pb->popKind == SB2O_POP_ANALYTIC) {                  /* GeneTree.java:182 */
                for (int b = 0; b < S; b++) {
                    allN[(size_t)l * S + b] = n[b]; allK[(size_t)l * S + b] = k[b];
                    sb2o_coalescent_times(b, pb->sParent, pb->sHeight, k, times, blk, allT + ((size_t)l * S + b) * tcap);
                }
            } else {
                for (int b = 0; b < S; b++) {                             /* GeneTree.java:185-196 */
                    sb2o_coalescent_times(b, pb->sParent, pb->sHeight, k, times, blk, bt);
                    double t;
                    if (pb->popKind == SB2O_POP_CONSTANT) t = sb2o_constant_logp(pb->popA[b], pb->ploidy[l], bt, n[b], k[b]);
                    else t = sb2o_lwcr_branch_logp(b, pb->nSpeciesLeaves, pb->sParent, pb->sLeft, pb->sRight, pb->popA, pb->popB, pb->ploidy[l], bt, n[b], k[b]);
                    coal += t;
                }
            }
            perLocusCoal[l] = coal;
            /* CompoundDistribution short-circuit: likelihood is not evaluated once the coalescent is -inf;
               still computed per locus here only if compatible. */
            if (!ok) { perLocusLik[l] = NAN; continue; }
            const double *br;
            if (pb->clockKind[l] == SB2O_CLOCK_SPECIES_RATES) {
                sb2o_starbeast_clock(G, S, pb->clockRate[l], pb->speciesRates, occ, rates); br = rates;
            } else if (pb->clockKind[l] == SB2O_CLOCK_EXPLICIT) {
                br = pb->explicitRates + no;
            } else {
                for (int i = 0; i < G; i++) rates[i] = pb->clockRate[l];   /* StrictClockModel */
                br = rates;
            }
            perLocusLik[l] = sb2o_tree_likelihood(G, nTips, pb->nPat[l], pb->nCat[l], pb->gParent + no, pb->gLeft + no, pb->gRight + no,
                                                  pb->gHeight + no, pb->tipStates + pb->stateOff[l], pb->weights + pb->patOff[l],
                                                  pb->substKind[l], pb->substParams + (size_t)l * SB2O_SUBST_STRIDE,
                                                  pb->catRates + pb->catOff[l], pb->catProps + pb->catOff[l], pb->pInv[l],
                                                  pb->constMask ? pb->constMask + pb->patOff[l] : NULL, br, useScaling, NULL, NULL, NULL, NULL);
        }
        free(n); free(k); free(assign); free(times); free(occ); free(bt); free(rates);
    }
    /* CompoundDistribution.calculateLogP over the GeneTrees, then MultispeciesCoalescent.java:150-198 */
    double coalTotal = 0.0; int infinite = 0;
    for (int l = 0; l < L; l++) { coalTotal += perLocusCoal[l]; if (isinf(perLocusCoal[l]) || isnan(perLocusCoal[l])) { infinite = 1; break; } }
    if (pb->popKind == SB2O_POP_ANALYTIC && !infinite) {
        int32_t *bn = (int32_t *)malloc(sizeof(int32_t) * L), *bk = (int32_t *)malloc(sizeof(int32_t) * L);
        int64_t *off = (int64_t *)malloc(sizeof(int64_t) * L);
        for (int b = 0; b < S; b++) {
            for (int l = 0; l < L; l++) { bn[l] = allN[(size_t)l * S + b]; bk[l] = allK[(size_t)l * S + b]; off[l] = ((int64_t)l * S + b) * tcap; }
            const double t = sb2o_analytical_logp(pb->alpha, pb->beta, L, pb->ploidy, allT, off, bn, bk);
            if (perBranch) perBranch[b] = t;
            coalTotal += t;                                               /* MultispeciesCoalescent.java:194 */
        }
        free(bn); free(bk); free(off);
    }
    free(allN); free(allK); free(allT);
    double likTotal = 0.0;
    if (!infinite) for (int l = 0; l < L; l++) likTotal += perLocusLik[l];
    else likTotal = NAN;
    total[0] = coalTotal; total[1] = likTotal; total[2] = infinite ? coalTotal : coalTotal + likTotal;
    return 0;
}

SB2O_API int sb2o_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* (4) regime-B CPU baseline: one MCMC step the way the reference evaluates it  */
/* ------------------------------------------------------------------------------------------ */
/*
 * What the reference does when ONE gene tree changed (single Java thread):
 *   - GeneTree.update() of EVERY locus (requiresRecalculation() is unconditionally true, GeneTree.java:73-76, quirk Q2),
 *     including the G x S occupancy fill;
 *   - per-branch population terms only where the embedding changed (GeneTree.java:187); here: all branches of the changed
 *     locus (an upper bound), cached values for the others;
 *   - StarBeastClock rates + TreeLikelihood.traverse for the changed locus only: matrices of branches whose length x rate
 *     changed, partials of the nodes above them, root integration over all patterns.
 * Non-analytic population models only.  The caller owns the problem; the proposed tree of `locus` is already in place.
 */
typedef struct {
    int L, maxG, S;
    double **partials, **matrices, **branchLen;   /* per locus: unscaled partials, matrices, cached length*rate per node */
    double *coal, *logL;
    int32_t *n, *k, *assign; double *times, *occ, *bt, *rates, *rootPartials; uint8_t *changed;
} sb2o_incr;

static double incr_locus(sb2o_incr *c, const sb2o_problem *pb, int l, int full)
{
    const int S = pb->S, nTips = pb->nTips[l], G = 2 * nTips - 1, blk = nTips;
    const int64_t no = pb->nodeOff[l];
    const int ok = sb2o_gene_tree_update(G, nTips, pb->gParent + no, pb->gHeight + no, pb->tipSpecies + pb->tipOff[l],
                                         S, pb->nSpeciesLeaves, pb->sParent, pb->sHeight, blk, c->n, c->k, c->times, c->assign, c->occ);
    if (!full) return ok ? c->coal[l] : -INFINITY;          /* clean locus: embedding only, cached terms */
    if (!ok) { c->coal[l] = -INFINITY; return -INFINITY; }
    double coal = 0.0;
    for (int b = 0; b < S; b++) {
        sb2o_coalescent_times(b, pb->sParent, pb->sHeight, c->k, c->times, blk, c->bt);
        coal += pb->popKind == SB2O_POP_CONSTANT
                    ? sb2o_constant_logp(pb->popA[b], pb->ploidy[l], c->bt, c->n[b], c->k[b])
                    : sb2o_lwcr_branch_logp(b, pb->nSpeciesLeaves, pb->sParent, pb->sLeft, pb->sRight, pb->popA, pb->popB, pb->ploidy[l], c->bt, c->n[b], c->k[b]);
    }
    c->coal[l] = coal;
    const double *br;
    if (pb->clockKind[l] == SB2O_CLOCK_SPECIES_RATES) { sb2o_starbeast_clock(G, S, pb->clockRate[l], pb->speciesRates, c->occ, c->rates); br = c->rates; }
    else if (pb->clockKind[l] == SB2O_CLOCK_EXPLICIT) br = pb->explicitRates + no;
    else { for (int i = 0; i < G; i++) c->rates[i] = pb->clockRate[l]; br = c->rates; }
    for (int i = 0; i < G; i++) {                            /* TreeLikelihood.traverse: branchTime != m_branchLengths[i] */
        const int par = pb->gParent[no + i];
        const double len = par >= 0 ? (pb->gHeight[no + par] - pb->gHeight[no + i]) * br[i] : 0.0;
        c->changed[i] = (uint8_t)(c->branchLen[l][i] != len);
        c->branchLen[l][i] = len;
    }
    c->logL[l] = sb2o_tree_likelihood_update(G, nTips, pb->nPat[l], pb->nCat[l], pb->gParent + no, pb->gLeft + no, pb->gRight + no,
                                             pb->gHeight + no, pb->tipStates + pb->stateOff[l], pb->weights + pb->patOff[l],
                                             pb->substKind[l], pb->substParams + (size_t)l * SB2O_SUBST_STRIDE,
                                             pb->catRates + pb->catOff[l], pb->catProps + pb->catOff[l], pb->pInv[l],
                                             pb->constMask ? pb->constMask + pb->patOff[l] : NULL, br, c->changed,
                                             c->partials[l], c->matrices[l], c->rootPartials);
    return coal;
}

SB2O_API sb2o_incr *sb2o_incr_create(const sb2o_problem *pb)
{
    if (pb->popKind == SB2O_POP_ANALYTIC) return NULL;
    sb2o_incr *c = (sb2o_incr *)calloc(1, sizeof(sb2o_incr));
    const int L = pb->nLoci, S = pb->S;
    int maxG = 0, maxP = 0;
    for (int l = 0; l < L; l++) { const int G = 2 * pb->nTips[l] - 1; if (G > maxG) maxG = G; if (pb->nPat[l] > maxP) maxP = pb->nPat[l]; }
    c->L = L; c->maxG = maxG; c->S = S;
    c->partials = (double **)calloc(L, sizeof(double *)); c->matrices = (double **)calloc(L, sizeof(double *)); c->branchLen = (double **)calloc(L, sizeof(double *));
    c->coal = (double *)calloc(L, sizeof(double)); c->logL = (double *)calloc(L, sizeof(double));
    c->n = (int32_t *)malloc(sizeof(int32_t) * S); c->k = (int32_t *)malloc(sizeof(int32_t) * S); c->assign = (int32_t *)malloc(sizeof(int32_t) * maxG);
    c->times = (double *)malloc(sizeof(double) * (size_t)S * maxG); c->occ = (double *)malloc(sizeof(double) * (size_t)maxG * S);
    c->bt = (double *)malloc(sizeof(double) * (maxG + 2)); c->rates = (double *)malloc(sizeof(double) * maxG);
    c->rootPartials = (double *)malloc(sizeof(double) * (size_t)maxP * 4); c->changed = (uint8_t *)malloc(maxG);
    for (int l = 0; l < L; l++) {
        const int nTips = pb->nTips[l], G = 2 * nTips - 1;
        c->partials[l] = (double *)calloc((size_t)(G - nTips) * pb->nCat[l] * pb->nPat[l] * 4, sizeof(double));
        c->matrices[l] = (double *)calloc((size_t)G * pb->nCat[l] * 16, sizeof(double));
        c->branchLen[l] = (double *)malloc(sizeof(double) * G);
        for (int i = 0; i < G; i++) c->branchLen[l][i] = NAN;          /* everything dirty the first time */
        incr_locus(c, pb, l, 1);
    }
    return c;
}

/* one step: the tree of `locus` changed (already written into pb).  Returns the posterior; -inf if incompatible. */
SB2O_API double sb2o_incr_step(sb2o_incr *c, const sb2o_problem *pb, int locus)
{
    double coal = 0.0, lik = 0.0;
    for (int l = 0; l < c->L; l++) coal += incr_locus(c, pb, l, l == locus);
    if (!(coal > -INFINITY)) return -INFINITY;                 /* CompoundDistribution short-circuit */
    for (int l = 0; l < c->L; l++) lik += c->logL[l];
    return coal + lik;
}

SB2O_API void sb2o_incr_destroy(sb2o_incr *c)
{
    if (!c) return;
    for (int l = 0; l < c->L; l++) { free(c->partials[l]); free(c->matrices[l]); free(c->branchLen[l]); }
    free(c->partials); free(c->matrices); free(c->branchLen); free(c->coal); free(c->logL);
    free(c->n); free(c->k); free(c->assign); free(c->times); free(c->occ); free(c->bt); free(c->rates); free(c->rootPartials); free(c->changed);
    free(c);
}

