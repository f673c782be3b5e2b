/*
 * sb2gpu_jni.c -- JNI shim between starbeast2.gpu.Sb2Gpu and libsb2gpu's C ABI (include/sb2gpu.h).
 * No JDK in the build image: compiled and executed there against tests/jni_stub (stand-in jni.h + mock JNIEnv,
 * tests/test_jni_shim.py); see java/README.md.  Arrays are pinned with
 * GetPrimitiveArrayCritical for the duration of one C call: the engine copies inputs before returning
 * and writes outputs into the caller's buffers, so no reference outlives the call.
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>

#include "sb2gpu.h"

#define E(h) ((sb2_engine *)(intptr_t)(h))
#define PIN(a) ((a) ? (*env)->GetPrimitiveArrayCritical(env, (a), NULL) : NULL)
#define UNPIN(a, p, mode) do { if (a) (*env)->ReleasePrimitiveArrayCritical(env, (a), (p), (mode)); } while (0)
#define FN(name) Java_starbeast2_gpu_Sb2Gpu_##name

JNIEXPORT jlong JNICALL FN(create)(JNIEnv *env, jclass c, jint device, jboolean exact)
{
    sb2_config cfg = {0};
    sb2_engine *e = NULL;
    cfg.device = device; cfg.exact_order = exact ? 1 : 0;
    if (sb2_create(&cfg, &e) != SB2_OK) {
        jclass ex = (*env)->FindClass(env, "java/lang/IllegalStateException");
        (*env)->ThrowNew(env, ex, sb2_last_error(e));   /* no CPU fallback: fail loudly */
        sb2_destroy(e);
        return 0;
    }
    return (jlong)(intptr_t)e;
}
JNIEXPORT void JNICALL FN(destroy)(JNIEnv *env, jclass c, jlong h) { sb2_destroy(E(h)); }
JNIEXPORT jstring JNICALL FN(lastError)(JNIEnv *env, jclass c, jlong h) { return (*env)->NewStringUTF(env, sb2_last_error(E(h))); }

JNIEXPORT jint JNICALL FN(addLocus)(JNIEnv *env, jclass c, jlong h, jint nTips, jint nPat, jbyteArray states, jintArray weights,
                                    jint nCat, jintArray tipSp, jdouble ploidy)
{
    void *s = PIN(states), *w = PIN(weights), *t = PIN(tipSp);
    int rc = sb2_add_locus(E(h), nTips, nPat, (const uint8_t *)s, (const int32_t *)w, nCat, (const int32_t *)t, ploidy);
    UNPIN(tipSp, t, JNI_ABORT); UNPIN(weights, w, JNI_ABORT); UNPIN(states, s, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(addLocusK)(JNIEnv *env, jclass c, jlong h, jint nTips, jint nPat, jint nStates, jbyteArray states, jintArray weights,
                                     jint nCat, jintArray tipSp, jdouble ploidy, jintArray excluded)
{
    const jint nExcl = excluded ? (*env)->GetArrayLength(env, excluded) : 0;
    void *s = PIN(states), *w = PIN(weights), *t = PIN(tipSp), *x = PIN(excluded);
    int rc = sb2_add_locus_k(E(h), nTips, nPat, nStates, (const uint8_t *)s, (const int32_t *)w, nCat, (const int32_t *)t, ploidy,
                             nExcl, (const int32_t *)x);
    UNPIN(excluded, x, JNI_ABORT); UNPIN(tipSp, t, JNI_ABORT); UNPIN(weights, w, JNI_ABORT); UNPIN(states, s, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(finalizeEngine)(JNIEnv *env, jclass c, jlong h, jint S, jint leaves) { return sb2_finalize(E(h), S, leaves); }

static jint set_tree(JNIEnv *env, jlong h, jint locus, jintArray p, jintArray l, jintArray r, jdoubleArray ht)
{
    void *pp = PIN(p), *ll = PIN(l), *rr = PIN(r), *hh = PIN(ht);
    int rc = locus < 0 ? sb2_set_species_tree(E(h), pp, ll, rr, hh) : sb2_set_gene_tree(E(h), locus, pp, ll, rr, hh);
    UNPIN(ht, hh, JNI_ABORT); UNPIN(r, rr, JNI_ABORT); UNPIN(l, ll, JNI_ABORT); UNPIN(p, pp, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(setSpeciesTree)(JNIEnv *env, jclass c, jlong h, jintArray p, jintArray l, jintArray r, jdoubleArray ht)
{ return set_tree(env, h, -1, p, l, r, ht); }
JNIEXPORT jint JNICALL FN(setGeneTree)(JNIEnv *env, jclass c, jlong h, jint locus, jintArray p, jintArray l, jintArray r, jdoubleArray ht)
{ return set_tree(env, h, locus, p, l, r, ht); }

JNIEXPORT jint JNICALL FN(setSubstModel)(JNIEnv *env, jclass c, jlong h, jint locus, jint kind, jdoubleArray params)
{
    void *p = PIN(params);
    int rc = sb2_set_subst_model(E(h), locus, kind, p);
    UNPIN(params, p, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(setSiteModel)(JNIEnv *env, jclass c, jlong h, jint locus, jdoubleArray rates, jdoubleArray props, jdouble pInv)
{
    void *r = PIN(rates), *p = PIN(props);
    int rc = sb2_set_site_model(E(h), locus, r, p, pInv);
    UNPIN(props, p, JNI_ABORT); UNPIN(rates, r, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(setClock)(JNIEnv *env, jclass c, jlong h, jint locus, jint kind, jdouble mean, jdoubleArray rates)
{
    void *r = PIN(rates);
    int rc = sb2_set_clock(E(h), locus, kind, mean, r);
    UNPIN(rates, r, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(setSpeciesRates)(JNIEnv *env, jclass c, jlong h, jdoubleArray rates)
{
    void *r = PIN(rates);
    int rc = sb2_set_species_rates(E(h), r);
    UNPIN(rates, r, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(setPopModel)(JNIEnv *env, jclass c, jlong h, jint kind, jdoubleArray a, jdoubleArray b, jdouble alpha, jdouble beta)
{
    void *pa = PIN(a), *pb = PIN(b);
    int rc = sb2_set_pop_model(E(h), kind, pa, pb, alpha, beta);
    UNPIN(b, pb, JNI_ABORT); UNPIN(a, pa, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(evalCoalescent)(JNIEnv *env, jclass c, jlong h, jdoubleArray perLocus, jdoubleArray perBranch, jdoubleArray total)
{
    void *l = PIN(perLocus), *b = PIN(perBranch), *t = PIN(total);
    int rc = sb2_eval_coalescent(E(h), l, b, t);
    UNPIN(total, t, 0); UNPIN(perBranch, b, 0); UNPIN(perLocus, l, 0);
    return rc;
}
JNIEXPORT jint JNICALL FN(evalLikelihood)(JNIEnv *env, jclass c, jlong h, jdoubleArray perLocus, jdoubleArray total)
{
    void *l = PIN(perLocus), *t = PIN(total);
    int rc = sb2_eval_likelihood(E(h), l, t);
    UNPIN(total, t, 0); UNPIN(perLocus, l, 0);
    return rc;
}
JNIEXPORT jint JNICALL FN(evalPosterior)(JNIEnv *env, jclass c, jlong h, jdoubleArray coal, jdoubleArray lik, jdoubleArray br, jdoubleArray tot3)
{
    void *a = PIN(coal), *l = PIN(lik), *b = PIN(br), *t = PIN(tot3);
    int rc = sb2_eval_posterior(E(h), a, l, b, t);
    UNPIN(tot3, t, 0); UNPIN(br, b, 0); UNPIN(lik, l, 0); UNPIN(coal, a, 0);
    return rc;
}
JNIEXPORT jint JNICALL FN(store)(JNIEnv *env, jclass c, jlong h) { return sb2_store(E(h)); }
JNIEXPORT jint JNICALL FN(restore)(JNIEnv *env, jclass c, jlong h) { return sb2_restore(E(h)); }
JNIEXPORT jint JNICALL FN(accept)(JNIEnv *env, jclass c, jlong h) { return sb2_accept(E(h)); }
JNIEXPORT jint JNICALL FN(markAllDirty)(JNIEnv *env, jclass c, jlong h) { return sb2_mark_all_dirty(E(h)); }
JNIEXPORT jint JNICALL FN(getBranchRates)(JNIEnv *env, jclass c, jlong h, jint locus, jdoubleArray rates)
{
    void *r = PIN(rates);
    int rc = sb2_get_branch_rates(E(h), locus, r);
    UNPIN(rates, r, 0);
    return rc;
}
JNIEXPORT jint JNICALL FN(getEmbedding)(JNIEnv *env, jclass c, jlong h, jint locus, jintArray n, jintArray k, jintArray assign, jdoubleArray br)
{
    void *pn = PIN(n), *pk = PIN(k), *pa = PIN(assign), *pb = PIN(br);
    int rc = sb2_get_embedding(E(h), locus, pn, pk, pa, pb);
    UNPIN(br, pb, 0); UNPIN(assign, pa, 0); UNPIN(k, pk, 0); UNPIN(n, pn, 0);
    return rc;
}
JNIEXPORT jint JNICALL FN(connectingNodes)(JNIEnv *env, jclass c, jlong h, jint speciesNode, jbyteArray mask, jdoubleArray freedoms, jlongArray count)
{
    void *m = PIN(mask), *f = PIN(freedoms), *n = PIN(count);
    int rc = sb2_connecting_nodes(E(h), speciesNode, (uint8_t *)m, (double *)f, (int64_t *)n);
    UNPIN(count, n, 0); UNPIN(freedoms, f, 0); UNPIN(mask, m, 0);
    return rc;
}
JNIEXPORT jint JNICALL FN(maxHeight)(JNIEnv *env, jclass c, jlong h, jbyteArray leafIsRight, jdoubleArray out)
{
    void *l = PIN(leafIsRight), *o = PIN(out);
    int rc = sb2_max_height(E(h), (const uint8_t *)l, (double *)o);
    UNPIN(out, o, 0); UNPIN(leafIsRight, l, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(evalBegin)(JNIEnv *env, jclass c, jlong h) { return sb2_eval_begin(E(h)); }
JNIEXPORT jint JNICALL FN(evalEnd)(JNIEnv *env, jclass c, jlong h, jdoubleArray coal, jdoubleArray lik, jdoubleArray br, jdoubleArray tot)
{
    void *pc = PIN(coal), *pl = PIN(lik), *pb = PIN(br), *pt = PIN(tot);
    int rc = sb2_eval_end(E(h), pc, pl, pb, pt);
    UNPIN(tot, pt, 0); UNPIN(br, pb, 0); UNPIN(lik, pl, 0); UNPIN(coal, pc, 0);
    return rc;
}
JNIEXPORT jint JNICALL FN(reduceVectorLength)(JNIEnv *env, jclass c, jlong h)
{
    void *dev = NULL; int32_t n = 0;
    return sb2_reduce_vector(E(h), &dev, &n) == SB2_OK ? n : -1;
}
JNIEXPORT jint JNICALL FN(readReduceVector)(JNIEnv *env, jclass c, jlong h, jdoubleArray host)
{
    void *p = PIN(host);
    int rc = sb2_read_reduce_vector(E(h), p);
    UNPIN(host, p, 0);
    return rc;
}
JNIEXPORT jint JNICALL FN(writeReduceVector)(JNIEnv *env, jclass c, jlong h, jdoubleArray host)
{
    void *p = PIN(host);
    int rc = sb2_write_reduce_vector(E(h), p);
    UNPIN(host, p, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(setGeneTrees)(JNIEnv *env, jclass c, jlong h, jint first, jint n, jintArray par, jintArray left, jintArray right, jdoubleArray height)
{
    void *p = PIN(par), *l = PIN(left), *r = PIN(right), *hh = PIN(height);
    int rc = sb2_set_gene_trees(E(h), first, n, p, l, r, hh);
    UNPIN(height, hh, JNI_ABORT); UNPIN(right, r, JNI_ABORT); UNPIN(left, l, JNI_ABORT); UNPIN(par, p, JNI_ABORT);
    return rc;
}
JNIEXPORT jint JNICALL FN(numLoci)(JNIEnv *env, jclass c, jlong h) { return sb2_num_loci(E(h)); }
JNIEXPORT jlong JNICALL FN(launchCount)(JNIEnv *env, jclass c, jlong h) { return sb2_launch_count(E(h)); }
