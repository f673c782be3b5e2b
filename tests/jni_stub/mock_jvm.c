/*
 * mock_jvm.c -- a mock JNIEnv for tests/test_jni_shim.py (test infrastructure; see tests/jni_stub/jni.h).
 *
 * "Java arrays" are heap objects that point at caller-owned memory; Get/ReleasePrimitiveArrayCritical hand that memory
 * out and count pin / unpin pairs so that the test can check the shim releases everything it pins.  ThrowNew records
 * the pending exception's class and message.
 */
#include <stdlib.h>
#include <string.h>

#include "jni.h"

struct _jobject {
    int kind;            /* 1 array, 2 class, 3 string */
    void *data;
    int length;
    int pins;
    char text[256];
};

static int g_pins = 0, g_unpins = 0, g_live = 0;
static char g_exClass[256], g_exMsg[512];

static jclass JNICALL m_FindClass(JNIEnv *env, const char *name)
{
    (void)env;
    struct _jobject *o = calloc(1, sizeof *o);
    o->kind = 2;
    strncpy(o->text, name, sizeof o->text - 1);
    return o;
}
static jint JNICALL m_ThrowNew(JNIEnv *env, jclass clazz, const char *msg)
{
    (void)env;
    strncpy(g_exClass, clazz ? clazz->text : "", sizeof g_exClass - 1);
    strncpy(g_exMsg, msg ? msg : "", sizeof g_exMsg - 1);
    return 0;
}
static jint JNICALL m_GetArrayLength(JNIEnv *env, jarray a) { (void)env; return a->length; }
static jstring JNICALL m_NewStringUTF(JNIEnv *env, const char *utf)
{
    (void)env;
    struct _jobject *o = calloc(1, sizeof *o);
    o->kind = 3;
    strncpy(o->text, utf ? utf : "", sizeof o->text - 1);
    return o;
}
static void *JNICALL m_GetCritical(JNIEnv *env, jarray a, jboolean *isCopy)
{
    (void)env;
    if (isCopy) *isCopy = 0;
    a->pins++; g_pins++; g_live++;
    return a->data;
}
static void JNICALL m_ReleaseCritical(JNIEnv *env, jarray a, void *carray, jint mode)
{
    (void)env; (void)mode;
    if (carray != a->data) abort();   /* the shim must hand back the pointer it was given */
    a->pins--; g_unpins++; g_live--;
}

static const struct JNINativeInterface_ g_table = {m_FindClass, m_ThrowNew, m_GetArrayLength, m_NewStringUTF, m_GetCritical, m_ReleaseCritical};
static JNIEnv g_env = &g_table;

JNIEXPORT JNIEnv *mock_env(void) { return &g_env; }
JNIEXPORT jarray mock_array(void *data)
{
    struct _jobject *o = calloc(1, sizeof *o);
    o->kind = 1; o->data = data;
    return o;
}
JNIEXPORT jarray mock_array_n(void *data, int length)
{
    jarray o = mock_array(data);
    o->length = length;
    return o;
}
JNIEXPORT void mock_free(jobject o) { free(o); }
JNIEXPORT const char *mock_string(jstring s) { return s ? s->text : ""; }
JNIEXPORT const char *mock_exception_class(void) { return g_exClass; }
JNIEXPORT const char *mock_exception_message(void) { return g_exMsg; }
JNIEXPORT void mock_clear_exception(void) { g_exClass[0] = 0; g_exMsg[0] = 0; }
JNIEXPORT int mock_pins(void) { return g_pins; }
JNIEXPORT int mock_unpins(void) { return g_unpins; }
JNIEXPORT int mock_live_pins(void) { return g_live; }
