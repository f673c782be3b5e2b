/*
 * jni.h -- MINIMAL STAND-IN for the JDK's <jni.h>, test infrastructure only (the build image has no JDK).
 *
 * It declares exactly the JNI types, constants and JNIEnv entries that java/jni/sb2gpu_jni.c uses, with the JDK's
 * spelling and calling shape (JNIEnv is a pointer to a table of function pointers that take the env as first argument),
 * so that the shim goes through a real C compiler (-Wall -Werror) and can be EXECUTED against a mock JNIEnv
 * (tests/jni_stub/mock_jvm.c).  It is NOT binary compatible with a JVM: the real table has ~230 slots in a fixed order.
 * On a machine with a JDK the shim is compiled against $JAVA_HOME/include/jni.h instead (java/README.md).
 */
#ifndef SB2_TEST_JNI_STUB_H
#define SB2_TEST_JNI_STUB_H

#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;

struct _jobject;
typedef struct _jobject *jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jthrowable;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jdoubleArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
    jclass (JNICALL *FindClass)(JNIEnv *env, const char *name);
    jint (JNICALL *ThrowNew)(JNIEnv *env, jclass clazz, const char *msg);
    jint (JNICALL *GetArrayLength)(JNIEnv *env, jarray array);
    jstring (JNICALL *NewStringUTF)(JNIEnv *env, const char *utf);
    void *(JNICALL *GetPrimitiveArrayCritical)(JNIEnv *env, jarray array, jboolean *isCopy);
    void (JNICALL *ReleasePrimitiveArrayCritical)(JNIEnv *env, jarray array, void *carray, jint mode);
};

#endif
