/* Minimal stand-in for <jni.h>: just enough of the JNI 1.6 surface for `gcc -fsyntax-only` on
 * kmers.anno_b200/java/jni/kmerengine_jni.c (tests/test_host.py).  There is no JDK in this image;
 * names, argument orders and types follow the JNI specification.  Test infrastructure only. */
#ifndef STUB_JNI_H
#define STUB_JNI_H
#include <stdint.h>
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef jint jsize;
typedef void* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jbyteArray;
typedef jobject jthrowable;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv*, const char*);
    jint (*ThrowNew)(JNIEnv*, jclass, const char*);
    jboolean (*ExceptionCheck)(JNIEnv*);
    jsize (*GetArrayLength)(JNIEnv*, jarray);
    jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*);
    void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
    jbyteArray (*NewByteArray)(JNIEnv*, jsize);
    void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
    jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
    void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
    jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
};
#endif
