/*
 * kmerengine_jni.c — JNI shim between org.theseed.proteins.kmers.gpu.KmerEngine and the C ABI
 * of libkmeranno.so (include/kmeranno.h).  NOT BUILT HERE (no JDK / jni.h in this image):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       kmerengine_jni.c -Lkmers.anno_b200 -lkmeranno -o libkmerengine_jni.so
 * Error convention: a negative ka_* code is thrown as java.io.IOException with ka_last_error.
 */
#include <jni.h>
#include <stdint.h>

#include "kmeranno.h"

static void throw_io(JNIEnv* env, int code, const char* msg) {
    char buf[600];
    snprintf(buf, sizeof buf, "kmeranno error %d: %s", code, msg ? msg : "");
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/io/IOException"), buf);
}

JNIEXPORT jlong JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_create(JNIEnv* env, jclass cls, jintArray devices) {
    jsize n = (*env)->GetArrayLength(env, devices);
    jint* d = (*env)->GetIntArrayElements(env, devices, NULL);
    ka_engine* e = NULL;
    int rc = ka_create((const int*)d, (int)n, &e);
    (*env)->ReleaseIntArrayElements(env, devices, d, JNI_ABORT);
    if (rc != KA_OK) { throw_io(env, rc, ka_last_error(NULL)); return 0; }
    return (jlong)(intptr_t)e;
}

JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_destroy(JNIEnv* env, jclass cls, jlong h) {
    ka_destroy((ka_engine*)(intptr_t)h);
}

JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_dbLoad(JNIEnv* env, jclass cls, jlong h,
        jbyteArray kmers, jintArray roles, jlong n, jint k) {
    ka_engine* e = (ka_engine*)(intptr_t)h;
    jbyte* km = (*env)->GetPrimitiveArrayCritical(env, kmers, NULL);
    jint* ro = (*env)->GetPrimitiveArrayCritical(env, roles, NULL);
    int rc = ka_db_load(e, (const uint8_t*)km, (const int32_t*)ro, (uint64_t)n, (int)k);
    (*env)->ReleasePrimitiveArrayCritical(env, roles, ro, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, kmers, km, JNI_ABORT);
    if (rc != KA_OK) throw_io(env, rc, ka_last_error(e));
}

JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_annotate(JNIEnv* env, jclass cls, jlong h,
        jbyteArray residues, jlongArray offsets, jlong n, jint minHits, jintArray role, jintArray hits, jbyteArray flag) {
    ka_engine* e = (ka_engine*)(intptr_t)h;
    jbyte* res = (*env)->GetPrimitiveArrayCritical(env, residues, NULL);
    jlong* off = (*env)->GetPrimitiveArrayCritical(env, offsets, NULL);
    jint* ro = (*env)->GetPrimitiveArrayCritical(env, role, NULL);
    jint* hi = (*env)->GetPrimitiveArrayCritical(env, hits, NULL);
    jbyte* fl = (*env)->GetPrimitiveArrayCritical(env, flag, NULL);
    int rc = ka_annotate(e, (const uint8_t*)res, (const uint64_t*)off, (uint64_t)n, (int32_t)minHits,
                         (int32_t*)ro, (int32_t*)hi, (uint8_t*)fl);
    (*env)->ReleasePrimitiveArrayCritical(env, flag, fl, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, hits, hi, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, role, ro, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, offsets, off, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, residues, res, JNI_ABORT);
    if (rc != KA_OK) throw_io(env, rc, ka_last_error(e));
}

JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_kmerDistance(JNIEnv* env, jclass cls, jlong h,
        jbyteArray residues, jlongArray offsets, jlong n, jint k, jintArray querySeq, jlongArray groupOff, jlong q,
        jintArray cand, jintArray common, jdoubleArray dist) {
    ka_engine* e = (ka_engine*)(intptr_t)h;
    jbyte* res = (*env)->GetPrimitiveArrayCritical(env, residues, NULL);
    jlong* off = (*env)->GetPrimitiveArrayCritical(env, offsets, NULL);
    jint* qs = (*env)->GetPrimitiveArrayCritical(env, querySeq, NULL);
    jlong* go = (*env)->GetPrimitiveArrayCritical(env, groupOff, NULL);
    jint* cs = (*env)->GetPrimitiveArrayCritical(env, cand, NULL);
    jint* co = (*env)->GetPrimitiveArrayCritical(env, common, NULL);
    jdouble* di = (*env)->GetPrimitiveArrayCritical(env, dist, NULL);
    int rc = ka_kmer_distance(e, (const uint8_t*)res, (const uint64_t*)off, (uint64_t)n, (int)k, (const uint32_t*)qs,
                              (const uint64_t*)go, (uint64_t)q, (const uint32_t*)cs, NULL, (int32_t*)co, (double*)di);
    (*env)->ReleasePrimitiveArrayCritical(env, dist, di, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, common, co, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, cand, cs, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, groupOff, go, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, querySeq, qs, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, offsets, off, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, residues, res, JNI_ABORT);
    if (rc != KA_OK) throw_io(env, rc, ka_last_error(e));
}
