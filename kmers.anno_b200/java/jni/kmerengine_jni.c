/*
 * kmerengine_jni.c — JNI shim between org.theseed.proteins.kmers.gpu.KmerEngine and the C ABI
 * of libkmeranno.so (include/kmeranno.h).  No JDK exists in the authoring image: the file is
 * syntax-checked against tests/stub_jni/jni.h (tests/test_host.py) and built on a real system with
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       kmerengine_jni.c -Lkmers.anno_b200 -lkmeranno -o libkmerengine_jni.so
 *
 * Every bulk argument is a DIRECT ByteBuffer over pinned host memory handed out by allocPinned()
 * (ka_host_alloc): the GPU calls run for milliseconds to seconds, so no primitive-array critical
 * section is ever held across them (that would stall the JVM's collector), and the engine copies
 * straight from the buffer the Java side filled — no staging copy.
 * Error convention: a negative ka_* code is thrown as java.io.IOException with ka_last_error.
 */
#include <jni.h>
#include <stdint.h>
#include <stdio.h>

#include "kmeranno.h"

static void throw_io(JNIEnv* env, int code, const char* msg) {
    char buf[600];
    jclass cls;
    snprintf(buf, sizeof buf, "kmeranno error %d: %s", code, msg ? msg : "");
    cls = (*env)->FindClass(env, "java/io/IOException");
    if (cls) (*env)->ThrowNew(env, cls, buf);
}

/* address of a direct buffer that must hold at least `need` bytes; NULL after throwing */
static void* direct(JNIEnv* env, jobject buf, jlong need, const char* what) {
    void* p = buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL;
    if (!p || (*env)->GetDirectBufferCapacity(env, buf) < need) {
        char msg[160];
        snprintf(msg, sizeof msg, "%s must be a direct ByteBuffer of at least %lld bytes (KmerEngine.allocPinned)", what, (long long)need);
        throw_io(env, KA_ERR_INVALID, msg);
        return NULL;
    }
    return p;
}

JNIEXPORT jlong JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_create(JNIEnv* env, jclass cls, jintArray devices) {
    jsize n = (*env)->GetArrayLength(env, devices);
    jint* d = (*env)->GetIntArrayElements(env, devices, NULL);
    ka_engine* e = NULL;
    int rc;
    (void)cls;
    if (!d) return 0;                                   /* OutOfMemoryError already pending */
    rc = ka_create((const int*)d, (int)n, &e);
    (*env)->ReleaseIntArrayElements(env, devices, d, JNI_ABORT);
    if (rc != KA_OK) { throw_io(env, rc, ka_last_error(NULL)); return 0; }
    return (jlong)(intptr_t)e;
}

JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_destroy(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    ka_destroy((ka_engine*)(intptr_t)h);
}

JNIEXPORT jobject JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_allocPinned(JNIEnv* env, jclass cls, jlong bytes) {
    void* p = ka_host_alloc((size_t)bytes);
    (void)cls;
    if (!p) { throw_io(env, KA_ERR_OOM, "ka_host_alloc failed"); return NULL; }
    return (*env)->NewDirectByteBuffer(env, p, bytes);
}

JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_freePinned(JNIEnv* env, jclass cls, jobject buf) {
    (void)cls;
    if (buf) ka_host_free((*env)->GetDirectBufferAddress(env, buf));
}

JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_dbLoad(JNIEnv* env, jclass cls, jlong h,
        jobject kmers, jobject roles, jlong n, jint k) {
    ka_engine* e = (ka_engine*)(intptr_t)h;
    const uint8_t* km = direct(env, kmers, n * k, "kmers");
    const int32_t* ro = km ? direct(env, roles, n * 4, "roles") : NULL;
    int rc;
    (void)cls;
    if (!ro) return;
    rc = ka_db_load(e, km, ro, (uint64_t)n, (int)k);
    if (rc != KA_OK) throw_io(env, rc, ka_last_error(e));
}

/* code_of_byte[256] of the loaded DB, for the Java-side packer */
JNIEXPORT jbyteArray JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_alphabet(JNIEnv* env, jclass cls, jlong h) {
    ka_engine* e = (ka_engine*)(intptr_t)h;
    uint8_t lut[256];
    jbyteArray out;
    int rc = ka_db_get_alphabet(e, lut);
    (void)cls;
    if (rc != KA_OK) { throw_io(env, rc, ka_last_error(e)); return NULL; }
    out = (*env)->NewByteArray(env, 256);
    if (out) (*env)->SetByteArrayRegion(env, out, 0, 256, (const jbyte*)lut);
    return out;
}

/* ka_annotate_packed: codes = 5-bit stream, offsets = int32 little-endian [n + 1], results int32 / int32 / byte [n] */
JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_annotatePacked(JNIEnv* env, jclass cls, jlong h,
        jobject codes, jlong codeBytes, jobject offsets, jlong n, jint minHits, jobject role, jobject hits, jobject flag) {
    ka_engine* e = (ka_engine*)(intptr_t)h;
    const uint8_t* co = direct(env, codes, codeBytes, "codes");
    const uint32_t* of = co ? direct(env, offsets, (n + 1) * 4, "offsets") : NULL;
    int32_t* ro = of ? direct(env, role, n * 4, "role") : NULL;
    int32_t* hi = ro ? direct(env, hits, n * 4, "hits") : NULL;
    uint8_t* fl = hi ? direct(env, flag, n, "flag") : NULL;
    int rc;
    (void)cls;
    if (!fl) return;
    rc = ka_annotate_packed(e, co, of, (uint64_t)n, (int32_t)minHits, ro, hi, fl);
    if (rc != KA_OK) throw_io(env, rc, ka_last_error(e));
}

/* ka_annotate: residues = one byte per residue, offsets = int64 little-endian [n + 1] */
JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_annotate(JNIEnv* env, jclass cls, jlong h,
        jobject residues, jlong residueBytes, jobject offsets, jlong n, jint minHits, jobject role, jobject hits, jobject flag) {
    ka_engine* e = (ka_engine*)(intptr_t)h;
    const uint8_t* re = direct(env, residues, residueBytes, "residues");
    const uint64_t* of = re ? direct(env, offsets, (n + 1) * 8, "offsets") : NULL;
    int32_t* ro = of ? direct(env, role, n * 4, "role") : NULL;
    int32_t* hi = ro ? direct(env, hits, n * 4, "hits") : NULL;
    uint8_t* fl = hi ? direct(env, flag, n, "flag") : NULL;
    int rc;
    (void)cls;
    if (!fl) return;
    rc = ka_annotate(e, re, of, (uint64_t)n, (int32_t)minHits, ro, hi, fl);
    if (rc != KA_OK) throw_io(env, rc, ka_last_error(e));
}

/* ka_kmer_distance: all arrays as direct buffers (int32 / int64 / float64 little-endian) */
JNIEXPORT void JNICALL Java_org_theseed_proteins_kmers_gpu_KmerEngine_kmerDistance(JNIEnv* env, jclass cls, jlong h,
        jobject residues, jlong residueBytes, jobject offsets, jlong n, jint k, jobject querySeq, jobject groupOff, jlong q,
        jobject cand, jlong m, jobject common, jobject dist) {
    ka_engine* e = (ka_engine*)(intptr_t)h;
    const uint8_t* re = direct(env, residues, residueBytes, "residues");
    const uint64_t* of = re ? direct(env, offsets, (n + 1) * 8, "offsets") : NULL;
    const uint32_t* qs = of ? direct(env, querySeq, q * 4, "querySeq") : NULL;
    const uint64_t* go = qs ? direct(env, groupOff, (q + 1) * 8, "groupOff") : NULL;
    const uint32_t* cs = go ? direct(env, cand, m * 4, "cand") : NULL;
    int32_t* cm = cs ? direct(env, common, m * 4, "common") : NULL;
    double* di = cm ? direct(env, dist, m * 8, "dist") : NULL;
    int rc;
    (void)cls;
    if (!di) return;
    rc = ka_kmer_distance(e, re, of, (uint64_t)n, (int)k, qs, go, (uint64_t)q, cs, NULL, cm, di);
    if (rc != KA_OK) throw_io(env, rc, ka_last_error(e));
}
