/* gkd_jni.c -- JNI glue between org.theseed.sequence.gpu.GkdNative and the C ABI of libgkd.so.
 * Build where a JDK exists (not in the build image):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       java/jni/gkd_jni.c -Lgenome/distance_b200 -lgkd -o libgkd_jni.so
 * No callbacks into the JVM are made from CUDA threads; arrays are pinned with
 * Get/ReleasePrimitiveArrayCritical only around the copying ABI call. */
#include <jni.h>
#include <stdint.h>
#include <stdlib.h>

#include "gkd.h"

#define CTX(h) ((gkd_ctx *)(intptr_t)(h))
#define FN(name) Java_org_theseed_sequence_gpu_GkdNative_##name

JNIEXPORT jlong JNICALL FN(create)(JNIEnv *env, jclass c, jint device, jint k, jint alphabet, jint strand) {
    gkd_config cfg = {0};
    cfg.device = device;
    cfg.k = k;
    cfg.alphabet = alphabet;
    cfg.strand_mode = strand;
    gkd_ctx *ctx = NULL;
    return gkd_create(&ctx, &cfg) == GKD_OK ? (jlong)(intptr_t)ctx : 0;
}

JNIEXPORT void JNICALL FN(destroy)(JNIEnv *env, jclass c, jlong h) { gkd_destroy(CTX(h)); }

JNIEXPORT jstring JNICALL FN(lastError)(JNIEnv *env, jclass c, jlong h) {
    return (*env)->NewStringUTF(env, gkd_last_error(CTX(h)));
}

JNIEXPORT jint JNICALL FN(addSequences)(JNIEnv *env, jclass c, jlong h, jobjectArray contigs) {
    jsize n = (*env)->GetArrayLength(env, contigs);
    const char **ptrs = (const char **)calloc((size_t)n + 1, sizeof(char *));
    uint64_t *lens = (uint64_t *)calloc((size_t)n + 1, sizeof(uint64_t));
    jbyteArray *arrs = (jbyteArray *)calloc((size_t)n + 1, sizeof(jbyteArray));
    if (!ptrs || !lens || !arrs) return GKD_ENOMEM;
    for (jsize i = 0; i < n; i++) {
        arrs[i] = (jbyteArray)(*env)->GetObjectArrayElement(env, contigs, i);
        lens[i] = (uint64_t)(*env)->GetArrayLength(env, arrs[i]);
        ptrs[i] = (const char *)(*env)->GetByteArrayElements(env, arrs[i], NULL);
    }
    uint32_t id = 0;
    int rc = gkd_add_sequences(CTX(h), ptrs, lens, (uint32_t)n, &id); /* copies before returning */
    for (jsize i = 0; i < n; i++) (*env)->ReleaseByteArrayElements(env, arrs[i], (jbyte *)ptrs[i], JNI_ABORT);
    free(ptrs);
    free(lens);
    free(arrs);
    return rc == GKD_OK ? (jint)id : rc;
}

JNIEXPORT jintArray JNICALL FN(addFastaFile)(JNIEnv *env, jclass c, jlong h, jstring path, jboolean perRecord) {
    const char *p = (*env)->GetStringUTFChars(env, path, NULL);
    uint32_t first = 0, n = 0;
    int rc = gkd_add_fasta_file(CTX(h), p, perRecord ? 1 : 0, &first, &n);
    (*env)->ReleaseStringUTFChars(env, path, p);
    if (rc != GKD_OK) return NULL;
    jint out[2] = {(jint)first, (jint)n};
    jintArray r = (*env)->NewIntArray(env, 2);
    (*env)->SetIntArrayRegion(env, r, 0, 2, out);
    return r;
}

JNIEXPORT jstring JNICALL FN(label)(JNIEnv *env, jclass c, jlong h, jint id) {
    return (*env)->NewStringUTF(env, gkd_label(CTX(h), (uint32_t)id));
}
JNIEXPORT jstring JNICALL FN(comment)(JNIEnv *env, jclass c, jlong h, jint id) {
    return (*env)->NewStringUTF(env, gkd_comment(CTX(h), (uint32_t)id));
}

JNIEXPORT jint JNICALL FN(buildSets)(JNIEnv *env, jclass c, jlong h) { return gkd_build_sets(CTX(h)); }

JNIEXPORT jint JNICALL FN(allVsAll)(JNIEnv *env, jclass c, jlong h, jdoubleArray dist) {
    jdouble *d = (*env)->GetDoubleArrayElements(env, dist, NULL);
    int rc = gkd_all_vs_all(CTX(h), NULL, d);
    (*env)->ReleaseDoubleArrayElements(env, dist, d, 0);
    return rc;
}

JNIEXPORT jint JNICALL FN(queryVsRef)(JNIEnv *env, jclass c, jlong h, jintArray q, jintArray r, jdoubleArray dist) {
    jsize nq = (*env)->GetArrayLength(env, q), nr = (*env)->GetArrayLength(env, r);
    jint *qa = (*env)->GetIntArrayElements(env, q, NULL);
    jint *ra = (*env)->GetIntArrayElements(env, r, NULL);
    jdouble *d = (*env)->GetDoubleArrayElements(env, dist, NULL);
    int rc = gkd_query_vs_ref(CTX(h), (const uint32_t *)qa, (uint32_t)nq, (const uint32_t *)ra, (uint32_t)nr, NULL, d);
    (*env)->ReleaseDoubleArrayElements(env, dist, d, 0);
    (*env)->ReleaseIntArrayElements(env, r, ra, JNI_ABORT);
    (*env)->ReleaseIntArrayElements(env, q, qa, JNI_ABORT);
    return rc;
}

JNIEXPORT jdouble JNICALL FN(pairDistance)(JNIEnv *env, jclass c, jlong h, jint a, jint b) {
    double d = 1.0;
    gkd_pair(CTX(h), (uint32_t)a, (uint32_t)b, NULL, NULL, &d);
    return d;
}

JNIEXPORT jlong JNICALL FN(setSize)(JNIEnv *env, jclass c, jlong h, jint id) {
    uint64_t n = 0;
    gkd_set_size(CTX(h), (uint32_t)id, &n, NULL, NULL);
    return (jlong)n;
}
