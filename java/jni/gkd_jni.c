/* gkd_jni.c -- JNI glue between org.theseed.sequence.gpu.GkdNative and the C ABI of libgkd.so.
 * Build where a JDK exists (not in the build image; there it is syntax-checked against tests/stubs/jni.h):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude \
 *       java/jni/gkd_jni.c -Lgenome/distance_b200 -lgkd -o libgkd_jni.so
 * No callbacks into the JVM are made from CUDA threads.  Every Java array that a native call fills is
 * length-checked against what the call writes; status codes are returned, never swallowed. */
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
    (void)env;
    (void)c;
    return gkd_create(&ctx, &cfg) == GKD_OK ? (jlong)(intptr_t)ctx : 0;
}

JNIEXPORT void JNICALL FN(destroy)(JNIEnv *env, jclass c, jlong h) {
    (void)env;
    (void)c;
    gkd_destroy(CTX(h));
}

JNIEXPORT jstring JNICALL FN(lastError)(JNIEnv *env, jclass c, jlong h) {
    (void)c;
    return (*env)->NewStringUTF(env, gkd_last_error(CTX(h)));
}

JNIEXPORT jint JNICALL FN(addSequences)(JNIEnv *env, jclass c, jlong h, jobjectArray contigs) {
    (void)c;
    jsize n = (*env)->GetArrayLength(env, contigs);
    const char **ptrs = (const char **)calloc((size_t)n + 1, sizeof(char *));
    uint64_t *lens = (uint64_t *)calloc((size_t)n + 1, sizeof(uint64_t));
    jbyteArray *arrs = (jbyteArray *)calloc((size_t)n + 1, sizeof(jbyteArray));
    int rc = GKD_OK;
    uint32_t id = 0;
    jsize got = 0;
    if (!ptrs || !lens || !arrs) rc = GKD_ENOMEM;
    for (; rc == GKD_OK && got < n; got++) {
        arrs[got] = (jbyteArray)(*env)->GetObjectArrayElement(env, contigs, got);
        lens[got] = (uint64_t)(*env)->GetArrayLength(env, arrs[got]);
        ptrs[got] = (const char *)(*env)->GetByteArrayElements(env, arrs[got], NULL);
        if (!ptrs[got]) {
            rc = GKD_ENOMEM;
            break;
        }
    }
    /* the JVM's byte arrays are pageable memory: the library copies them before it returns */
    if (rc == GKD_OK) rc = gkd_add_sequences(CTX(h), ptrs, lens, (uint32_t)n, &id);
    for (jsize i = 0; i < got; i++) (*env)->ReleaseByteArrayElements(env, arrs[i], (jbyte *)ptrs[i], JNI_ABORT);
    free(ptrs);
    free(lens);
    free(arrs);
    return rc == GKD_OK ? (jint)id : rc;
}

JNIEXPORT jintArray JNICALL FN(addFastaFile)(JNIEnv *env, jclass c, jlong h, jstring path, jboolean perRecord) {
    (void)c;
    const char *p = (*env)->GetStringUTFChars(env, path, NULL);
    if (!p) return NULL;
    uint32_t first = 0, n = 0;
    int rc = gkd_add_fasta_file(CTX(h), p, perRecord ? 1 : 0, &first, &n);
    (*env)->ReleaseStringUTFChars(env, path, p);
    if (rc != GKD_OK) return NULL;
    jint out[2] = {(jint)first, (jint)n};
    jintArray r = (*env)->NewIntArray(env, 2);
    if (r) (*env)->SetIntArrayRegion(env, r, 0, 2, out);
    return r;
}

JNIEXPORT jstring JNICALL FN(label)(JNIEnv *env, jclass c, jlong h, jint id) {
    (void)c;
    return (*env)->NewStringUTF(env, gkd_label(CTX(h), (uint32_t)id));
}
JNIEXPORT jstring JNICALL FN(comment)(JNIEnv *env, jclass c, jlong h, jint id) {
    (void)c;
    return (*env)->NewStringUTF(env, gkd_comment(CTX(h), (uint32_t)id));
}

JNIEXPORT jint JNICALL FN(count)(JNIEnv *env, jclass c, jlong h) {
    (void)env;
    (void)c;
    return (jint)gkd_count(CTX(h));
}

JNIEXPORT jint JNICALL FN(buildSets)(JNIEnv *env, jclass c, jlong h) {
    (void)env;
    (void)c;
    return gkd_build_sets(CTX(h));
}

JNIEXPORT jint JNICALL FN(truncate)(JNIEnv *env, jclass c, jlong h, jint keep) {
    (void)env;
    (void)c;
    return keep < 0 ? GKD_EINVAL : gkd_truncate(CTX(h), (uint32_t)keep);
}

JNIEXPORT jint JNICALL FN(allVsAllRange)(JNIEnv *env, jclass c, jlong h, jint n, jlong first, jlong count,
                                         jdoubleArray dist) {
    (void)c;
    if (n < 0 || first < 0 || count < 0 || !dist) return GKD_EINVAL;
    /* the native call writes exactly `count` doubles: the Java array must hold them */
    if ((jlong)(*env)->GetArrayLength(env, dist) < count) return GKD_EINVAL;
    jdouble *d = (*env)->GetDoubleArrayElements(env, dist, NULL);
    if (!d) return GKD_ENOMEM;
    int rc = gkd_all_vs_all_range(CTX(h), (uint32_t)n, (uint64_t)first, (uint64_t)count, NULL, d);
    (*env)->ReleaseDoubleArrayElements(env, dist, d, 0);
    return rc;
}

JNIEXPORT jint JNICALL FN(queryVsRef)(JNIEnv *env, jclass c, jlong h, jintArray q, jintArray r, jdoubleArray dist) {
    (void)c;
    if (!q || !r || !dist) return GKD_EINVAL;
    jsize nq = (*env)->GetArrayLength(env, q), nr = (*env)->GetArrayLength(env, r);
    if ((jlong)(*env)->GetArrayLength(env, dist) < (jlong)nq * (jlong)nr) return GKD_EINVAL;
    jint *qa = (*env)->GetIntArrayElements(env, q, NULL);
    jint *ra = (*env)->GetIntArrayElements(env, r, NULL);
    jdouble *d = (*env)->GetDoubleArrayElements(env, dist, NULL);
    int rc = (qa && ra && d) ? gkd_query_vs_ref(CTX(h), (const uint32_t *)qa, (uint32_t)nq, (const uint32_t *)ra,
                                                (uint32_t)nr, NULL, d)
                             : GKD_ENOMEM;
    if (d) (*env)->ReleaseDoubleArrayElements(env, dist, d, 0);
    if (ra) (*env)->ReleaseIntArrayElements(env, r, ra, JNI_ABORT);
    if (qa) (*env)->ReleaseIntArrayElements(env, q, qa, JNI_ABORT);
    return rc;
}

JNIEXPORT jint JNICALL FN(pair)(JNIEnv *env, jclass c, jlong h, jint a, jint b, jlongArray inter, jdoubleArray dist) {
    (void)c;
    if (a < 0 || b < 0) return GKD_EINVAL;
    if ((inter && (*env)->GetArrayLength(env, inter) < 1) || (dist && (*env)->GetArrayLength(env, dist) < 1))
        return GKD_EINVAL;
    uint64_t I = 0;
    double d = 1.0;
    int rc = gkd_pair(CTX(h), (uint32_t)a, (uint32_t)b, &I, NULL, &d);
    if (rc != GKD_OK) return rc; /* a failed call is an error, never "distance 1.0" */
    if (inter) {
        jlong v = (jlong)I;
        (*env)->SetLongArrayRegion(env, inter, 0, 1, &v);
    }
    if (dist) (*env)->SetDoubleArrayRegion(env, dist, 0, 1, &d);
    return GKD_OK;
}

JNIEXPORT jint JNICALL FN(setSize)(JNIEnv *env, jclass c, jlong h, jint id, jlongArray out) {
    (void)c;
    if (id < 0 || !out || (*env)->GetArrayLength(env, out) < 1) return GKD_EINVAL;
    uint64_t n = 0;
    int rc = gkd_set_size(CTX(h), (uint32_t)id, &n, NULL, NULL);
    if (rc != GKD_OK) return rc;
    jlong v = (jlong)n;
    (*env)->SetLongArrayRegion(env, out, 0, 1, &v);
    return GKD_OK;
}
