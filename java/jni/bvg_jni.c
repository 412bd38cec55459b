/*
 * bvg_jni.c -- JNI shim between it.unimi.dsi.webgraph.b200.B200BVGraph and the C ABI of libbvgraph_b200.so
 * (include/bvgraph_b200.h).  Build on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include bvg_jni.c \
 *       -L../../webgraph_b200 -lbvgraph_b200 -o libbvgraph_b200_jni.so
 * Not compiled in this repository's image (no jni.h); see INTEGRATION.md.
 * bvg_status -> exception mapping follows SURVEY 8b / the comments of enum bvg_status.
 */
#include <jni.h>
#include <stdlib.h>
#include "bvgraph_b200.h"

static void throw_status(JNIEnv* env, int rc) {
    const char* cls;
    switch (rc) {
        case BVG_EINVAL: cls = "java/lang/IllegalArgumentException"; break;
        case BVG_ESTATE: cls = "java/lang/IllegalStateException"; break;
        case BVG_EUNSUPPORTED: cls = "java/lang/UnsupportedOperationException"; break;
        case BVG_EEND: cls = "java/util/NoSuchElementException"; break;
        case BVG_ENOMEM: cls = "java/lang/OutOfMemoryError"; break;
        case BVG_EIO: case BVG_EFORMAT: cls = "java/io/IOException"; break;
        default: cls = "java/lang/RuntimeException"; break;
    }
    (*env)->ThrowNew(env, (*env)->FindClass(env, cls), bvg_strerror(rc));
}

#define G(h) ((bvg_graph*)(intptr_t)(h))
#define JNI(name) Java_it_unimi_dsi_webgraph_b200_B200BVGraph_##name

JNIEXPORT jlong JNICALL JNI(nativeOpen)(JNIEnv* env, jclass c, jstring basename, jint offsetType) {
    const char* b = (*env)->GetStringUTFChars(env, basename, NULL);
    bvg_graph* g = NULL;
    int rc = bvg_open(b, offsetType, NULL, 0, &g);
    (*env)->ReleaseStringUTFChars(env, basename, b);
    if (rc) { throw_status(env, rc); return 0; }
    return (jlong)(intptr_t)g;
}
JNIEXPORT void JNICALL JNI(nativeClose)(JNIEnv* env, jclass c, jlong h) { bvg_close(G(h)); }
JNIEXPORT jint JNICALL JNI(nativeNumNodes)(JNIEnv* env, jclass c, jlong h) { int32_t n = 0; bvg_info(G(h), &n, 0, 0, 0, 0, 0, 0, 0); return n; }
JNIEXPORT jlong JNICALL JNI(nativeNumArcs)(JNIEnv* env, jclass c, jlong h) { int64_t m = 0; bvg_info(G(h), 0, &m, 0, 0, 0, 0, 0, 0); return m; }
JNIEXPORT jint JNICALL JNI(nativeOutdegree)(JNIEnv* env, jclass c, jlong h, jint x) {
    int32_t d = 0;
    int rc = bvg_outdegree(G(h), x, &d);
    if (rc) throw_status(env, rc);
    return d;
}
JNIEXPORT jintArray JNICALL JNI(nativeSuccessorArray)(JNIEnv* env, jclass c, jlong h, jint x) {
    int32_t d = 0;
    int rc = bvg_outdegree(G(h), x, &d);
    if (rc == BVG_ESTATE) rc = BVG_EUNSUPPORTED;  /* successors() without offsets is UOE (BVGraph.java:901) */
    if (rc) { throw_status(env, rc); return NULL; }
    jintArray out = (*env)->NewIntArray(env, d);
    if (d) {
        jint* p = (*env)->GetPrimitiveArrayCritical(env, out, NULL);
        rc = bvg_successors(G(h), x, (int32_t*)p, d, &d);
        (*env)->ReleasePrimitiveArrayCritical(env, out, p, 0);
        if (rc) { throw_status(env, rc); return NULL; }
    }
    return out;
}
JNIEXPORT jlong JNICALL JNI(nativeRangeArcs)(JNIEnv* env, jclass c, jlong h, jint from, jint to) {
    int64_t a = 0;
    int rc = bvg_range_arcs(G(h), from, to, &a);
    if (rc) throw_status(env, rc);
    return a;
}
JNIEXPORT jlongArray JNICALL JNI(nativeDecodeRange)(JNIEnv* env, jclass c, jlong h, jint from, jint to, jobjectArray succOut) {
    int64_t arcs = 0;
    int rc = bvg_range_arcs(G(h), from, to, &arcs);
    if (rc) { throw_status(env, rc); return NULL; }
    jlongArray off = (*env)->NewLongArray(env, to - from + 1);
    jintArray succ = (*env)->NewIntArray(env, (jsize)arcs);
    jlong* po = (*env)->GetPrimitiveArrayCritical(env, off, NULL);
    jint* ps = (*env)->GetPrimitiveArrayCritical(env, succ, NULL);
    rc = bvg_decode_range(G(h), from, to, (int64_t*)po, (int32_t*)ps, arcs, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, succ, ps, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, off, po, 0);
    if (rc) { throw_status(env, rc); return NULL; }
    (*env)->SetObjectArrayElement(env, succOut, 0, succ);
    return off;
}
