/*
 * bvg_jni.c -- JNI shim between it.unimi.dsi.webgraph.b200.B200BVGraph and the C ABI of libbvgraph_b200.so
 * (include/bvgraph_b200.h).  Build on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include bvg_jni.c \
 *       -L../../webgraph_b200 -lbvgraph_b200 -o libbvgraph_b200_jni.so
 * Not compiled in this repository's image (no jni.h); see INTEGRATION.md.
 * bvg_status -> exception mapping follows SURVEY 8b / the comments of enum bvg_status.
 *
 * Rules kept here: no JNI critical region is ever held across a call into the library (those calls launch kernels and
 * block on the device; a critical region would stall the collector); every allocation result is checked; the sequential
 * route hands whole decoded batches to Java as direct ByteBuffers over the cursor's pinned memory (no copy, no per-node
 * call).
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>
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
    jclass k = (*env)->FindClass(env, cls);
    if (k) (*env)->ThrowNew(env, k, bvg_strerror(rc));
}

#define G(h) ((bvg_graph*)(intptr_t)(h))
#define CUR(h) ((bvg_cursor*)(intptr_t)(h))
#define JNI(name) Java_it_unimi_dsi_webgraph_b200_B200BVGraph_##name

JNIEXPORT jlong JNICALL JNI(nativeOpen)(JNIEnv* env, jclass c, jstring basename, jint offsetType) {
    const char* b = (*env)->GetStringUTFChars(env, basename, NULL);
    if (!b) return 0;  /* OutOfMemoryError already pending */
    bvg_graph* g = NULL;
    int rc = bvg_open(b, offsetType, NULL, 0, &g);
    (*env)->ReleaseStringUTFChars(env, basename, b);
    if (rc) { throw_status(env, rc); return 0; }
    return (jlong)(intptr_t)g;
}
JNIEXPORT void JNICALL JNI(nativeClose)(JNIEnv* env, jclass c, jlong h) { if (h) bvg_close(G(h)); }
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
    if (!out) return NULL;  /* OutOfMemoryError pending */
    if (d) {
        int32_t* tmp = (int32_t*)malloc((size_t)d * sizeof(int32_t));  /* decoded outside any critical region, then copied in */
        if (!tmp) { throw_status(env, BVG_ENOMEM); return NULL; }
        rc = bvg_successors(G(h), x, tmp, d, &d);
        if (!rc) (*env)->SetIntArrayRegion(env, out, 0, d, (const jint*)tmp);
        free(tmp);
        if (rc) { throw_status(env, rc); return NULL; }
    }
    return out;
}

/* ---- sequential route: a cursor per NodeIterator; a batch is two direct buffers over the cursor's pinned memory ---- */
JNIEXPORT jlong JNICALL JNI(nativeCursorOpen)(JNIEnv* env, jclass c, jlong h, jint from, jint upper) {
    bvg_cursor* cur = NULL;
    int rc = bvg_cursor_open(G(h), from, upper, &cur);  /* BVG_ESTATE when from != 0 without offsets (BVGraph.java:1174) */
    if (rc) { throw_status(env, rc); return 0; }
    return (jlong)(intptr_t)cur;
}
JNIEXPORT void JNICALL JNI(nativeCursorClose)(JNIEnv* env, jclass c, jlong cur) { if (cur) bvg_cursor_close(CUR(cur)); }
/* meta[0] = first node, meta[1] = nodes in the batch; bufs[0] = offsets (count + 1 longs, native order), bufs[1] = successors
 * (ints, native order; indexed by the offsets, which are relative to the start of this buffer).  false at the end. */
JNIEXPORT jboolean JNICALL JNI(nativeCursorNextBatch)(JNIEnv* env, jclass c, jlong cur, jintArray meta, jobjectArray bufs) {
    int32_t first = 0, count = 0;
    const int64_t* off = NULL;
    const int32_t* succ = NULL;
    int rc = bvg_cursor_next_batch(CUR(cur), &first, &count, &off, &succ);
    if (rc == BVG_EEND) return JNI_FALSE;
    if (rc) { throw_status(env, rc); return JNI_FALSE; }
    const jint m[2] = { first, count };
    (*env)->SetIntArrayRegion(env, meta, 0, 2, m);
    jobject bo = (*env)->NewDirectByteBuffer(env, (void*)off, (jlong)(count + 1) * 8);
    /* the successor buffer is addressed by absolute offsets: expose everything up to the batch's last arc */
    jobject bs = (*env)->NewDirectByteBuffer(env, (void*)succ, (jlong)off[count] * 4);
    if (!bo || !bs) return JNI_FALSE;  /* OutOfMemoryError pending */
    (*env)->SetObjectArrayElement(env, bufs, 0, bo);
    (*env)->SetObjectArrayElement(env, bufs, 1, bs);
    return JNI_TRUE;
}
