/*
 * rnabloom_jni.c -- JNI shim between rnabloom.gpu.Native (java/rnabloom/gpu/Native.java) and librnabloom_gpu.so.
 *
 * Build (needs a JDK; none exists in the build image, so this file is only syntax-checked there against tests/jni_stub/jni.h):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/rnabloom_jni.c \
 *       -Lrna-bloom_b200 -lrnabloom_gpu -o librnabloom_jni.so
 * Handles travel as jlong; bulk data as direct ByteBuffers (GetDirectBufferAddress: no copies, may be cudaHostRegister'ed).
 * A non-zero C-ABI return becomes a RuntimeException, matching how the reference's workers fail (RNABloom.java:631-633).
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include "rnabloom_gpu.h"

static void* addr(JNIEnv* env, jobject buf) { return buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL; }

static jint check(JNIEnv* env, rb_ctx* ctx, int32_t rc) {
    if (rc != RB_OK) {
        jclass ex = (*env)->FindClass(env, "java/lang/RuntimeException");
        if (ex) (*env)->ThrowNew(env, ex, rb_last_error(ctx));
    }
    return rc;
}

JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_ctxCreate(JNIEnv* env, jclass cls, jint device) {
    rb_ctx* ctx = NULL;
    (void)cls;
    check(env, NULL, rb_ctx_create(device, &ctx));
    return (jlong)(intptr_t)ctx;
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_ctxDestroy(JNIEnv* env, jclass cls, jlong ctx) {
    (void)env; (void)cls;
    rb_ctx_destroy((rb_ctx*)(intptr_t)ctx);
}
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_graphCreate(JNIEnv* env, jclass cls, jlong ctx, jlong dbgbfNumBits, jlong cbfNumBytes,
                                                            jlong pkbfNumBits, jint dbgbfNumHash, jint cbfNumHash, jint pkbfNumHash, jint k,
                                                            jboolean stranded, jboolean useReadPairedKmers) {
    rb_graph* g = NULL;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_create((rb_ctx*)(intptr_t)ctx, dbgbfNumBits, cbfNumBytes, pkbfNumBits, dbgbfNumHash, cbfNumHash,
                                                       pkbfNumHash, k, stranded, useReadPairedKmers, &g));
    return (jlong)(intptr_t)g;
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphDestroy(JNIEnv* env, jclass cls, jlong g) {
    (void)env; (void)cls;
    rb_graph_destroy((rb_graph*)(intptr_t)g);
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphSetDistances(JNIEnv* env, jclass cls, jlong ctx, jlong g, jint dRead, jint dFrag) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_set_distances((rb_graph*)(intptr_t)g, dRead, dFrag));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphSetEngine(JNIEnv* env, jclass cls, jlong ctx, jlong g, jint engine) {
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_set_engine((rb_graph*)(intptr_t)g, engine));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphInitFpkbf(JNIEnv* env, jclass cls, jlong ctx, jlong g, jlong bits, jint numHash) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_init_fpkbf((rb_graph*)(intptr_t)g, bits, numHash));
}
/* FastqToGraphWorker / FastaToGraphWorker body for a chunk of records (RNABloom.java:551-634, 677-716) */
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_graphAddReadsAscii(JNIEnv* env, jclass cls, jlong ctx, jlong g, jobject bases, jobject quals,
                                                                   jobject offsets, jlong nReads, jint minQual, jint flags) {
    int64_t n = 0;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_add_reads_ascii((rb_graph*)(intptr_t)g, (const char*)addr(env, bases), (const char*)addr(env, quals),
                                                                (const int64_t*)addr(env, offsets), nReads, minQual, (uint32_t)flags, &n));
    return n;
}
/* FragmentsToGraphWorker body: already 2-bit packed fragments (RNABloom.java:1489-1516) */
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_graphAddReadsPacked(JNIEnv* env, jclass cls, jlong ctx, jlong g, jobject packed, jobject mask,
                                                                    jobject readOff, jobject readLen, jlong nReads, jint flags) {
    int64_t n = 0;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_add_reads((rb_graph*)(intptr_t)g, (const uint64_t*)addr(env, packed), (const uint32_t*)addr(env, mask),
                                                          (const int64_t*)addr(env, readOff), (const int32_t*)addr(env, readLen), nReads, 0, 0,
                                                          (uint32_t)flags, &n));
    return n;
}
/* graph.getKmers over many sequences (graph/BloomFilterDeBruijnGraph.java:1224-1226) */
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_graphCountReads(JNIEnv* env, jclass cls, jlong ctx, jlong g, jobject packed, jobject mask,
                                                                jobject readOff, jobject readLen, jlong nReads, jobject counts, jobject fHash,
                                                                jobject rHash) {
    int64_t n = 0;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_count_reads((rb_graph*)(intptr_t)g, (const uint64_t*)addr(env, packed), (const uint32_t*)addr(env, mask),
                                                            (const int64_t*)addr(env, readOff), (const int32_t*)addr(env, readLen), nReads, 0, 0,
                                                            (float*)addr(env, counts), (int64_t*)addr(env, fHash), (int64_t*)addr(env, rHash), &n));
    return n;
}
/* the same without waiting for the results (rb_graph_count_reads_async): returns the ticket for ctxWait; the buffers must be direct
 * ByteBuffers over pinned memory and stay untouched until the wait returns */
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_graphCountReadsAsync(JNIEnv* env, jclass cls, jlong ctx, jlong g, jobject packed, jobject mask,
                                                                     jobject readOff, jobject readLen, jlong nReads, jobject counts, jobject fHash,
                                                                     jobject rHash) {
    int64_t n = 0, ticket = 0;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_count_reads_async((rb_graph*)(intptr_t)g, (const uint64_t*)addr(env, packed), (const uint32_t*)addr(env, mask),
                                                                  (const int64_t*)addr(env, readOff), (const int32_t*)addr(env, readLen), nReads, 0, 0,
                                                                  (float*)addr(env, counts), (int64_t*)addr(env, fHash), (int64_t*)addr(env, rHash), &n, &ticket));
    return ticket;
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_ctxWait(JNIEnv* env, jclass cls, jlong ctx, jlong ticket) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_ctx_wait((rb_ctx*)(intptr_t)ctx, ticket));
}
/* graph.getKmers(String) over a chunk of sequences: counts + hashes, bit-exact for every byte value (graph :1224-1234) */
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_graphCountReadsAscii(JNIEnv* env, jclass cls, jlong ctx, jlong g, jobject bases, jobject offsets, jlong nReads,
                                                                     jobject counts, jobject fHash, jobject rHash) {
    int64_t n = 0;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_count_reads_ascii((rb_graph*)(intptr_t)g, (const char*)addr(env, bases), (const int64_t*)addr(env, offsets), nReads,
                                                                  (float*)addr(env, counts), (int64_t*)addr(env, fHash), (int64_t*)addr(env, rHash), &n));
    return n;
}
/* graph.add(long[]) / addCountIfPresent / addDbgOnly for an array of hVals[0] (graph :405-436) */
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphAddHashes(JNIEnv* env, jclass cls, jlong ctx, jlong g, jobject hashes, jlong n, jint flags) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_add_hashes((rb_graph*)(intptr_t)g, (const int64_t*)addr(env, hashes), n, (uint32_t)flags));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphCountHashes(JNIEnv* env, jclass cls, jlong ctx, jlong g, jobject hashes, jlong n, jobject counts) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_count_hashes((rb_graph*)(intptr_t)g, (const int64_t*)addr(env, hashes), n, (float*)addr(env, counts)));
}
/* Kmer.getSuccessors / getPredecessors for a batch of k-mers (graph/Kmer.java:213-253): counts (and hashes) of the 4 + 4 candidates */
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphNeighborCounts(JNIEnv* env, jclass cls, jlong ctx, jlong g, jobject fhash, jobject rhash, jobject firstBases,
                                                                    jobject lastBases, jlong n, jobject counts, jobject nbrF, jobject nbrR) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx,
          rb_graph_neighbor_counts((rb_graph*)(intptr_t)g, (const int64_t*)addr(env, fhash), (const int64_t*)addr(env, rhash), (const uint8_t*)addr(env, firstBases),
                                   (const uint8_t*)addr(env, lastBases), n, (float*)addr(env, counts), (int64_t*)addr(env, nbrF), (int64_t*)addr(env, nbrR)));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphAddPairHashes(JNIEnv* env, jclass cls, jlong ctx, jlong g, jint which, jobject hashes, jlong n) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_add_pair_hashes((rb_graph*)(intptr_t)g, which, (const int64_t*)addr(env, hashes), n));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphLookupPairHashes(JNIEnv* env, jclass cls, jlong ctx, jlong g, jint which, jobject hashes, jlong n,
                                                                     jobject out) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_lookup_pair_hashes((rb_graph*)(intptr_t)g, which, (const int64_t*)addr(env, hashes), n, (uint8_t*)addr(env, out)));
}
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_graphFilter(JNIEnv* env, jclass cls, jlong ctx, jlong g, jint which) {
    rb_filter* f = NULL;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_filter((rb_graph*)(intptr_t)g, which, &f));
    return (jlong)(intptr_t)f;
}
/* barrier + host mirror refresh: destinations are the native addresses of the inherited filters' Unsafe buffers (0 = skip) */
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphSyncToHost(JNIEnv* env, jclass cls, jlong ctx, jlong g, jlong dbgbfAddr, jlong cbfAddr, jlong rpkbfAddr,
                                                                jlong fpkbfAddr) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_sync_to_host((rb_graph*)(intptr_t)g, (void*)(intptr_t)dbgbfAddr, (void*)(intptr_t)cbfAddr,
                                                             (void*)(intptr_t)rpkbfAddr, (void*)(intptr_t)fpkbfAddr));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphSync(JNIEnv* env, jclass cls, jlong ctx, jlong g) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_sync((rb_graph*)(intptr_t)g));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_graphSave(JNIEnv* env, jclass cls, jlong ctx, jlong g, jstring path) {
    const char* p = (*env)->GetStringUTFChars(env, path, NULL);
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_save((rb_graph*)(intptr_t)g, p));
    (*env)->ReleaseStringUTFChars(env, path, p);
}
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_graphLoad(JNIEnv* env, jclass cls, jlong ctx, jstring path, jboolean loadDbgbf, jboolean loadFpkbf) {
    rb_graph* g = NULL;
    const char* p = (*env)->GetStringUTFChars(env, path, NULL);
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_graph_load((rb_ctx*)(intptr_t)ctx, p, loadDbgbf, loadFpkbf, &g));
    (*env)->ReleaseStringUTFChars(env, path, p);
    return (jlong)(intptr_t)g;
}
/* stand-alone filters: BloomFilter / CountingBloomFilter (bloom/BloomFilter.java, bloom/CountingBloomFilter.java) */
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_filterCreate(JNIEnv* env, jclass cls, jlong ctx, jint kind, jlong size, jint numHash, jint k) {
    rb_filter* f = NULL;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_create((rb_ctx*)(intptr_t)ctx, kind, size, numHash, k, &f));
    return (jlong)(intptr_t)f;
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_filterDestroy(JNIEnv* env, jclass cls, jlong f) {
    (void)env; (void)cls;
    rb_filter_destroy((rb_filter*)(intptr_t)f);
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_filterEmpty(JNIEnv* env, jclass cls, jlong ctx, jlong f) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_empty((rb_filter*)(intptr_t)f));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_filterAddHashes(JNIEnv* env, jclass cls, jlong ctx, jlong f, jobject hashes, jlong n) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_add_hashes((rb_filter*)(intptr_t)f, (const int64_t*)addr(env, hashes), n));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_filterLookupHashes(JNIEnv* env, jclass cls, jlong ctx, jlong f, jobject hashes, jlong n, jobject out) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_lookup_hashes((rb_filter*)(intptr_t)f, (const int64_t*)addr(env, hashes), n, (uint8_t*)addr(env, out)));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_filterLookupThenAddHashes(JNIEnv* env, jclass cls, jlong ctx, jlong f, jobject hashes, jlong n, jobject out) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_lookup_then_add_hashes((rb_filter*)(intptr_t)f, (const int64_t*)addr(env, hashes), n, (uint8_t*)addr(env, out)));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_cbfIncrementHashes(JNIEnv* env, jclass cls, jlong ctx, jlong f, jobject hashes, jlong n) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_cbf_increment_hashes((rb_filter*)(intptr_t)f, (const int64_t*)addr(env, hashes), n));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_cbfCountHashes(JNIEnv* env, jclass cls, jlong ctx, jlong f, jobject hashes, jlong n, jobject out) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_cbf_count_hashes((rb_filter*)(intptr_t)f, (const int64_t*)addr(env, hashes), n, (float*)addr(env, out)));
}
JNIEXPORT jlong JNICALL Java_rnabloom_gpu_Native_filterPopcount(JNIEnv* env, jclass cls, jlong ctx, jlong f) {
    int64_t v = 0;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_popcount((rb_filter*)(intptr_t)f, &v));
    return v;
}
JNIEXPORT jfloat JNICALL Java_rnabloom_gpu_Native_filterFpr(JNIEnv* env, jclass cls, jlong ctx, jlong f) {
    float v = 0;
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_fpr((rb_filter*)(intptr_t)f, &v));
    return v;
}
/* host mirror: fills the Unsafe-backed buffer the unchanged per-k-mer Java code (GraphUtils etc.) keeps reading */
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_filterDownload(JNIEnv* env, jclass cls, jlong ctx, jlong f, jlong dstAddress, jlong nBytes) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_download((rb_filter*)(intptr_t)f, (void*)(intptr_t)dstAddress, nBytes));
}
JNIEXPORT void JNICALL Java_rnabloom_gpu_Native_filterUpload(JNIEnv* env, jclass cls, jlong ctx, jlong f, jlong srcAddress, jlong nBytes) {
    (void)cls;
    check(env, (rb_ctx*)(intptr_t)ctx, rb_filter_upload((rb_filter*)(intptr_t)f, (const void*)(intptr_t)srcAddress, nBytes));
}
