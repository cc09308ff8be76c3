package rnabloom.gpu;

import java.nio.ByteBuffer;

/**
 * Native entry points of librnabloom_jni.so (jni/rnabloom_jni.c) -> librnabloom_gpu.so (include/rnabloom_gpu.h).
 * Handles are opaque longs; bulk data travels in direct ByteBuffers (native byte order).
 * Not compiled in the build image (no JDK there); see INTEGRATION.md.
 */
public final class Native {
    static { System.loadLibrary("rnabloom_jni"); }
    private Native() {}

    // insert flags (include/rnabloom_gpu.h)
    public static final int REVCOMP = 1, ADD_COUNT_IF_PRESENT = 2, DBG_ONLY = 4, STORE_READ_PAIRS = 8, STORE_FRAG_PAIRS = 16,
                            PAIRS_EXISTING_ONLY = 32;
    public static final int BLOOM = 0, COUNTING = 1;
    public static final int DBGBF = 0, CBF = 1, RPKBF = 2, FPKBF = 3;

    public static native long ctxCreate(int device);
    public static native void ctxDestroy(long ctx);
    public static native long graphCreate(long ctx, long dbgbfNumBits, long cbfNumBytes, long pkbfNumBits, int dbgbfNumHash,
                                          int cbfNumHash, int pkbfNumHash, int k, boolean stranded, boolean useReadPairedKmers);
    public static native void graphDestroy(long graph);
    public static native void graphSetDistances(long ctx, long graph, int readPairedKmersDistance, int fragPairedKmersDistance);
    /** RB_ENGINE_DIRECT = 0, RB_ENGINE_SLICED = 2, RB_ENGINE_AUTO = 3 (default): same results, different HBM schedule. */
    public static native void graphSetEngine(long ctx, long graph, int engine);
    public static native void graphInitFpkbf(long ctx, long graph, long numBits, int numHash);
    public static native long graphAddReadsAscii(long ctx, long graph, ByteBuffer bases, ByteBuffer quals, ByteBuffer offsets,
                                                 long numReads, int minBaseQual, int flags);
    public static native long graphAddReadsPacked(long ctx, long graph, ByteBuffer packed, ByteBuffer mask, ByteBuffer readOff,
                                                  ByteBuffer readLen, long numReads, int flags);
    public static native long graphCountReads(long ctx, long graph, ByteBuffer packed, ByteBuffer mask, ByteBuffer readOff,
                                              ByteBuffer readLen, long numReads, ByteBuffer counts, ByteBuffer fHash, ByteBuffer rHash);
    /** graphCountReads without waiting: returns a ticket; counts / hashes are complete once ctxWait(ticket) has returned. */
    public static native long graphCountReadsAsync(long ctx, long graph, ByteBuffer packed, ByteBuffer mask, ByteBuffer readOff,
                                                   ByteBuffer readLen, long numReads, ByteBuffer counts, ByteBuffer fHashVals, ByteBuffer rHashVals);
    public static native void ctxWait(long ctx, long ticket);
    /** graph.getKmers(String) for a chunk of sequences: counts + forward / reverse hashes of every window, exact for every character. */
    public static native long graphCountReadsAscii(long ctx, long graph, ByteBuffer bases, ByteBuffer offsets, long numReads, ByteBuffer counts,
                                                   ByteBuffer fHash, ByteBuffer rHash);
    public static native void graphAddHashes(long ctx, long graph, ByteBuffer hashes, long n, int flags);
    public static native void graphCountHashes(long ctx, long graph, ByteBuffer hashes, long n, ByteBuffer counts);
    /** Batched Kmer.getSuccessors/getPredecessors: counts[n][2][4] (successors A,C,G,T then predecessors), optional neighbour hashes. */
    public static native void graphNeighborCounts(long ctx, long graph, ByteBuffer fHashVals, ByteBuffer rHashVals, ByteBuffer firstBases,
                                                  ByteBuffer lastBases, long n, ByteBuffer counts, ByteBuffer nbrFHashVals, ByteBuffer nbrRHashVals);
    public static native void graphAddPairHashes(long ctx, long graph, int which, ByteBuffer hashes, long n);
    public static native void graphLookupPairHashes(long ctx, long graph, int which, ByteBuffer hashes, long n, ByteBuffer out);
    public static native long graphFilter(long ctx, long graph, int which);
    /** Barrier + refresh of the host mirror: native addresses of the Unsafe buffers behind the inherited filters (0 = skip). */
    public static native void graphSyncToHost(long ctx, long graph, long dbgbfAddress, long cbfAddress, long rpkbfAddress, long fpkbfAddress);
    public static native void graphSync(long ctx, long graph);
    public static native void graphSave(long ctx, long graph, String path);
    public static native long graphLoad(long ctx, String path, boolean loadDbgbf, boolean loadFpkbf);
    public static native long filterCreate(long ctx, int kind, long size, int numHash, int k);
    public static native void filterDestroy(long filter);
    public static native void filterEmpty(long ctx, long filter);
    public static native void filterAddHashes(long ctx, long filter, ByteBuffer hashes, long n);
    public static native void filterLookupHashes(long ctx, long filter, ByteBuffer hashes, long n, ByteBuffer out);
    public static native void filterLookupThenAddHashes(long ctx, long filter, ByteBuffer hashes, long n, ByteBuffer out);
    public static native void cbfIncrementHashes(long ctx, long filter, ByteBuffer hashes, long n);
    public static native void cbfCountHashes(long ctx, long filter, ByteBuffer hashes, long n, ByteBuffer out);
    public static native long filterPopcount(long ctx, long filter);
    public static native float filterFpr(long ctx, long filter);
    public static native void filterDownload(long ctx, long filter, long dstAddress, long numBytes);
    public static native void filterUpload(long ctx, long filter, long srcAddress, long numBytes);
}
