package rnabloom.gpu;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.ArrayList;

import rnabloom.graph.BloomFilterDeBruijnGraph;

/**
 * Drop-in subclass of rnabloom.graph.BloomFilterDeBruijnGraph whose BULK work runs on the B200:
 *   - the insert workers (RNABloom.java:364-732,1463-1539) call {@link #addReads} with a chunk of records instead of looping over
 *     NTHashIterator + graph.add per k-mer;
 *   - whole-read lookups (graph.getKmers) can go through {@link #countReads}.
 * Per-k-mer calls made by the unchanged assembler (GraphUtils: getCount, getSuccessors, contains ...) keep running on the inherited
 * host filters, which {@link #syncToHost()} refreshes from the device arrays (byte-identical layout), so nothing above this class changes.
 * Sketch for maintainers; not compiled in the build image (no JDK).
 */
public class GpuBloomFilterDeBruijnGraph extends BloomFilterDeBruijnGraph {
    private final long ctx;
    private final long handle;

    public GpuBloomFilterDeBruijnGraph(int device, long dbgbfNumBits, long cbfNumBytes, long pkbfNumBits, int dbgbfNumHash, int cbfNumHash,
                                       int pkbfNumHash, int k, boolean stranded, boolean useReadPairedKmers) {
        super(dbgbfNumBits, cbfNumBytes, pkbfNumBits, dbgbfNumHash, cbfNumHash, pkbfNumHash, k, stranded, useReadPairedKmers);
        ctx = Native.ctxCreate(device);
        handle = Native.graphCreate(ctx, dbgbfNumBits, cbfNumBytes, pkbfNumBits, dbgbfNumHash, cbfNumHash, pkbfNumHash, k, stranded,
                                    useReadPairedKmers);
    }

    /** One chunk of FASTQ/FASTA records: segmentation, k-merisation and graph.add on the GPU. quals may be null (FASTA). */
    public long addReads(ArrayList<String> seqs, ArrayList<String> quals, int minBaseQual, boolean reverseComplement,
                         boolean addCountsOnly, boolean storeReadPairedKmers) {
        int total = 0;
        for (String s : seqs) total += s.length();
        ByteBuffer bases = ByteBuffer.allocateDirect(total + 1);
        ByteBuffer q = quals == null ? null : ByteBuffer.allocateDirect(total + 1);
        ByteBuffer off = ByteBuffer.allocateDirect(8 * (seqs.size() + 1)).order(ByteOrder.nativeOrder());
        long pos = 0;
        for (int i = 0; i < seqs.size(); ++i) {
            off.putLong(pos);
            bases.put(seqs.get(i).getBytes(java.nio.charset.StandardCharsets.ISO_8859_1));
            if (q != null) q.put(quals.get(i).getBytes(java.nio.charset.StandardCharsets.ISO_8859_1));
            pos += seqs.get(i).length();
        }
        off.putLong(pos);
        int flags = (reverseComplement ? Native.REVCOMP : 0) | (addCountsOnly ? Native.ADD_COUNT_IF_PRESENT : 0)
                  | (storeReadPairedKmers ? Native.STORE_READ_PAIRS : 0);
        Native.graphSetDistances(ctx, handle, getReadPairedKmerDistance(), getFragPairedKmerDistance());
        return Native.graphAddReadsAscii(ctx, handle, bases, q, off, seqs.size(), minBaseQual, flags);
    }

    /**
     * Copies dbgbf / cbf / rpkbf / fpkbf from HBM into the inherited Unsafe buffers (call before FPR checks on the host, save by the JVM,
     * and stage 2).  Needs the four one-line getters of INTEGRATION.md edit 4 in the reference:
     *   UnsafeByteBuffer.getAddress()      { return start; }                          (bloom/buffer/UnsafeByteBuffer.java:30)
     *   UnsafeBitBuffer.getAddress()       { return backingByteBuffer.getAddress(); } (bloom/buffer/UnsafeBitBuffer.java:31)
     *   BloomFilter.getBufferAddress()     { return ((UnsafeBitBuffer) bitArray).getAddress(); }   (bloom/BloomFilter.java:41)
     *   CountingBloomFilter.getBufferAddress() { return ((UnsafeByteBuffer) counts).getAddress(); } (bloom/CountingBloomFilter.java:42)
     * The device arrays have exactly the Unsafe layout (bit i = byte i/8, mask 1 << (i % 8); one byte per counter), so this is a copy.
     */
    public void syncToHost() {
        Native.graphSyncToHost(ctx, handle,
                               getDbgbf().getBufferAddress(),
                               getCbf().getBufferAddress(),
                               getRpkbf() == null ? 0L : getRpkbf().getBufferAddress(),
                               getFpkbf() == null ? 0L : getFpkbf().getBufferAddress());
    }

    @Override
    public void save(java.io.File graphFile) throws java.io.IOException {
        Native.graphSave(ctx, handle, graphFile.getPath());   // same files as graph/BloomFilterDeBruijnGraph.java:307-329
    }

    public void destroyGpu() {
        Native.graphDestroy(handle);
        Native.ctxDestroy(ctx);
    }
}
