/*
 * rnabloom_gpu.h -- C-ABI of librnabloom_gpu.so: RNA-Bloom's k-mer / Bloom-filter hot path on B200 (sm_100a).
 *
 * This is the drop-in boundary.  The reference (bcgsc/RNA-Bloom, 100 % Java) has no FFI seam, so each entry
 * point below names the Java method(s) it replaces; a thin JNI shim (jni/rnabloom_jni.c) binds
 * them to rnabloom.gpu.* classes that subclass the reference's BloomFilter / CountingBloomFilter /
 * BloomFilterDeBruijnGraph.  Citations are relative to /root/reference/src/rnabloom/.
 *
 * Conventions
 *   - every call returns int32: RB_OK or a negative RB_E* code; rb_last_error() gives the text.
 *   - all pointers are HOST pointers unless the function name ends in _dev (then: device pointers on the
 *     context's GPU).  Host buffers are caller-owned and may be reused as soon as the call returns.
 *   - calls on one graph/filter from several host threads are serialised internally (the reference shares one
 *     graph between N unsynchronised workers, RNABloom.java:1189-1205); different contexts are independent.
 *   - results are complete (device-synchronised) when a host-pointer call returns; _dev calls are
 *     stream-ordered on the context's stream and need rb_ctx_sync() before the host looks at device memory.
 *
 * Read ingest layout (ours; not the reference's .2bit record format, see DESIGN.md "Ingest"):
 *   packed : 2-bit codes A0 C1 G2 T/U3, base b at bits 2*(b&31).. of little-endian 64-bit word b>>5
 *   mask   : optional 1 bit per base, bit (b&31) of 32-bit word b>>5, set = base unusable (non-ACGTU, or
 *            below the PHRED33 quality floor).  NULL = every base usable.
 *   read i covers bases [read_off[i], read_off[i]+read_len[i]) of those two streams (absolute base indices;
 *   any alignment).  read_off == NULL selects the uniform layout: read i starts at i*uniform_stride and has
 *   uniform_len bases.
 *   A k-mer is inserted iff none of its k bases is masked -- identical to the reference's "maximal runs of
 *   good-quality ACGTU of length >= k" segmentation (RNABloom.java:567-595, util/SeqUtils.java:1430-1438).
 */
#ifndef RNABLOOM_GPU_H
#define RNABLOOM_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB_API __attribute__((visibility("default")))

typedef struct rb_ctx rb_ctx;
typedef struct rb_filter rb_filter;
typedef struct rb_graph rb_graph;

enum {
    RB_OK = 0,
    RB_EINVAL = -1,  /* bad argument */
    RB_ENOMEM = -2,  /* host or device allocation failed */
    RB_ECUDA = -3,   /* CUDA runtime error (text in rb_last_error) */
    RB_ENCCL = -4,   /* multi-GPU exchange: NCCL missing or an NCCL / transport call failed */
    RB_EIO = -5,     /* file I/O */
    RB_ESTATE = -6   /* object not in a state that allows the call (e.g. no rpkbf) */
};

enum { RB_BLOOM = 0, RB_COUNTING = 1 };                      /* rb_filter kinds */
enum { RB_DBGBF = 0, RB_CBF = 1, RB_RPKBF = 2, RB_FPKBF = 3 }; /* filters of a graph */
enum { RB_MODE_FWD = 0, RB_MODE_RC = 1, RB_MODE_CANON = 2 }; /* k-merizer strand mode */

/* flags of rb_graph_add_reads* -- which insert worker body is reproduced */
#define RB_REVCOMP 1u               /* ReverseComplement iterators (stranded graphs only; CanonicalHashFunction.java:188-206 ignores it) */
#define RB_ADD_COUNT_IF_PRESENT 2u  /* graph.addCountIfPresent instead of graph.add  (graph/BloomFilterDeBruijnGraph.java:424-428) */
#define RB_DBG_ONLY 4u              /* graph.addDbgOnly                            (:430-436; FragmentsToGraphWorker RNABloom.java:1489-1516) */
#define RB_STORE_READ_PAIRS 8u      /* also rpkbf.add(pair hash) at distance d_read (RNABloom.java:587-592) */
#define RB_STORE_FRAG_PAIRS 16u     /* also fpkbf.add(pair hash) at distance d_frag (RNABloom.java:1503-1512) */
#define RB_PAIRS_EXISTING_ONLY 32u  /* pair add only if both k-mers are in dbgbf   (RNABloom.java:389-399); implies no k-mer insert */

/* ---- context ------------------------------------------------------------------------------------------ */
RB_API int32_t rb_version(void);
RB_API int32_t rb_ctx_create(int32_t device, rb_ctx** out);
RB_API int32_t rb_ctx_destroy(rb_ctx* ctx);
RB_API const char* rb_last_error(rb_ctx* ctx);          /* ctx may be NULL: error of the last failed create */
RB_API int32_t rb_ctx_set_stream(rb_ctx* ctx, void* cuda_stream); /* run on a caller stream (NULL = back to the own stream) */
RB_API int32_t rb_ctx_sync(rb_ctx* ctx);
RB_API int32_t rb_ctx_set_rng_seed(rb_ctx* ctx, uint64_t seed);  /* MiniFloat coin flips (util/MiniFloat.java:31-38 uses Math.random) */
RB_API int32_t rb_ctx_set_subbatch_kmers(rb_ctx* ctx, int64_t kmers); /* tuning: k-mers per kernel launch */
RB_API int64_t rb_ctx_kernel_launches(rb_ctx* ctx);     /* number of kernels this context has launched so far */
/* device-side stopwatch on the context's stream (CUDA events): start, then stop returns the elapsed milliseconds */
/* Per-kernel device time (CUDA events around every launch of the read-level engines) -- measurement only (bench.py's roofline):
 * enable, run, then read: names receives the kernel names joined by '\n', ms / calls the summed time and launch count of each. */
RB_API int32_t rb_ctx_profile_enable(rb_ctx* ctx, int32_t on);
RB_API int32_t rb_ctx_profile_read(rb_ctx* ctx, char* names, int64_t names_len, float* ms, int32_t* calls, int32_t max_entries,
                                   int32_t* n_out);
RB_API int32_t rb_timer_start(rb_ctx* ctx);
RB_API int32_t rb_timer_stop(rb_ctx* ctx, float* elapsed_ms);
RB_API int32_t rb_host_alloc(void** p, int64_t bytes);  /* pinned host memory for fast transfers */
RB_API int32_t rb_host_free(void* p);
RB_API int32_t rb_dev_alloc(rb_ctx* ctx, void** p, int64_t bytes);
RB_API int32_t rb_dev_free(rb_ctx* ctx, void* p);
RB_API int32_t rb_memcpy_h2d(rb_ctx* ctx, void* dst_dev, const void* src, int64_t bytes);
RB_API int32_t rb_memcpy_d2h(rb_ctx* ctx, void* dst, const void* src_dev, int64_t bytes);

/* ---- host-side helpers (pure host logic, no device traffic) ------------------------------------------------ */
/* BloomFilter.getExpectedSize (bloom/BloomFilter.java:196-199; same in CountingBloomFilter.java:265-268) */
RB_API int64_t rb_expected_size(int64_t n_elements, float fpr, int32_t num_hash);
/* MiniFloat.toFloat (util/MiniFloat.java:40-45) */
RB_API float rb_minifloat_to_float(int8_t b);
/* Pack ASCII reads into the ingest layout on the host.  quals may be NULL (FASTA path).  Read i occupies
 * ascii[ascii_off[i], ascii_off[i+1]); its bases land at base offset out_read_off[i] (reads are padded to a
 * multiple of 32 bases so they start on word boundaries).  Returns the number of bases of stream written. */
RB_API int64_t rb_pack_reads_host(const char* bases, const char* quals, const int64_t* ascii_off, int64_t n_reads,
                                  int32_t min_qual, uint64_t* packed, uint32_t* mask, int64_t* out_read_off,
                                  int32_t* out_read_len);
/* exclusive prefix sum of max(0, len-k+1): where read i's k-mers start in the outputs of rb_graph_count_reads */
RB_API int64_t rb_kmer_offsets(const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int32_t k, int64_t* off);

/* ---- stand-alone filters -------------------------------------------------------------------------------------
 * BloomFilter(size bits, numHash, hashFunction)            bloom/BloomFilter.java:47-59
 * CountingBloomFilter(size bytes, numHash, hashFunction)   bloom/CountingBloomFilter.java:48-61
 * k is what HashFunction carries: it enters the multi-hash expansion NTM64 (bloom/hash/NTHash.java:518-527). */
RB_API int32_t rb_filter_create(rb_ctx* ctx, int32_t kind, int64_t size, int32_t num_hash, int32_t k, rb_filter** out);
RB_API int32_t rb_filter_destroy(rb_filter* f);                       /* destroy()  BloomFilter.java:246-251 */
RB_API int32_t rb_filter_empty(rb_filter* f);                         /* empty()    BloomFilter.java:240-244 */
RB_API int64_t rb_filter_size(const rb_filter* f);                    /* bits (Bloom) or bytes (counting) */
RB_API int64_t rb_filter_num_bytes(const rb_filter* f);               /* bytes of the backing array = file size */
RB_API int32_t rb_filter_num_hash(const rb_filter* f);                /* getNumHash() */
RB_API int32_t rb_filter_device_ptr(rb_filter* f, void** p);          /* HBM address of the byte array */
/* bulk forms of the `long hashVal` overloads: base[i] is hVals[0]; the other numHash-1 values are derived with NTM64
 *   add(long)            BloomFilter.java:139-141     lookup(long)          :180-182
 *   lookupThenAdd(long)  BloomFilter.java:143-145     (out[i] = 1 if all bits were already set; duplicates inside one
 *                                                      call behave as first-then-repeat)
 *   increment(long)      CountingBloomFilter.java:126-128   getCount(long) :231-233   incrementAndGet :196-222 */
RB_API int32_t rb_filter_add_hashes(rb_filter* f, const int64_t* base, int64_t n);
RB_API int32_t rb_filter_lookup_hashes(rb_filter* f, const int64_t* base, int64_t n, uint8_t* out);
RB_API int32_t rb_filter_lookup_then_add_hashes(rb_filter* f, const int64_t* base, int64_t n, uint8_t* out);
RB_API int32_t rb_cbf_increment_hashes(rb_filter* f, const int64_t* base, int64_t n);
RB_API int32_t rb_cbf_increment_and_get_hashes(rb_filter* f, const int64_t* base, int64_t n, float* out);
RB_API int32_t rb_cbf_count_hashes(rb_filter* f, const int64_t* base, int64_t n, float* out);
/* getPopCount / getFPR: set bits (Bloom) or non-zero bytes (counting); fpr = (pop/size)^numHash in double, cast to float
 *   BloomFilter.java:185-194,201-203   CountingBloomFilter.java:254-263   buffer/UnsafeByteBuffer.java:121-150 */
RB_API int32_t rb_filter_popcount(rb_filter* f, int64_t* out);
RB_API int32_t rb_filter_fpr(rb_filter* f, float* out);
/* host mirror: the device array is byte-identical to UnsafeByteBuffer's (bit i = byte i/8, mask 1<<(i%8)) */
RB_API int32_t rb_filter_download(rb_filter* f, void* dst, int64_t nbytes);
RB_API int32_t rb_filter_upload(rb_filter* f, const void* src, int64_t nbytes);
/* save(desc, bits) / file constructors: BloomFilter.java:70-124, CountingBloomFilter.java:66-118 */
RB_API int32_t rb_filter_save(rb_filter* f, const char* desc_path, const char* bits_path);
RB_API int32_t rb_filter_load(rb_ctx* ctx, int32_t kind, const char* desc_path, const char* bits_path, int32_t k,
                              int32_t load_bits, rb_filter** out);

/* getIndex(hashVal, size) = (hashVal >>> 1) % size for an arbitrary positive 63-bit size (bloom/BloomFilter.java:108-111,
 * CountingBloomFilter.java:101-104): the device index arithmetic exposed for parity checks. */
/* ---- CascadingBloomFilter (bloom/CascadingBloomFilter.java:34-100): num_levels Bloom filters of size / num_levels bits each ------------
 * add (:66-72): lookupThenAdd level by level until a level did not have the key yet; lookup (:79-85): the top level;
 * lookupThenAdd (:93-100): 1 iff every level already had the key.  rb_cascade_level lends a level as a filter (popcount, FPR, download). */
typedef struct rb_cascade rb_cascade;
RB_API int32_t rb_cascade_create(rb_ctx* ctx, int64_t size, int32_t num_hash, int32_t k, int32_t num_levels, rb_cascade** out);
RB_API int32_t rb_cascade_destroy(rb_cascade* c);
RB_API int32_t rb_cascade_level(rb_cascade* c, int32_t level, rb_filter** out);
RB_API int32_t rb_cascade_add_hashes(rb_cascade* c, const int64_t* base, int64_t n);
RB_API int32_t rb_cascade_lookup_hashes(rb_cascade* c, const int64_t* base, int64_t n, uint8_t* out);
RB_API int32_t rb_cascade_lookup_then_add_hashes(rb_cascade* c, const int64_t* base, int64_t n, uint8_t* out);

RB_API int32_t rb_index_hashes(rb_ctx* ctx, const int64_t* hash, int64_t n, int64_t size, int64_t* index_out);

/* ---- k-merizer alone (hash parity, and the "hash only" operator) --------------------------------------------
 * NTHashIterator / CanonicalNTHashIterator / ReverseComplementNTHashIterator (bloom/hash/ *NTHashIterator.java): for every k-mer
 * position of every read writes fhash, rhash, base (= hVals[0]) at rb_kmer_offsets()[read] + pos.  Outputs may be NULL.
 * Masked bases hash as seed 0 on both strands (like 'N', NTHash.java:135-168); that is exact for N and for every k-mer that is inserted or
 * counted (> 0), but not for the reverse-strand hash of windows over other non-ACGTU characters: rb_kmerize_ascii /
 * rb_graph_count_reads_ascii carry the c & 0x07 row of NTHash.java:367-373 as well. */
RB_API int32_t rb_kmerize(rb_ctx* ctx, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                          const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride,
                          int32_t k, int32_t mode, int64_t* fhash, int64_t* rhash, int64_t* base);
/* Paired*NTHashIterator (bloom/hash/PairedNTHashIterator.java:55-85 and twins): pair base hash hValsP[0] for positions
 * pos <= len-k-d; out index = rb_kmer_offsets(k+d)[read] + pos. */
RB_API int32_t rb_kmerize_pairs(rb_ctx* ctx, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride,
                                int32_t k, int32_t d, int32_t mode, int64_t* pair_base);

/* ---- BloomFilterDeBruijnGraph ---------------------------------------------------------------------------------
 * ctor  graph/BloomFilterDeBruijnGraph.java:75-104 ; initializePairKmersBloomFilter :352-359 ;
 * set{Read,Frag}PairedKmerDistance :361-389 ; destroy :211-275 */
RB_API int32_t rb_graph_create(rb_ctx* ctx, int64_t dbgbf_bits, int64_t cbf_bytes, int64_t pkbf_bits, int32_t dbgbf_num_hash,
                               int32_t cbf_num_hash, int32_t pkbf_num_hash, int32_t k, int32_t stranded,
                               int32_t use_read_paired_kmers, rb_graph** out);
RB_API int32_t rb_graph_destroy(rb_graph* g);
RB_API int32_t rb_graph_init_fpkbf(rb_graph* g, int64_t pkbf_bits, int32_t pkbf_num_hash);
RB_API int32_t rb_graph_set_distances(rb_graph* g, int32_t d_read, int32_t d_frag);
RB_API int32_t rb_graph_filter(rb_graph* g, int32_t which, rb_filter** out); /* borrowed handle; NULL if absent */
RB_API int32_t rb_graph_clear(rb_graph* g);
/* Execution engine of the read-level insert/lookup calls (different HBM schedule; DESIGN.md sections 3 and 4):
 *   RB_ENGINE_DIRECT  one fused kernel, every probe an isolated random HBM access (one DRAM line request per probe); counter updates are
 *                     linearisable per k-mer instance
 *   RB_ENGINE_SLICED  probes tile-sorted by 64 MiB filter slice, applied slice by slice out of L2, answers picked up per tile;
 *                     rounds of up to 2^29 k-mers.  Needs numHash(dbgbf) <= 3 and numHash(cbf) <= 3; heavy hitters (one k-mer with
 *                     thousands of copies in a round) take a spill path, a round beyond even that is done by the direct engine (before
 *                     anything was modified)
 *   RB_ENGINE_AUTO    (default) sliced for rounds of at least 2^20 k-mers, direct below
 * Results: the bit filters (dbgbf, rpkbf, fpkbf) are byte-identical for either engine, any round size and any GPU count.  The counting filter
 * is byte-identical to the sequential reference wherever no two distinct k-mers share a counter.  Where they do, the sliced engine has
 * SNAPSHOT-PER-ROUND semantics: duplicates of one k-mer inside a round are aggregated exactly (m copies = m-1 or m increments), but two
 * different k-mers of the same round that share a counter both start from the counter's value at the start of the round and the larger
 * result is kept, so such a counter can end one increment below any serial order (measured: ~1e-4 of the counters at 30 % occupancy).
 * The result therefore depends (only on such counters) on the engine and on the round size.
 * The environment variable RB_ENGINE = direct | sliced | auto sets the engine of graphs created afterwards (and of graphs loaded from files). */
enum { RB_ENGINE_DIRECT = 0, RB_ENGINE_SLICED = 2, RB_ENGINE_AUTO = 3 };
RB_API int32_t rb_graph_set_engine(rb_graph* g, int32_t engine);                                 /* clearDbgbf/Cbf/Rpkbf/Fpkbf :211-245 */

/* Bulk insert = the body of the five live insert workers (RNABloom.java:364-732,1463-1539): for every usable k-mer of
 * every read, graph.add / addCountIfPresent / addDbgOnly (graph :405-436), plus pair adds when flagged (:455-461).
 * n_kmers_out (nullable) receives the number of k-mer instances processed. */
RB_API int32_t rb_graph_add_reads(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                  const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride,
                                  uint32_t flags, int64_t* n_kmers_out);
RB_API int32_t rb_graph_add_reads_dev(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                      const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride,
                                      uint32_t flags, int64_t* n_kmers_out);
/* FASTQ/FASTA convenience: ASCII bases (+ PHRED33 qualities or NULL), read i = [ascii_off[i], ascii_off[i+1]).
 * Segmentation and 2-bit packing run on the GPU. */
RB_API int32_t rb_graph_add_reads_ascii(rb_graph* g, const char* bases, const char* quals, const int64_t* ascii_off,
                                        int64_t n_reads, int32_t min_qual, uint32_t flags, int64_t* n_kmers_out);
/* ---- the reference's packed fragment records ("2bit": io/NucleotideBitsWriter.java:24-31, io/NucleotideBitsReader.java:39-49,
 * util/SeqBitsUtils.java:138-262): 4-byte big-endian length + ceil(len/4) bytes, each (b0*64 + b1*16 + b2*4 + b3) - 128 with the first base
 * in the top bits.  rb_graph_add_reads_2bit takes a buffer of such records in host memory (what FragmentsToGraphWorker reads for the stage-3
 * graph rebuild, RNABloom.java:1489-1516) and re-packs them on the GPU; the host helpers are the codec itself. */
RB_API int64_t rb_2bit_record_bytes(int32_t seq_len);
RB_API int64_t rb_2bit_encode_records(const char* bases, const int64_t* ascii_off, int64_t n_reads, uint8_t* out /* NULL: size only */);
RB_API int64_t rb_2bit_index_records(const uint8_t* records, int64_t n_bytes, int64_t* data_off, int32_t* read_len, int64_t max_reads);
RB_API int32_t rb_graph_add_reads_2bit(rb_graph* g, const uint8_t* records, int64_t n_bytes, uint32_t flags, int64_t* n_reads_out, int64_t* n_kmers_out);
/* graph.getKmers(String) over a chunk of sequences (graph :1224-1234, bloom/hash/HashFunction.java:55-85, CanonicalHashFunction.java:46-78):
 * count (0 for windows over a non-ACGTU character), forward and reverse hash of every window, results in host memory.  Unlike the
 * packed entry points (whose unusable bases hash as 0 on both strands, like 'N'), the ASCII entry points reproduce NTHash for every
 * byte value: the reverse strand reads msTab row c & 0x07 (NTHash.java:100-101,367-373), so IUPAC codes such as Y K M S W D and '-'
 * contribute a seed there. */
RB_API int32_t rb_graph_count_reads_ascii(rb_graph* g, const char* bases, const int64_t* ascii_off, int64_t n_reads, float* counts, int64_t* fhash,
                                          int64_t* rhash, int64_t* n_kmers_out);
RB_API int32_t rb_kmerize_ascii(rb_ctx* ctx, const char* bases, const int64_t* ascii_off, int64_t n_reads, int32_t k, int32_t mode, int64_t* fhash,
                                int64_t* rhash, int64_t* base);

/* Bulk lookup = graph.getKmers(seq) (graph :1224-1226 -> HashFunction.java:55-85 / CanonicalHashFunction.java:46-78):
 * per k-mer position count = dbgbf.lookup ? cbf.getCount + 1 : 0 (graph :562-570), 0 when the k-mer covers a masked base;
 * fhash/rhash (nullable) are Kmer.fHashVal / CanonicalKmer.rHashVal.  Output index = rb_kmer_offsets(k)[read] + pos. */
RB_API int32_t rb_graph_count_reads(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                    const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride,
                                    float* counts, int64_t* fhash, int64_t* rhash, int64_t* n_kmers_out);
RB_API int32_t rb_graph_count_reads_dev(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                        const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride,
                                        float* counts, int64_t* fhash, int64_t* rhash, int64_t* n_kmers_out);
/* rb_graph_count_reads without waiting for the results: returns when the kernels of the last round are queued; the device->host copies
 * of counts / hashes run behind whatever the caller does next (the next rb_graph_add_reads, typically: 4 B of count per k-mer take longer
 * over PCIe than the look-up itself).  The read buffers and the (pinned) result buffers must stay untouched until rb_ctx_wait(ticket)
 * returns; tickets complete in order.  rb_graph_sync also waits for everything.  The Java side of it is a pair of direct ByteBuffers
 * used alternately by the worker that consumes the counts. */
RB_API int32_t rb_graph_count_reads_async(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                          const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride,
                                          float* counts, int64_t* fhash, int64_t* rhash, int64_t* n_kmers_out, int64_t* ticket);
RB_API int32_t rb_ctx_wait(rb_ctx* ctx, int64_t ticket);
/* per-hash forms: graph.add(long[]) :405, contains :538, getCount(long) :552 */
RB_API int32_t rb_graph_add_hashes(rb_graph* g, const int64_t* base, int64_t n, uint32_t flags);
RB_API int32_t rb_graph_count_hashes(rb_graph* g, const int64_t* base, int64_t n, float* counts);
/* pair ops: addReadSingleKmerPair / addFragmentSingleKmerPair :455-461, lookup{Read,Fragment}KmerPair :526-532;
 * pair_hash[i] is Kmer.getKmerPairHashValue (graph/Kmer.java:65-67, CanonicalKmer.java:69-72) */
RB_API int32_t rb_graph_add_pair_hashes(rb_graph* g, int32_t which, const int64_t* pair_hash, int64_t n);
RB_API int32_t rb_graph_lookup_pair_hashes(rb_graph* g, int32_t which, const int64_t* pair_hash, int64_t n, uint8_t* out);
/* Barrier, and barrier + refresh of the HOST MIRROR.  The assembler's per-k-mer calls (graph.getCount(long[]), contains, the neighbour
 * iterators: hundreds of latency-bound call sites in util/GraphUtils.java) must not cross the FFI; they keep reading the Unsafe buffers
 * of the inherited filters (bloom/buffer/UnsafeByteBuffer.java:30 `start`, UnsafeBitBuffer.java:31 `backingByteBuffer`), which this call
 * makes byte-identical to the device arrays.  Destinations are host addresses of rb_filter_num_bytes() bytes each; NULL skips a filter.
 * Call before FPR checks on the host, graph.save by the JVM, and stage 2. */
RB_API int32_t rb_graph_sync(rb_graph* g);
RB_API int32_t rb_graph_sync_to_host(rb_graph* g, void* dbgbf, void* cbf, void* rpkbf, void* fpkbf);
/* save / file constructor: graph :297-339,121-189.  Writes <path>, <path>.dbgbf[.desc], <path>.cbf[.desc], <path>.rpkbf[.desc]
 * (if present) and <path>.fpkbf[.desc] (if present) in the reference's format, so the unmodified JAR can restoreGraph() them. */
RB_API int32_t rb_graph_save(rb_graph* g, const char* path);
RB_API int32_t rb_graph_load(rb_ctx* ctx, const char* path, int32_t load_dbgbf, int32_t load_fpkbf, rb_graph** out);

/* ---- f4: a lone Bloom filter over whole sequences: the screening filter of the assembly stages -------------------
 * op RB_SEQ_ADD                for (Kmer kmer : kmers) bf.add(kmer.getHash())              RNABloom.java:1680,2536,2553,2607,2624
 *    RB_SEQ_CONTAINS_ALL       all_found[r] = containsAllKmers(bf, kmers of read r)         util/GraphUtils.java:627-640 (false for a read without k-mers
 *                                                                                           or with a k-mer over an unusable base: "kmer == null")
 *    RB_SEQ_LOOKUP_AND_ADD_ALL all_found[r] = lookupAndAddAllKmers(bf, kmers of read r)     util/GraphUtils.java:642-650, RNABloom.java:4264
 * mode: RB_MODE_FWD / RB_MODE_RC / RB_MODE_CANON = which hash Kmer.getHash() is (graph/Kmer.java, CanonicalKmer.java).  Host pointers, reads
 * as in rb_graph_add_reads; all_found (n_reads bytes, NULL for RB_SEQ_ADD) in host memory.  The final bit array is the same for any order
 * of the reads; the answers of RB_SEQ_LOOKUP_AND_ADD_ALL for two reads of one batch that share a k-mer depend on their order, exactly as
 * they do between the reference's worker threads, which share the screening filter without a lock. */
enum { RB_SEQ_ADD = 0, RB_SEQ_CONTAINS_ALL = 1, RB_SEQ_LOOKUP_AND_ADD_ALL = 2 };
RB_API int32_t rb_filter_seq_op(rb_filter* f, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int32_t mode, int32_t op, uint8_t* all_found);

/* The other two consumers of SURVEY 8f rank 4, as the batched primitives their loops are made of:
 *  - MinimizerHashIterator.next() for every window of w consecutive k-mers of every read (bloom/hash/MinimizerHashIterator.java:42-101,
 *    util/LongRollingWindow.java:43-73: the signed minimum of hVals[0]) -- the key generator of SeqSubsampler.minimizerBased
 *    (util/SeqSubsampler.java:50-117), whose keys then go through rb_cbf_increment_and_get_hashes sequence by sequence.  max(0, len - k - w + 2)
 *    values per read (offsets: rb_kmer_offsets with k + w - 1); w <= 64.
 *  - graph.lookupReadKmerPair / lookupFragmentKmerPair (:526-532) for every pair position of every read: the test inside
 *    breakWithReadPairedKmers / breakWithFragPairedKmers (util/GraphUtils.java:4184-4310).  One byte per pair position (rb_kmer_offsets
 *    with k + d); 0 where the span covers an unusable base. */
RB_API int32_t rb_minimizers(rb_ctx* ctx, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                             int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int32_t k, int32_t w, int32_t mode, int64_t* minimizers);
RB_API int32_t rb_graph_lookup_pairs_reads(rb_graph* g, int32_t which, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                           const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, uint32_t flags,
                                           uint8_t* found);

/* ---- f3: k-mer multiplicity histogram by hash sampling -----------------------------------------------------------
 * The reference sizes its filters from the histogram of the external `ntcard` binary (RNABloom.java:5745-5768, 6939-7010; parsed by
 * util/NTCardHistogram.java:33-63: F1 = k-mers, F0 = distinct k-mers, f_m = distinct k-mers of multiplicity m).  Here: a k-mer is sampled
 * when the top sample_bits bits of a multiplicative mix of its hash are zero; sampled k-mers are counted exactly in a device table, so the
 * sample's histogram is exact and F0 / f_m are that histogram times 2^sample_bits (sample_bits = 0: exact counts).  Accumulates over calls.
 * rb_card_histogram: totals[0] = F1, [1] = sampled instances, [2] = sampled distinct k-mers, [3] = 2^sample_bits; hist[m - 1] = sampled
 * distinct k-mers of multiplicity m for m <= max_mult (<= 65535), hist[max_mult] = those above (max_mult + 1 entries).  stage1.py writes
 * them in ntcard's file format, which the unmodified JAR parses instead of running ntcard when the file exists (RNABloom.java:5750). */
typedef struct rb_card rb_card;
RB_API int32_t rb_card_create(rb_ctx* ctx, int32_t k, int32_t stranded, int32_t sample_bits, int64_t table_slots, rb_card** out);
RB_API int32_t rb_card_destroy(rb_card* c);
RB_API int32_t rb_card_add_reads(rb_card* c, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                 int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int64_t* n_kmers_out);
RB_API int32_t rb_card_add_reads_dev(rb_card* c, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                     int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int64_t* n_kmers_out);
RB_API int32_t rb_card_histogram(rb_card* c, int64_t* totals, int64_t* hist, int32_t max_mult);

/* ---- synthetic workload generator (bench + fixtures; not a reference operator) ----------------------------------
 * Deterministic, counter-based: read r of a virtual genome (seed, genome_len), length L, err_ppm substitutions per
 * 1e6 bases; written in the uniform ingest layout (stride = stride_bases, multiple of 32) into device memory. */
RB_API int32_t rb_synth_reads_dev(rb_ctx* ctx, uint64_t seed, uint64_t genome_len, uint64_t first_read, int64_t n_reads,
                                  int32_t L, uint32_t err_ppm, int64_t stride_bases, uint64_t* packed_dev);

/* ---- neighbour query -------------------------------------------------------------------------------------
 * Kmer.getSuccessors / getPredecessors (graph/Kmer.java:213-253) and CanonicalKmer's (graph/CanonicalKmer.java:232-271), batched: for
 * every k-mer i (forward hash, reverse hash for an unstranded graph, 2-bit codes A0 C1 G2 T3 of its first and last base) the
 * graph.getCount of its four candidate successors then its four candidate predecessors (candidate order A, C, G, T), and their
 * hashes: counts / nbr_fhash / nbr_rhash are [n][2][4]; the caller keeps the candidates with count >= minKmerCov (the hash arrays
 * are nullable; rhash and nbr_rhash are ignored for a stranded graph). */
RB_API int32_t rb_graph_neighbor_counts(rb_graph* g, const int64_t* fhash, const int64_t* rhash, const uint8_t* first_base,
                                        const uint8_t* last_base, int64_t n, float* counts, int64_t* nbr_fhash, int64_t* nbr_rhash);
/* Kmer.getLeftVariants / getRightVariants (graph/Kmer.java:357-405, CanonicalKmer.java:381-519; bloom/hash/LeftVariantsNTHashIterator.java:40-46,
 * RightVariantsNTHashIterator.java:38-44, Canonical*VariantsNTHashIterator.java:42-52) for a batch: counts / hashes [n][2][4] = the k-mers that
 * carry base A, C, G, T in the FIRST position, then in the LAST position (the entry of the k-mer's own base is the k-mer itself). */
RB_API int32_t rb_graph_variant_counts(rb_graph* g, const int64_t* fhash, const int64_t* rhash, const uint8_t* first_base, const uint8_t* last_base,
                                       int64_t n, float* counts, int64_t* var_fhash, int64_t* var_rhash);
/* Kmer.getMaxCovSuccessor / getMaxCovPredecessor (graph/Kmer.java:301-355): best[n][2] = base code of the successor / predecessor with the
 * largest count >= min_cov (first of equal ones in A,C,G,T order), -1 if none; optional count and hashes of it. */
RB_API int32_t rb_graph_max_cov_neighbors(rb_graph* g, const int64_t* fhash, const int64_t* rhash, const uint8_t* first_base, const uint8_t* last_base,
                                          int64_t n, float min_cov, int8_t* best, float* best_count, int64_t* best_fhash, int64_t* best_rhash);
/* GraphUtils.greedyExtendRight / greedyExtendLeft with lookahead <= 1 (util/GraphUtils.java:501-527,1961-1976) for a batch of start k-mers, one
 * GPU thread per walk: kmer_bits = 2 x uint64 per k-mer (2-bit codes, base i at bits 2 * (i & 31) of word i >> 5; k <= 64); up to `bound`
 * k-mers are added, each the max-count neighbour >= min_cov; ext_codes[n][bound] = the added bases in walking order, ext_len[n] their number. */
RB_API int32_t rb_graph_greedy_extend(rb_graph* g, const uint64_t* kmer_bits, const int64_t* fhash, const int64_t* rhash, int64_t n, int32_t right,
                                      int32_t bound, float min_cov, int32_t* ext_len, uint8_t* ext_codes, float* ext_counts);

/* ---- hash-sharded graph (one process per GPU) ------------------------------------------------
 * The filters of BloomFilterDeBruijnGraph (graph/BloomFilterDeBruijnGraph.java:75-104) are split by index range over n_ranks GPUs; the
 * ranks' shares reassemble to exactly the arrays a single GPU or the JVM produces (rb_mgraph_layout).  Reads are data-parallel: every
 * rank feeds its own reads, the library routes every probe to the owner of its index and the answers back.  The exchange belongs to
 * the library: rb_mgraph_create_nccl drives NCCL itself (libnccl.so.2 is bound at run time; all ranks pass the 128-byte id that rank 0
 * obtained from rb_nccl_unique_id and distributed over any host-side channel), rb_mgraph_create takes a transport of two functions
 * instead (MPI, gloo, a test double ...).  A transport works on DEVICE pointers and is ordered on the given stream.
 *
 * A round is collective: every rank calls rb_mgraph_add_round_dev / rb_mgraph_count_round_dev the same number of times (with zero
 * reads if it has none left); one round holds at most max_kmers_per_round k-mers per rank.  Pointers of the round calls are DEVICE
 * pointers.  insert = graph.add (:405-412) and its policies via the RB_* flags; count = graph.getKmers counts (:562-570).
 * Errors: a hash skew that overflows the fixed-capacity regions of a round is detected on the device, agreed between the ranks and
 * reported as RB_ESTATE with NOTHING modified (retry with smaller rounds); once the probes of a round are routed the round always completes
 * (counter raises travel back over the probes' own answer bytes: there is no second routing step that could overflow). */
typedef struct rb_mgraph rb_mgraph;
typedef struct rb_transport {
    void* user;
    /* equal-split all-to-all: piece p (bytes_per_rank bytes) of `send` on rank r arrives as piece r of `recv` on rank p */
    int32_t (*all_to_all)(void* user, const void* send, void* recv, int64_t bytes_per_rank, void* stream);
    /* element-wise maximum of n int32 over all ranks, in place */
    int32_t (*all_reduce_max)(void* user, int32_t* buf, int64_t n, void* stream);
    /* optional (may be NULL): every rank contributes `bytes` bytes, recv holds the contributions in rank order.  With it, ranks on one box
     * exchange CUDA IPC handles at creation and the round kernels then read each other's arenas directly over NVLink (peer-to-peer mode:
     * no staged exchange, the transport is only used for barriers); without it every exchange is a staged all_to_all. */
    int32_t (*all_gather)(void* user, const void* send, void* recv, int64_t bytes, void* stream);
} rb_transport;
RB_API int32_t rb_nccl_unique_id(void* id128, int64_t len);   /* ncclGetUniqueId; len >= 128 */
RB_API int32_t rb_mgraph_create_nccl(rb_ctx* ctx, int32_t n_ranks, int32_t rank, const void* nccl_unique_id, int64_t dbgbf_bits, int64_t cbf_bytes,
                                     int32_t dbgbf_num_hash, int32_t cbf_num_hash, int32_t k, int32_t stranded, int64_t max_kmers_per_round,
                                     rb_mgraph** out);
RB_API int32_t rb_mgraph_create(rb_ctx* ctx, int32_t n_ranks, int32_t rank, const rb_transport* transport, int64_t dbgbf_bits, int64_t cbf_bytes,
                                int32_t dbgbf_num_hash, int32_t cbf_num_hash, int32_t k, int32_t stranded, int64_t max_kmers_per_round,
                                rb_mgraph** out);
RB_API int32_t rb_mgraph_destroy(rb_mgraph* mg);
/* layout[0] paired slices (0/1)  [1] dbgbf bits of a full share  [2] cbf bytes of a full share  [3] local dbgbf bits  [4] local cbf bytes
 * [5] chunks = dbgbf_bits / cbf_bytes (paired)  [6] max k-mers per round and rank  [7] / [8] bytes one insert / lookup round sends per rank.
 * unpaired: rank r holds dbgbf bits [r * layout[1], ...) and cbf bytes [r * layout[2], ...);
 * paired:   rank r holds cbf bytes [r * layout[2], ...) and, for every chunk c, the global bits c * cbf_bytes + r * layout[2] + x
 *           (x < layout[2]) as its local bits c * layout[2] + x. */
RB_API int32_t rb_mgraph_layout(rb_mgraph* mg, int64_t* layout9);
RB_API int32_t rb_mgraph_filter(rb_mgraph* mg, int32_t which, rb_filter** out);   /* this rank's share (borrowed): popcount, download, empty */
RB_API int32_t rb_mgraph_stats(rb_mgraph* mg, int64_t* exchanged_bytes, int64_t* rounds);
RB_API int32_t rb_mgraph_peer_to_peer(rb_mgraph* mg);   /* 1: the round kernels read / write peer memory over NVLink (CUDA IPC); 0: staged all_to_all */
RB_API int32_t rb_mgraph_add_round_dev(rb_mgraph* mg, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                       int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, uint32_t flags, int64_t* n_kmers_out);
RB_API int32_t rb_mgraph_count_round_dev(rb_mgraph* mg, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                         int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, float* counts, int64_t* fhash, int64_t* rhash,
                                         int64_t* n_kmers_out);

/* Long reads for BASELINE configs[4] (ONT-like): read r has rb_synth_long_read_len(seed, r) bases (500..3497, mean ~2 kb) with
 * substitutions, insertions and deletions at the given rates per 1e6 emitted bases; ragged layout: read i of the call starts at base
 * read_off_dev[i] (a multiple of 32) of packed_dev.  Same generator, bit for bit, in the CPU checker. */
RB_API int32_t rb_synth_long_read_len(uint64_t seed, uint64_t read);
RB_API int32_t rb_synth_long_reads_dev(rb_ctx* ctx, uint64_t seed, uint64_t genome_len, uint64_t first_read, int64_t n_reads, uint32_t sub_ppm,
                                       uint32_t ins_ppm, uint32_t del_ppm, const int64_t* read_off_dev, uint64_t* packed_dev);

#ifdef __cplusplus
}
#endif
#endif /* RNABLOOM_GPU_H */
