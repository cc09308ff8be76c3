// rb_sliced.cuh -- the sliced engine: graph.add / graph.getKmers without isolated random HBM probes.
//
// Why (profiles/r01_notes.md): an isolated probe costs one 128 B DRAM line request whatever it uses of it, and the part serves
// ~47.5 G of those per second -- 23 % of the 32 B-per-probe roofline.  The same probe against an L2-resident slice of the filter
// costs one L1TEX wavefront (measured ~1.4 cycles per lane, ~200 G/s), and streaming runs at ~6 TB/s.  So every probe is first
// written, as a 4-byte slice-local index, into the region of the filter slice it falls into; then the regions are consumed in
// order by all CTAs together (work handed out by one in-order counter), so that only one or two slices are live in L2 at any
// time; the answer to a probe is one byte at the probe's own position in a parallel array, and a tile picks the answers of
// its k-mers up through the run positions its sort recorded (coalesced copies of whole runs into shared memory, no second sort).
// The logical bit / byte arrays are untouched: this is a schedule, not a blocked Bloom filter.  The same kernels run the
// hash-sharded multi-GPU graph (rb_sshard_host.inl): there the tile sort's regions are (owner rank, slice) and travel by all-to-all.
//
//   lookup (graph.getKmers / getCount, graph/BloomFilterDeBruijnGraph.java:562-570)
//     S1 ks_route_lookup    hash every k-mer, tile-sort its h_d + h_c probes by filter slice, remember the positions
//        (_u: uniform read layout -- the k-mers of a CTA are hashed through XOR-prefix arrays of rotated base seeds)
//     S2 ks_apply_probes<0> slice by slice: read bit / counter, write the answer byte
//     S3 ks_combine_lookup  stage the tile's answer runs in shared memory: count = MiniFloat(min counter) + 1 if all bits are set
//   insert (graph.add, :405-412; addCountIfPresent :424-428; addDbgOnly :430-436)
//     I1 ks_route_keys      tile-sort the base hashes by key range (top bits of a multiplicative hash of the key)
//     I2 ks_split_keys      tile-sort every range again by the next hash bits: sub-ranges of ~1 Ki keys
//     I3 ks_dedup           one CTA per sub-range: (key -> multiplicity) in a shared-memory hash table -> dense distinct keys
//     I4 ks_emit_probes     per distinct key: probes tile-sorted by filter slice, positions remembered
//     I5 ks_apply_probes<1> test-and-set the dbgbf bits (old bit is the answer), read the counters
//     I6 ks_combine_insert  present = AND(old bits); replay m-1+present min-increments on the counter values
//                           (bloom/CountingBloomFilter.java:170-194); the new value of every counter that grew is written over the
//                           probe's answer byte -- the raise travels back to where the probe record already sits, sorted by slice
//     I7 ks_apply_raises    second sweep over the probe regions: counter = max(counter, raise byte), slice by slice
// Linearisation is the one DESIGN.md section 4 states for batches: duplicates of a k-mer inside a round are aggregated, so exactly
// one of them is the first sighting; k-mers that share a counter inside one round see the counter's value at the start of the round.
//
// The tile sort (TileSort below) is a CTA-wide multisplit: shared-memory histogram (the atomic's return value is the record's
// rank inside its bucket), one global cursor bump per (tile, bucket), records staged in shared memory in bucket order and
// copied out so that consecutive threads write consecutive addresses.  Regions have fixed capacities derived from the
// expected load; a round whose hashes are skewed beyond the slack raises the overflow flag *before* any filter is modified
// and is redone by the direct engine.
#pragma once
#include "rb_kernels.cuh"

namespace rb {

constexpr int kSlThreads = 256;        // every kernel here runs 256-thread CTAs (cta_exclusive_scan relies on it)
constexpr int kSlMaxH = 3;             // hashes per filter the engine is built for
constexpr int kSlNJ = 2 * kSlMaxH;     // probe slots per k-mer, separate records: dbgbf hashes at 0..2, cbf hashes at 3..5
constexpr int kSlRoundKmers = 4;       // k-mers per thread and sort round (key sorts; probe sorts: SlShape<NJ>::KPT)
constexpr int kSlTileRecords = 24;     // probe records per thread and tile sort
// Probe records per k-mer.  NJ = 6: one record per probe (any filter sizes).  NJ = 3 ("paired"): hash j of a k-mer indexes both
// filters -- (h_j >>> 1) % dbg_bits and (h_j >>> 1) % cbf_bytes (bloom/BloomFilter.java:108-111, bloom/CountingBloomFilter.java:101-104) --
// so when cbf_bytes is a power of two that divides dbg_bits the counter index is the bit index mod cbf_bytes, and ONE record per
// hash serves both filters if a slice is cut as {counters [s*W, (s+1)*W)} + {bits whose index mod cbf_bytes falls in that range}
// (dbg_bits / cbf_bytes chunks of W bits).  The answer byte carries both: bit 7 = the dbgbf bit, bits 0..6 = the counter (<= 127).
// Half the records to sort, move and answer; the logical arrays are untouched.
template <int NJ>
struct SlShape {
    static constexpr int KPT = kSlTileRecords / NJ;      // k-mers per thread and tile: 4 (NJ = 6) or 8 (NJ = 3)
    static constexpr int TILE = kSlThreads * KPT;        // k-mers per tile: 1024 or 2048
    static constexpr int PFX_PER = NJ == 3 ? 12 : 8;     // bases per thread of the prefix k-merizer's span
};
constexpr int kSlPad = 32;             // one cursor per 128 B line (atomics to one line serialise in L2)
constexpr int kSlMaxRegions = 2048;    // bucket ids are kept in 12 bits, 0xFFF = no record
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;

struct SlArena {
    void* data;               // records; region b = [roff[b], roff[b+1])
    unsigned int* cursor;     // [B * cursor_stride] records appended to region b so far (may pass the capacity: overflow)
    const uint32_t* roff;     // [B + 1] region offsets in records (the whole arena holds < 2^32 records), or nullptr (see cap)
    const uint32_t* rlo;      // [B] or nullptr: explicit region starts (regions of cap records in an order of the consumer's choosing)
    int B;
    int chunk;                // records per work item of the kernels that consume the arena region by region
    uint32_t cap;             // roff == nullptr: every region holds cap records, region b = [b * cap, (b + 1) * cap)
    int cursor_stride;        // cursor of region b = cursor[b * cursor_stride] (kSlPad when the regions are few and hot)
    // TileSort<.., SPILL = true> only: records that do not fit their region go here instead of failing the round (nullptr: fail)
    void* spill_data;
    unsigned int* spill_cursor;
    uint32_t spill_cap;
    // consumer of a sharded round in peer-to-peer mode: region b lives in the arena of source rank b % n_peers, which this GPU reads
    // directly over NVLink (peer_data[src] = that rank's send arena, mapped with CUDA IPC); answers are written straight into the
    // source's answer array peer_ans[src].  nullptr: one local arena (`data`).
    void* const* peer_data;
    uint8_t* const* peer_ans;
    int n_peers;
    // consumer: every region is walked `passes` times (work items of pass 0 first, then pass 1 ...): a paired region may span several
    // L2-resident sub-slices, pass i handles the records of sub-slice i (SlGeom::pair_sub_log2).  1 everywhere else.
    int passes;
    // consumer: the apply kernels bring the next work item into shared memory with cp.async while the current one is applied (host:
    // sl_apply_stage -- on when the records are a peer's, where it hides the NVLink latency; on local records it was measured to cost
    // 0.8 ms per kernel and is off unless RB_SLICED_STAGE=1)
    int stage;
};
template <typename REC>
__device__ __forceinline__ const REC* sl_region_records(const SlArena& a, int region) {
    return reinterpret_cast<const REC*>(a.peer_data ? a.peer_data[region % a.n_peers] : a.data);
}
struct SlGeom {
    FastMod dbg_fm, cbf_fm;   // global index arithmetic (reference semantics)
    int hd, hc;
    int dbg_log2, cbf_log2;   // slice sizes: 2^dbg_log2 bits, 2^cbf_log2 bytes
    int n_dbg, n_cbf;         // probe region = dbgbf slice, or n_dbg + cbf slice
    // paired records (NJ = 3): slice s = counters [s << pair_log2, (s + 1) << pair_log2) and, for every chunk c < dbg_bits / cbf_bytes,
    // the bits c * cbf_bytes + the same range; record = chunk << pair_log2 | offset inside the slice.  n_dbg = n_cbf = 0, n_pair regions.
    int paired, pair_log2, cbf_size_log2, n_pair;
    // cells (paired slices with dbg_bits / cbf_bytes <= 8): the engine works on a co-located copy of both filters, one 16-bit cell per
    // counter index ci: bits 0..7 = counter ci, bit 8 + c = dbgbf bit c * C + ci (C = counters of the share).  A paired record then costs ONE
    // random 32-bit access instead of two (the apply kernels are bound by exactly those: ~218 G divergent L2 loads per second).
    // The kernels receive the cell array through their dbg_words argument; k_cells_pack / k_cells_unpack convert to and from the logical
    // arrays, which stay the layout of everything else (direct kernels, download, save, host mirror).
    int cells;
    // A region of 2^pair_log2 counters is consumed in 2^pair_sub_log2 passes over sub-slices of 2^(pair_log2 - pair_sub_log2) counters:
    // only the sub-slice has to stay L2-resident, so a producer with many owners / slices can sort into fewer, wider regions (the tile
    // sort's cost per tile grows with the number of regions) at the price of streaming the region's records once per pass.
    int pair_sub_log2;
    uint64_t pair_local_c;    // counters (= bits per chunk) of the consumer's share: cbf_bytes on one GPU, shard_p << pair_log2 when sharded
    int shard_p;              // paired slices per rank (sharded graph): region = global slice, owner = slice / shard_p
    // hash-sharded graph (rb_sshard_*, one process per GPU): rank r owns dbgbf slices [r * shard_d, (r+1) * shard_d) and cbf slices
    // [r * shard_c, ...); a producer's probe region = owner * (shard_d + shard_c) + (local dbgbf slice | shard_d + local cbf slice).
    // A consumer sees the regions it received ordered by local region first, source rank second: region / region_div = local
    // region.  Single GPU: shard_d = 0, region_div = 1.
    int shard_d, shard_c, region_div;
};
__device__ __forceinline__ int sl_dbg_region(const SlGeom& sg, uint64_t gi) {
    const int s = (int)(gi >> sg.dbg_log2);
    return sg.shard_d ? (s / sg.shard_d) * (sg.shard_d + sg.shard_c) + s % sg.shard_d : s;
}
__device__ __forceinline__ int sl_cbf_region(const SlGeom& sg, uint64_t gi) {
    const int s = (int)(gi >> sg.cbf_log2);
    return sg.shard_d ? (s / sg.shard_c) * (sg.shard_d + sg.shard_c) + sg.shard_d + s % sg.shard_c : sg.n_dbg + s;
}
// producer view: does region b hold counter probes (its answer bytes carry the raises back)?
__device__ __forceinline__ bool sl_region_has_counters(const SlGeom& sg, int b) {
    if (sg.paired) return true;
    return sg.shard_d ? (b % (sg.shard_d + sg.shard_c)) >= sg.shard_d : b >= sg.n_dbg;
}
__device__ __forceinline__ uint64_t sl_mixkey(uint64_t key) { return key * 0x9E3779B97F4A7C15ULL; }

// Exclusive prefix sum of v[0..n) in shared memory, in place.  Every thread of the (256-thread) CTA calls it; returns the total.
// scratch: 296 words.  No warp shuffles on purpose: the same code runs under the host emulation of tests/emu.
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t* v, int n, uint32_t* scratch) {
    const int t = threadIdx.x;
    const int per = (n + kSlThreads - 1) / kSlThreads;
    const int lo = min(n, t * per), hi = min(n, lo + per);
    uint32_t s = 0;
    for (int i = lo; i < hi; ++i) s += v[i];
    scratch[t] = s;
    __syncthreads();
    if (t < 32) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const uint32_t x = scratch[t * 8 + i]; scratch[t * 8 + i] = acc; acc += x; }
        scratch[kSlThreads + t] = acc;
    }
    __syncthreads();
    if (t == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < 32; ++i) { const uint32_t x = scratch[kSlThreads + i]; scratch[kSlThreads + i] = acc; acc += x; }
        scratch[kSlThreads + 32] = acc;
    }
    __syncthreads();
    uint32_t acc = scratch[t] + scratch[kSlThreads + (t >> 3)];
    for (int i = lo; i < hi; ++i) { const uint32_t x = v[i]; v[i] = acc; acc += x; }
    const uint32_t total = scratch[kSlThreads + 32];
    __syncthreads();
    return total;
}

// ---- CTA-wide multisplit of up to 256 * E records into the regions of an arena -------------------------------------------------
// rank     shared-memory atomicAdd on the bucket's counter; its return value is the record's place inside the (tile, bucket) run.
//          (Warp-level rankings with warp-private counters and no atomics were measured too -- one __ballot_sync per bucket-id
//          bit: 30.4 ms, one __match_any_sync per row: 31.1 ms, against 21-23 ms per 504 M k-mers in ks_route_lookup_u with
//          ATOMS -- the atomic is the cheaper instruction mix on this part; git history has both variants.)
// reserve  one global cursor bump per (tile, bucket), issued before the scan and consumed after the staging, so its ~1 us round trip
//          is overlapped
// stage    records + their bucket ids in bucket order in shared memory
// copy     consecutive threads write consecutive addresses inside a bucket's run
// A run that does not fit its region sets *overflow and is written past the region's end (at most one tile of records: the arenas
// are allocated with that much slack); the host discards the round, so what it overwrites does not matter.
constexpr int kSlWarps = kSlThreads / 32;
constexpr int kSlBucketsPerThread = kSlMaxRegions / kSlThreads;
constexpr int kSlSpill = 8192;   // records of slack behind every arena (>= the largest tile)
__device__ __forceinline__ uint32_t sl_region_lo(const SlArena& a, int region) {
    return a.rlo ? __ldg(&a.rlo[region]) : a.roff ? __ldg(&a.roff[region]) : (uint32_t)region * a.cap;
}
__device__ __forceinline__ uint32_t sl_region_hi(const SlArena& a, int region) {
    return a.rlo ? __ldg(&a.rlo[region]) + a.cap : a.roff ? __ldg(&a.roff[region + 1]) : (uint32_t)(region + 1) * a.cap;
}
template <typename REC, int E, bool SPILL = false>
struct TileSort {
    uint32_t *start, *delta, *scratch;   // [B] [B] [296]
    uint32_t *thresh, *delta2;           // SPILL: [B] staged positions from thresh[b] on go to the spill list, at delta2[b] + position
    REC* stage;                          // [256 * E] records in bucket order
    uint16_t* tag;                       // [256 * E] bucket of each staged record
    int B;
    static __host__ __device__ size_t words_of(int B) { return ((size_t)(SPILL ? 4 : 2) * B + 296 + 3) & ~(size_t)3; }
    static __host__ __device__ size_t smem_bytes(int B) { return words_of(B) * 4 + (size_t)kSlThreads * E * sizeof(REC) + (size_t)kSlThreads * E * 2; }
    __device__ __forceinline__ void init(unsigned char* smem, int B_) {
        B = B_;
        start = reinterpret_cast<uint32_t*>(smem);
        delta = start + B;
        scratch = delta + B;
        thresh = scratch + 296;
        delta2 = thresh + B;
        stage = reinterpret_cast<REC*>(smem + words_of(B) * 4);
        tag = reinterpret_cast<uint16_t*>(smem + words_of(B) * 4 + (size_t)kSlThreads * E * sizeof(REC));
    }
    // In: place[e] = kNoSlot: no record, else the record goes to region region0 + place[e] of `out`.  Out: place[e] = bucket | rank inside
    // the (tile, bucket) run << 12 (kNoSlot: no record).  meta (nullable, B + 1 entries): x = arena position of the bucket's run,
    // y = position of the run inside the tile's bucket-ordered sequence; entry B: y = number of records of the tile.  With them a
    // later kernel finds the tile's answers again: answer of a record = ans[meta[b].x + rank].  Every thread of the CTA calls it.
    __device__ __forceinline__ void run(const SlArena& out, int region0, uint32_t (&place)[E], const REC (&rec)[E], int* overflow,
                                        uint2* __restrict__ meta) {
        const int t = threadIdx.x;
        for (int b = t; b < B; b += kSlThreads) start[b] = 0;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e) if (place[e] != kNoSlot) place[e] |= atomicAdd(&start[place[e]], 1u) << 12;
        __syncthreads();
        uint32_t at[kSlBucketsPerThread];
#pragma unroll
        for (int q = 0; q < kSlBucketsPerThread; ++q) {
            const int b = t + q * kSlThreads;
            at[q] = 0;
            if (b < B) { const uint32_t cnt = start[b]; if (cnt) at[q] = atomicAdd(&out.cursor[(size_t)(region0 + b) * out.cursor_stride], cnt); }
        }
        const uint32_t total = cta_exclusive_scan(start, B, scratch);   // start[b] = staging position of the bucket's first record
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (place[e] != kNoSlot) {
                const uint32_t b = place[e] & 0xFFFu, p = start[b] + (place[e] >> 12);
                stage[p] = rec[e];
                tag[p] = (uint16_t)b;
            }
        }
#pragma unroll
        for (int q = 0; q < kSlBucketsPerThread; ++q) {
            const int b = t + q * kSlThreads;
            if (b < B) {
                const uint32_t cnt = (b + 1 < B ? start[b + 1] : total) - start[b];
                const uint32_t lo = sl_region_lo(out, region0 + b), cap = sl_region_hi(out, region0 + b) - lo;
                const uint32_t a = min(at[q], cap);
                delta[b] = lo + a - start[b];
                if (meta) meta[b] = make_uint2(lo + a, start[b]);
                bool failed = cnt && a + cnt > cap;
                if (SPILL) {
                    thresh[b] = 0xFFFFFFFFu; delta2[b] = 0;
                    if (failed && out.spill_data) {   // the part of the run that does not fit goes to the spill list
                        const uint32_t keep = cap - a, over = cnt - keep;
                        const uint32_t so = atomicAdd(out.spill_cursor, over);
                        if (so + over <= out.spill_cap) { thresh[b] = start[b] + keep; delta2[b] = so - (start[b] + keep); failed = false; }
                    }
                }
                if (failed) atomicOr(reinterpret_cast<unsigned int*>(overflow), 1u);
            }
        }
        if (meta && t == 0) meta[B] = make_uint2(0u, total);
        __syncthreads();
        REC* data = reinterpret_cast<REC*>(out.data);
        if (SPILL) {
            REC* spill = reinterpret_cast<REC*>(out.spill_data);
            for (uint32_t p = t; p < total; p += kSlThreads) {
                const uint32_t b = tag[p];
                if (p < thresh[b]) data[delta[b] + p] = stage[p]; else spill[delta2[b] + p] = stage[p];
            }
        } else {
            for (uint32_t p = t; p < total; p += kSlThreads) data[delta[tag[p]] + p] = stage[p];
        }
        __syncthreads();
    }
};

// the probes of one k-mer (bloom/hash/NTHash.java:518-527 + bloom/BloomFilter.java:108-111).  NJ = 6: slots 0..2 dbgbf, 3..5 cbf;
// NJ = 3: slot j = hash j against both filters (slots >= hc: the counter half of the answer is ignored)
template <int NJ>
__device__ __forceinline__ void sl_probes(const SlGeom& sg, const HashMults& hm, uint64_t base, bool with_cbf, uint32_t* bkt, uint32_t* rec) {
#pragma unroll
    for (int j = 0; j < kSlMaxH; ++j) {
        if (NJ == 3) {
            if (j < sg.hd) {
                const uint64_t gd = fm_index(expand_hash(base, j, hm), sg.dbg_fm);
                const uint64_t gc = gd & sg.cbf_fm.mask;   // == (h_j >>> 1) % cbf_bytes: cbf_bytes is a power of two dividing dbg_bits
                bkt[j] = (uint32_t)(gc >> sg.pair_log2);
                rec[j] = (uint32_t)((gd >> sg.cbf_size_log2) << sg.pair_log2) | (uint32_t)(gc & ((1ULL << sg.pair_log2) - 1));
            }
        } else {
            if (j < sg.hd) {
                const uint64_t gi = fm_index(expand_hash(base, j, hm), sg.dbg_fm);
                bkt[j] = (uint32_t)sl_dbg_region(sg, gi);
                rec[j] = (uint32_t)(gi & ((1ULL << sg.dbg_log2) - 1));
            }
            if (with_cbf && j < sg.hc) {
                const uint64_t gi = fm_index(expand_hash(base, j, hm), sg.cbf_fm);
                bkt[kSlMaxH + j] = (uint32_t)sl_cbf_region(sg, gi);
                rec[kSlMaxH + j] = (uint32_t)(gi & ((1ULL << sg.cbf_log2) - 1));
            }
        }
    }
}
// the NJ places of one k-mer (or distinct key) are 4 * NJ contiguous bytes of the position array
template <int NJ>
__device__ __forceinline__ void sl_store_places(uint32_t* pos, int64_t item, const uint32_t* place) {
    if (NJ % 2 == 0) {
        uint2* dst = reinterpret_cast<uint2*>(pos + item * NJ);
#pragma unroll
        for (int q = 0; q < NJ / 2; ++q) dst[q] = make_uint2(place[2 * q], place[2 * q + 1]);
    } else {
#pragma unroll
        for (int q = 0; q < NJ; ++q) pos[item * NJ + q] = place[q];
    }
}
template <int NJ>
__device__ __forceinline__ void sl_load_places(const uint32_t* pos, int64_t item, uint32_t* place) {
    if (NJ % 2 == 0) {
        const uint2* src = reinterpret_cast<const uint2*>(pos + item * NJ);
#pragma unroll
        for (int q = 0; q < NJ / 2; ++q) { const uint2 v = __ldg(src + q); place[2 * q] = v.x; place[2 * q + 1] = v.y; }
    } else {
#pragma unroll
        for (int q = 0; q < NJ; ++q) place[q] = __ldg(pos + item * NJ + q);
    }
}
// answer byte of a probe: bit 7 = the dbgbf bit (old value when test-and-set), bits 0..6 = the counter
__device__ __forceinline__ bool sl_ans_bit(uint32_t a) { return (a & 0x80u) != 0; }
template <int NJ>
__device__ __forceinline__ int sl_ans_counter(const uint32_t* a, int h) { return (int)((NJ == 3 ? a[h] : a[kSlMaxH + h]) & 0x7Fu); }
// The answers of a tile: its records sit in one run per bucket of the answer array (where the tile sort put them); the runs are
// copied into shared memory in bucket order -- one warp per run, consecutive lanes read consecutive bytes -- and every thread
// then picks its answers up at start[bucket] + rank.  (A per-record gather from global memory costs ~2 L1TEX cycles per
// answer: 20 ms per 504 M k-mers in ks_combine_lookup; the runs of a tile are a few hundred sectors.)
struct TileAnswers {
    uint32_t* start;   // [B + 1] (+ [B + 1] arena positions of the runs while loading)
    uint8_t* bytes;    // [tile records]
    static __host__ __device__ size_t smem_bytes(int B, int tile_records) { return ((size_t)(2 * B + 2) * 4 + (size_t)tile_records + 15) & ~(size_t)15; }
    // every thread of the CTA calls it; meta = the tile's B + 1 entries written by TileSort::run.  Eight lanes copy one run (a run
    // is ~12-24 bytes), 32 runs per CTA step, and nothing in a step depends on the step before: the loads of many runs overlap.
    // RO = false: the kernel writes the runs back later (store): no non-coherent loads of `ans`
    template <bool RO = true>
    __device__ __forceinline__ void load(unsigned char* smem, int B, const uint2* __restrict__ meta, const uint8_t* ans) {
        start = reinterpret_cast<uint32_t*>(smem);
        uint32_t* gpos = start + (B + 1);
        bytes = smem + (size_t)(2 * B + 2) * 4;
        for (int b = threadIdx.x; b <= B; b += kSlThreads) { const uint2 m = __ldg(&meta[b]); start[b] = m.y; gpos[b] = m.x; }
        __syncthreads();
        const int grp = threadIdx.x >> 3, l8 = threadIdx.x & 7;
        for (int b = grp; b < B; b += kSlThreads / 8) {
            const uint32_t lo = start[b], n = start[b + 1] - lo, g = gpos[b];
            for (uint32_t i = l8; i < n; i += 8) bytes[lo + i] = RO ? __ldg(ans + g + i) : __ldcg(ans + g + i);
        }
        __syncthreads();
    }
    __device__ __forceinline__ uint32_t get(uint32_t place) const { return place != kNoSlot ? (uint32_t)bytes[start[place & 0xFFFu] + (place >> 12)] : 0u; }
    // the way back (ks_combine_insert): a thread overwrites the staged bytes of its own records, then the runs go back where they came from
    __device__ __forceinline__ void put(uint32_t place, uint32_t v) { if (place != kNoSlot) bytes[start[place & 0xFFFu] + (place >> 12)] = (uint8_t)v; }
    // every thread of the CTA calls it (after a __syncthreads that follows the last put); ONLY_COUNTERS: runs of dbgbf-only regions stay as they are
    __device__ __forceinline__ void store(int B, uint8_t* __restrict__ ans, const SlGeom& sg) const {
        const uint32_t* gpos = start + (B + 1);
        const int grp = threadIdx.x >> 3, l8 = threadIdx.x & 7;
        for (int b = grp; b < B; b += kSlThreads / 8) {
            if (!sl_region_has_counters(sg, b)) continue;
            const uint32_t lo = start[b], n = start[b + 1] - lo, g = gpos[b];
            for (uint32_t i = l8; i < n; i += 8) ans[g + i] = bytes[lo + i];
        }
    }
};

// ---- prefix k-merizer (uniform read layout) ----------------------------------------------------------------------------------------------
// ntHash is a XOR of rotated per-base seeds (bloom/hash/NTHash.java:332-373), so with x = index of a base in the CTA's span of the
// packed stream,  Gf(x) = rotl(S[b_x], -x),  Gr(x) = rotl(S[3-b_x], x)  and their XOR-prefixes Pf, Pr:
//     forward  hash of the k-mer at x = rotl(Pf(x+k) ^ Pf(x), x+k-1)        reverse hash = rotr(Pr(x+k) ^ Pr(x), x)
// -- two prefix look-ups per strand and k-mer, no k-step seeding and no per-thread rolling state (the rolling recurrence :584-629 is
// this sum evaluated incrementally).  Masked / non-ACGT bases contribute 0 exactly as in the walker (NTHash.java:39-43 "N" seed),
// and a prefix count of them gives the number of unusable bases in any window.  The span (garbage between reads included:
// it cancels in Pf(x+k) ^ Pf(x)) is scanned once per CTA: 8 bases per thread, warp-shuffle XOR-scan of the thread totals.
constexpr int kSlTile = kSlThreads * kSlRoundKmers;   // k-mer positions per CTA of the uniform-layout key kernel
template <int PER, int TILE>   // PER bases per thread of the span, TILE k-mer positions per CTA
struct PrefixKmerizer {
    static constexpr int kSpan = kSlThreads * PER;   // bases a CTA can cover
    unsigned long long *lf, *lr, *of, *orv;   // [kSpan + 8] thread-local exclusive prefixes, [257] thread offsets
    uint16_t* lb; uint32_t* ob;               // masked-base counts
    int64_t abs_lo;
    static __host__ __device__ size_t smem_bytes() { return (size_t)(kSpan + 8) * 16 + 264 * 16 + (size_t)(kSpan + 8) * 2 + 264 * 4 + 3 * kSlWarps * 8; }
    __device__ __forceinline__ int64_t abs_of(const Ingest& g, int64_t p) const {   // first base of launch-local position p
        const int64_t read = p / g.uniform_npos;
        return g.first_base + read * g.uniform_stride + (p - read * g.uniform_npos);
    }
    // every thread of the CTA calls it; tile0 = first position of the CTA
    template <int MODE>
    __device__ __forceinline__ void build(unsigned char* smem, const Ingest& g, int k, int64_t tile0) {
        lf = reinterpret_cast<unsigned long long*>(smem);
        lr = lf + (kSpan + 8);
        of = lr + (kSpan + 8);
        orv = of + 264;
        unsigned long long* wtot = orv + 264;                     // [3 * kSlWarps]
        lb = reinterpret_cast<uint16_t*>(wtot + 3 * kSlWarps);
        ob = reinterpret_cast<uint32_t*>(lb + (kSpan + 8));
        const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
        const int64_t last = min(tile0 + TILE, g.n_pos) - 1;
        abs_lo = abs_of(g, tile0);
        const int span = (int)(abs_of(g, last) + k - abs_lo);   // <= kSpan (the host checks the layout)
        BaseCursor cur;
        cur.seek(g.packed, g.mask, abs_lo + (int64_t)t * PER, g.rcm);
        unsigned long long xf = 0, xr = 0;
        uint32_t xb = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int x = t * PER + i;
            lf[x] = xf; lr[x] = xr; lb[x] = (uint16_t)xb;
            if (x < span) {
                const int c = cur.next();
                if (c & 4) {
                    ++xb;
                    if ((c & 8) && MODE != 0) xr ^= rotl64(seed_of_code(c & 3), x & 63);   // unusable base with a reverse-strand seed (Ingest::rcm)
                } else {
                    if (MODE != 1) xf ^= rotl64(seed_of_code(c & 3), 64 - (x & 63));
                    if (MODE != 0) xr ^= rotl64(seed_of_code(3 - (c & 3)), x & 63);
                }
            }
        }
        // exclusive scan of the thread totals: inclusive warp scans, then the totals of the warps before
        unsigned long long sf = xf, sr = xr;
        uint32_t sb = xb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long yf = __shfl_up_sync(0xffffffffu, sf, o), yr = __shfl_up_sync(0xffffffffu, sr, o);
            const uint32_t yb = __shfl_up_sync(0xffffffffu, sb, o);
            if (lane >= o) { sf ^= yf; sr ^= yr; sb += yb; }
        }
        if (lane == 31) { wtot[warp] = sf; wtot[kSlWarps + warp] = sr; wtot[2 * kSlWarps + warp] = sb; }
        __syncthreads();
        unsigned long long ef = sf ^ xf, er = sr ^ xr;
        uint32_t eb = sb - xb;
        for (int w = 0; w < warp; ++w) { ef ^= wtot[w]; er ^= wtot[kSlWarps + w]; eb += (uint32_t)wtot[2 * kSlWarps + w]; }
        of[t] = ef; orv[t] = er; ob[t] = eb;
        if (t == kSlThreads - 1) {
            of[kSlThreads] = ef ^ xf; orv[kSlThreads] = er ^ xr; ob[kSlThreads] = eb + xb;
            lf[kSpan] = 0; lr[kSpan] = 0; lb[kSpan] = 0;
        }
        __syncthreads();
    }
    // hashes of launch-local position p; bad = number of unusable bases in the window
    template <int MODE>
    __device__ __forceinline__ void eval(const Ingest& g, int k, int64_t p, uint64_t& f, uint64_t& r, int& bad) const {
        const int x = (int)(abs_of(g, p) - abs_lo), y = x + k;
        const int tx = x / PER, ty = y / PER;
        f = 0; r = 0;
        if (MODE != 1) f = rotl64((lf[y] ^ of[ty]) ^ (lf[x] ^ of[tx]), (y - 1) & 63);
        if (MODE != 0) r = rotr64((lr[y] ^ orv[ty]) ^ (lr[x] ^ orv[tx]), x & 63);
        bad = (int)((lb[y] + ob[ty]) - (lb[x] + ob[tx]));
    }
    template <int MODE>
    static __device__ __forceinline__ uint64_t base_of(uint64_t f, uint64_t r) {
        if (MODE == 0) return f;
        if (MODE == 1) return r;
        return ((int64_t)r < (int64_t)f) ? r : f;   // canonical: signed min (NTHash.java:494)
    }
};
using KeyKmerizer = PrefixKmerizer<8, kSlTile>;

// ---- S1 (uniform layout): one CTA = SlShape<NJ>::TILE consecutive k-mer positions, hashed through the prefix arrays, one tile sort ------
template <int MODE, int NJ>
__global__ void __launch_bounds__(kSlThreads, 3) ks_route_lookup_u(const Ingest g, int k, const HashMults hm, const SlGeom sg, const SlArena arena,
                                                               uint32_t* __restrict__ pos, uint2* __restrict__ tile_meta, int64_t* __restrict__ fhash,
                                                               int64_t* __restrict__ rhash, int* overflow) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    constexpr int KPT = SlShape<NJ>::KPT, TILE = SlShape<NJ>::TILE;
    using PK = PrefixKmerizer<SlShape<NJ>::PFX_PER, TILE>;
    const int64_t tile0 = (int64_t)blockIdx.x * TILE;
    PK pk;
    pk.template build<MODE>(sl_smem, g, k, tile0);
    // item i of thread t = position tile0 + i * 256 + t: consecutive lanes read consecutive prefix entries (no bank conflicts) and
    // write consecutive place records
    uint32_t rec[KPT * NJ], slot[KPT * NJ];
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int64_t p = tile0 + i * kSlThreads + threadIdx.x;
#pragma unroll
        for (int j = 0; j < NJ; ++j) { slot[i * NJ + j] = kNoSlot; rec[i * NJ + j] = 0; }
        if (p < g.n_pos) {
            uint64_t f, r; int bad;
            pk.template eval<MODE>(g, k, p, f, r, bad);
            if (fhash) fhash[g.out_base + p] = (int64_t)f;
            if (rhash) rhash[g.out_base + p] = (int64_t)r;
            if (bad == 0) sl_probes<NJ>(sg, hm, PK::template base_of<MODE>(f, r), true, &slot[i * NJ], &rec[i * NJ]);
        }
    }
    __syncthreads();   // the tile sort reuses the shared memory of the prefix arrays
    TileSort<uint32_t, KPT * NJ> ts;
    ts.init(sl_smem, arena.B);
    ts.run(arena, 0, slot, rec, overflow, tile_meta + (size_t)blockIdx.x * (arena.B + 1));
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int64_t p = tile0 + i * kSlThreads + threadIdx.x;
        if (p < g.n_pos) sl_store_places<NJ>(pos, p, &slot[i * NJ]);
    }
}
// ---- I1 (uniform layout) -------------------------------------------------------------------------------------------------------------------
// One CTA = kKeyTile consecutive positions: kKeySub passes of the prefix k-merizer (1024 positions each, 4 keys per thread kept in
// registers), then ONE tile sort of 16 keys per thread -- the per-bucket work of a sort (zeroing, scan, one global cursor bump per
// bucket) is as large as its per-record work when a tile holds about as many keys as there are ranges.
constexpr int kKeySub = 4;
constexpr int kKeyE = kSlRoundKmers * kKeySub;    // keys per thread and tile sort
constexpr int kKeyTile = kSlTile * kKeySub;       // 4096
template <int MODE>
__global__ void __launch_bounds__(kSlThreads, 3) ks_route_keys_u(const Ingest g, int k, int n_ranges, int range_shift, const SlArena arena, int* overflow) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    const int64_t tile0 = (int64_t)blockIdx.x * kKeyTile;
    unsigned long long rec[kKeyE];
    uint32_t slot[kKeyE];
#pragma unroll
    for (int s = 0; s < kKeySub; ++s) {
        const int64_t sub0 = tile0 + (int64_t)s * kSlTile;
#pragma unroll
        for (int i = 0; i < kSlRoundKmers; ++i) { slot[s * kSlRoundKmers + i] = kNoSlot; rec[s * kSlRoundKmers + i] = 0; }
        if (sub0 < g.n_pos) {   // the whole CTA
            KeyKmerizer pk;
            pk.build<MODE>(sl_smem, g, k, sub0);
#pragma unroll
            for (int i = 0; i < kSlRoundKmers; ++i) {
                const int64_t p = sub0 + i * kSlThreads + threadIdx.x;
                if (p < g.n_pos) {
                    uint64_t f, r; int bad;
                    pk.eval<MODE>(g, k, p, f, r, bad);
                    if (bad == 0) {
                        const uint64_t b = KeyKmerizer::base_of<MODE>(f, r);
                        rec[s * kSlRoundKmers + i] = b;
                        slot[s * kSlRoundKmers + i] = n_ranges > 1 ? (uint32_t)(sl_mixkey(b) >> range_shift) : 0;
                    }
                }
            }
            __syncthreads();   // the next pass (or the tile sort) reuses the shared memory of the prefix arrays
        }
    }
    TileSort<unsigned long long, kKeyE, true> ts;
    ts.init(sl_smem, arena.B);
    ts.run(arena, 0, slot, rec, overflow, nullptr);
}

// ---- S1: k-merise, tile-sort the probes of every usable k-mer instance by filter slice --------------------------------------------
template <int MODE, int NJ>
__global__ void __launch_bounds__(kSlThreads, 2) ks_route_lookup(const Ingest g, int k, const HashMults hm, const SlGeom sg, const SlArena arena,
                                                             uint32_t* __restrict__ pos, uint2* __restrict__ tile_meta, int64_t* __restrict__ fhash,
                                                             int64_t* __restrict__ rhash, int* overflow) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    constexpr int KPT = SlShape<NJ>::KPT;
    __shared__ RollLut lut;
    build_lut(&lut, k);
    TileSort<uint32_t, KPT * NJ> ts;
    ts.init(sl_smem, arena.B);
    const int64_t pos0 = ((int64_t)blockIdx.x * kSlThreads + threadIdx.x) * kChunk;
    const int n = pos0 < g.n_pos ? (int)min((int64_t)kChunk, g.n_pos - pos0) : 0;
    PositionWalker<MODE> pw;
    if (n) pw.start(g, pos0, k, lut);
#pragma unroll 1
    for (int r0 = 0; r0 < kChunk; r0 += KPT) {
        uint32_t rec[KPT * NJ], slot[KPT * NJ];
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) { slot[i * NJ + j] = kNoSlot; rec[i * NJ + j] = 0; }
            if (r0 + i < n) {
                pw.advance(g, k, lut);
                const int64_t o = g.out_base + pos0 + r0 + i;
                if (fhash) fhash[o] = (int64_t)pw.wk.f;
                if (rhash) rhash[o] = (int64_t)pw.wk.r;
                if (pw.wk.bad == 0) sl_probes<NJ>(sg, hm, pw.wk.base(), true, &slot[i * NJ], &rec[i * NJ]);
            }
        }
        ts.run(arena, 0, slot, rec, overflow, tile_meta + ((size_t)blockIdx.x * (kChunk / KPT) + r0 / KPT) * (arena.B + 1));
#pragma unroll
        for (int i = 0; i < KPT; ++i) if (r0 + i < n) sl_store_places<NJ>(pos, pos0 + r0 + i, &slot[i * NJ]);
    }
}

// ---- work list of an arena: chunk_prefix[b] = number of arena.chunk-record work items in the regions before b --------------------------------
__global__ void __launch_bounds__(kSlThreads) ks_chunk_prefix(const SlArena arena, int* __restrict__ chunk_prefix) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    uint32_t* v = reinterpret_cast<uint32_t*>(sl_smem);
    uint32_t* scratch = v + ((arena.B + 3) & ~3);
    for (int b = threadIdx.x; b < arena.B; b += kSlThreads) {
        const uint32_t cap = sl_region_hi(arena, b) - sl_region_lo(arena, b);
        const uint32_t cnt = min(arena.cursor[(size_t)b * arena.cursor_stride], cap);
        v[b] = (cnt + (uint32_t)arena.chunk - 1) / (uint32_t)arena.chunk * (uint32_t)max(arena.passes, 1);
    }
    __syncthreads();
    const uint32_t total = cta_exclusive_scan(v, arena.B, scratch);
    for (int b = threadIdx.x; b < arena.B; b += kSlThreads) chunk_prefix[b] = (int)v[b];
    if (threadIdx.x == 0) { chunk_prefix[arena.B] = (int)total; chunk_prefix[arena.B + 1] = 0; }   // [B + 1]: the consumers' work counter
}
// Work items are handed out in region order through one global counter: every CTA takes the lowest unclaimed chunk, so the records
// in flight are always one contiguous window of gridDim.x chunks.  The window must be a fraction of a region (the host sizes grid
// and chunk for that): then one, at region boundaries two, filter slices are live and they stay L2-resident without any grid
// barrier.  Measured: window as large as a region -> 109 B of DRAM reads per probe (no residency at all); static round-robin
// (chunk c to CTA c % grid) -> the CTAs of the test-and-set kernel drift apart by several windows and every filter line is
// fetched ~6 times (106 GB instead of 23 GB per round).
__device__ __forceinline__ int sl_next_chunk(int* counter, int* s_c) {
    __syncthreads();   // everybody is done with the previous chunk and has read *s_c
    if (threadIdx.x == 0) *s_c = atomicAdd(counter, 1);
    __syncthreads();
    return *s_c;
}
struct SlWork { int b; uint32_t first, n; int pass; };
__device__ __forceinline__ void sl_load_prefix(int* pre, const int* chunk_prefix, int B) {
    for (int i = threadIdx.x; i <= B; i += kSlThreads) pre[i] = chunk_prefix[i];
    __syncthreads();
}
__device__ __forceinline__ SlWork sl_work_item(const SlArena& arena, const int* pre, int c) {
    int lo = 0, hi = arena.B;   // pre[lo] <= c < pre[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= c) lo = mid; else hi = mid; }
    SlWork w;
    w.b = lo;
    const uint32_t r_lo = sl_region_lo(arena, lo), cap = sl_region_hi(arena, lo) - r_lo;
    const uint32_t cnt = min(arena.cursor[(size_t)lo * arena.cursor_stride], cap);
    uint32_t local = (uint32_t)(c - pre[lo]);
    w.pass = 0;
    if (arena.passes > 1) {
        const uint32_t per_pass = (cnt + (uint32_t)arena.chunk - 1) / (uint32_t)arena.chunk;
        w.pass = (int)(local / per_pass);
        local -= (uint32_t)w.pass * per_pass;
    }
    const uint32_t off = local * (uint32_t)arena.chunk;
    w.first = r_lo + off;
    w.n = min((uint32_t)arena.chunk, cnt - off);
    return w;
}

// ---- asynchronous staging of a work item's records in shared memory (cp.async: SASS LDGSTS) ----------------------------------------------------
// The apply kernels walk their arena one work item (<= kSlStageRecords records, contiguous) at a time.  The records of the NEXT item are
// copied into shared memory with cp.async while the current item is applied: no registers are held for data in flight, so the latency of
// the copy -- a microsecond out of local HBM, several over NVLink when the arena is a peer's (peer-to-peer mode of the sharded graph) --
// hides behind the filter accesses of the current item.  (Holding the next item's records in registers instead was measured and lost:
// the occupancy it costs is worth more than the latency it hides.)  16-byte pieces: region starts and work-item sizes are multiples of
// 16 records (host: sl_capacity, sl_chunk), so every piece is aligned in global and shared memory.
constexpr int kSlStageRecords = 4096;   // largest work item that is staged (2 buffers of 16 KiB + 2 of 4 KiB); larger ones are read directly
__device__ __forceinline__ void sl_cp16(void* smem_dst, const void* gsrc) {
#ifdef RB_EMU
    memcpy(smem_dst, gsrc, 16);
#else
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void sl_cp_commit() {
#ifndef RB_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING>
__device__ __forceinline__ void sl_cp_wait() {
#ifndef RB_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}
// every thread of the CTA calls it: n_bytes (rounded up to 16) from src to dst, both 16-byte aligned
__device__ __forceinline__ void sl_stage(void* dst, const void* src, uint32_t n_bytes) {
    const uint32_t pieces = (n_bytes + 15u) >> 4;
    for (uint32_t i = threadIdx.x; i < pieces; i += kSlThreads) sl_cp16(reinterpret_cast<char*>(dst) + (size_t)i * 16, reinterpret_cast<const char*>(src) + (size_t)i * 16);
}

// ---- S2 / I5: apply the probes.  SET = 1: dbgbf probes are test-and-set (graph.add / addDbgOnly) -------------------------------------
// STAGED: cp.async staging of the work items (compile-time: a run-time choice between a shared-memory and a global load inside the
// unrolled loops kept the compiler from batching the record loads -- 8.5 -> 11.5 ms look-up)
template <int SET, bool STAGED>
__global__ void __launch_bounds__(kSlThreads) ks_apply_probes(const SlArena arena, int* chunk_prefix, const SlGeom sg,
                                                             uint32_t* __restrict__ dbg_words, const uint32_t* __restrict__ cbf_words,
                                                             uint8_t* __restrict__ ans, const int* abort) {
    if (abort && *abort) return;   // a region overflowed while the round was routed (on this or another rank): nothing may be modified
    RB_DYN_SMEM(unsigned char, sl_smem);
    int* pre = reinterpret_cast<int*>(sl_smem);
    sl_load_prefix(pre, chunk_prefix, arena.B);
    const int total = pre[arena.B];
    constexpr int U = 8;   // probes in flight per thread
    const L2Keep keep = l2_keep_policy();
    __shared__ int s_c;
    // two record buffers behind the prefix table (16-byte aligned)
    constexpr bool staged = STAGED;
    uint32_t* sbuf = reinterpret_cast<uint32_t*>(sl_smem + (((size_t)(arena.B + 1) * 4 + 15) & ~(size_t)15));
    int buf = 0;
    SlWork w_ahead;
    int c = sl_next_chunk(chunk_prefix + arena.B + 1, &s_c);
    if (staged && c < total) {
        w_ahead = sl_work_item(arena, pre, c);
        sl_stage(sbuf, sl_region_records<uint32_t>(arena, w_ahead.b) + w_ahead.first, w_ahead.n * 4u);
        sl_cp_commit();
    }
    // not staged: claim an item, apply it, claim the next (a CTA that claimed ahead would double the window of records in flight and with
    // it the slices that have to stay in L2: measured, 8.5 -> 9.8 ms look-up on local records); staged: the body claims the next item itself
    for (; c < total; c = staged ? c : sl_next_chunk(chunk_prefix + arena.B + 1, &s_c)) {
        SlWork w;
        if (staged) {
            w = w_ahead;
            c = sl_next_chunk(chunk_prefix + arena.B + 1, &s_c);   // its barrier: everybody is done with the buffer the next copy overwrites
            if (c < total) {
                w_ahead = sl_work_item(arena, pre, c);
                sl_stage(sbuf + (buf ^ 1) * arena.chunk, sl_region_records<uint32_t>(arena, w_ahead.b) + w_ahead.first, w_ahead.n * 4u);
                sl_cp_commit();
                sl_cp_wait<1>();   // this item's copy has landed (the next one's may still fly)
            } else sl_cp_wait<0>();
            __syncthreads();
        } else w = sl_work_item(arena, pre, c);
        const uint32_t* cur = sbuf + buf * arena.chunk;
        buf ^= 1;
        const int lr = w.b / sg.region_div;   // local region: dbgbf slices first, then cbf slices -- or paired slices
        const uint32_t* rec = sl_region_records<uint32_t>(arena, w.b);                      // local, or the source rank's arena over NVLink
        uint8_t* ans_out = arena.peer_ans ? arena.peer_ans[w.b % arena.n_peers] : ans;  // answers land where the producer will look
        if (sg.paired) {
            // record = chunk << pair_log2 | offset: counter byte (lr << pair_log2) + offset, bit chunk * pair_local_c + the same
            const uint64_t byte0 = (uint64_t)lr << sg.pair_log2;
            const uint32_t off_mask = (1u << sg.pair_log2) - 1u;
            const int sub_shift = sg.pair_log2 - sg.pair_sub_log2;
            for (uint32_t i0 = threadIdx.x; i0 < w.n; i0 += kSlThreads * U) {
                uint32_t li[U], wd[U], wc[U];
                bool act[U];   // in range and in the sub-slice of this pass
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool in = i0 + u * kSlThreads < w.n;
                    li[u] = !in ? 0u : staged ? cur[i0 + u * kSlThreads] : __ldcs(rec + w.first + i0 + u * kSlThreads);
                    act[u] = in && (int)((li[u] & off_mask) >> sub_shift) == w.pass;
                }
                if (sg.cells) {   // one access per record: the cell word holds the counter and the bit
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        wd[u] = 0;
                        if (act[u]) wd[u] = ld_cg_keep(dbg_words + ((byte0 + (li[u] & off_mask)) >> 1), keep);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (act[u]) {
                            const uint64_t ci = byte0 + (li[u] & off_mask);
                            const int sh = (int)(ci & 1) * 16;
                            const uint32_t bit = 1u << (sh + 8 + (int)(li[u] >> sg.pair_log2));
                            if (SET && !(wd[u] & bit)) wd[u] = atomic_or_keep(dbg_words + (ci >> 1), bit, keep);
                            __stcs(ans_out + w.first + i0 + u * kSlThreads, (uint8_t)(((wd[u] & bit) ? 0x80u : 0u) | ((wd[u] >> sh) & 0x7Fu)));
                        }
                    }
                    continue;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    wd[u] = 0; wc[u] = 0;
                    if (act[u]) {
                        const uint64_t ci = byte0 + (li[u] & off_mask);
                        const uint64_t bi = (uint64_t)(li[u] >> sg.pair_log2) * sg.pair_local_c + ci;
                        wd[u] = ld_cg_keep(dbg_words + (bi >> 5), keep);
                        wc[u] = ld_cg_keep(cbf_words + (ci >> 2), keep);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (act[u]) {
                        const uint64_t ci = byte0 + (li[u] & off_mask);
                        const uint64_t bi = (uint64_t)(li[u] >> sg.pair_log2) * sg.pair_local_c + ci;
                        const uint32_t bit = 1u << (bi & 31);
                        if (SET && !(wd[u] & bit)) wd[u] = atomic_or_keep(dbg_words + (bi >> 5), bit, keep);
                        __stcs(ans_out + w.first + i0 + u * kSlThreads, (uint8_t)(((wd[u] & bit) ? 0x80u : 0u) | ((wc[u] >> ((ci & 3) * 8)) & 0x7Fu)));
                    }
                }
            }
            continue;
        }
        const bool is_dbg = lr < sg.n_dbg;
        const int64_t word0 = is_dbg ? ((int64_t)lr << (sg.dbg_log2 - 5)) : ((int64_t)(lr - sg.n_dbg) << (sg.cbf_log2 - 2));
        for (uint32_t i0 = threadIdx.x; i0 < w.n; i0 += kSlThreads * U) {
            uint32_t li[U], wd[U];
#pragma unroll
            for (int u = 0; u < U; ++u) li[u] = !(i0 + u * kSlThreads < w.n) ? 0u : staged ? cur[i0 + u * kSlThreads] : __ldcs(rec + w.first + i0 + u * kSlThreads);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                wd[u] = 0;
                if (i0 + u * kSlThreads < w.n) wd[u] = is_dbg ? ld_cg_keep(dbg_words + word0 + (li[u] >> 5), keep) : ld_cg_keep(cbf_words + word0 + (li[u] >> 2), keep);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u * kSlThreads < w.n) {
                    uint32_t value;
                    if (is_dbg) {
                        const uint32_t bit = 1u << (li[u] & 31);
                        if (SET && !(wd[u] & bit)) wd[u] = atomic_or_keep(dbg_words + word0 + (li[u] >> 5), bit, keep);
                        value = (wd[u] & bit) ? 0x80u : 0u;
                    } else {
                        value = (wd[u] >> ((li[u] & 3) * 8)) & 0x7Fu;
                    }
                    __stcs(ans_out + w.first + i0 + u * kSlThreads, (uint8_t)value);
                }
            }
        }
    }
}

// ---- S3: gather the answers, write the counts (graph :562-570) ---------------------------------------------------------------------------
// Same CTA / thread -> k-mer mapping as the route kernel that wrote the positions (FLAT = 1: ks_route_lookup_u, 1024 consecutive
// k-mers per CTA; FLAT = 0: ks_route_lookup, 16 consecutive k-mers per thread in 4 rounds): the answers of a CTA round sit in
// the few hundred runs its tile sort wrote, so the 32 B sectors a CTA gathers from are shared by its own threads (L1 hits)
// instead of being fetched again by CTAs on other SMs (measured with mismatched mappings: 43 ms per 504 M k-mers).
template <int NJ>
__device__ __forceinline__ void sl_count_of_kmer(const uint32_t* __restrict__ pos, const TileAnswers& ta, int64_t inst, int hd, int hc,
                                                 float* __restrict__ counts, int64_t out_base) {
    uint32_t place[NJ], a[NJ];
    sl_load_places<NJ>(pos, inst, place);
#pragma unroll
    for (int j = 0; j < NJ; ++j) a[j] = ta.get(place[j]);
    float c = 0.f;
    bool all = place[0] != kNoSlot;   // unusable k-mers (masked base in the window) made no probes
#pragma unroll
    for (int h = 0; h < kSlMaxH; ++h) if (h < hd) all = all && sl_ans_bit(a[h]);
    if (all) {
        int mn = 127;
#pragma unroll
        for (int h = 0; h < kSlMaxH; ++h) if (h < hc) { const int v = sl_ans_counter<NJ>(a, h); mn = v < mn ? v : mn; }
        c = minifloat_to_float(mn) + 1.f;
    }
    counts[out_base + inst] = c;
}
// FLAT = 1: tiles of ks_route_lookup_u (TILE consecutive k-mers per CTA, item i of thread t = k-mer i * 256 + t);
// FLAT = 0: tiles of ks_route_lookup (16 consecutive k-mers per thread, 16 / KPT tile sorts per CTA).  Same grid as the route kernel.
template <int FLAT, int NJ>
__global__ void __launch_bounds__(kSlThreads) ks_combine_lookup(const uint32_t* __restrict__ pos, const uint2* __restrict__ tile_meta, int B,
                                                               const uint8_t* __restrict__ ans, int64_t n_inst, int hd, int hc,
                                                               float* __restrict__ counts, int64_t out_base) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    constexpr int KPT = SlShape<NJ>::KPT, TILE = SlShape<NJ>::TILE;
    TileAnswers ta;
    if (FLAT) {
        ta.load(sl_smem, B, tile_meta + (size_t)blockIdx.x * (B + 1), ans);
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int64_t inst = (int64_t)blockIdx.x * TILE + i * kSlThreads + threadIdx.x;
            if (inst < n_inst) sl_count_of_kmer<NJ>(pos, ta, inst, hd, hc, counts, out_base);
        }
    } else {
        const int64_t pos0 = ((int64_t)blockIdx.x * kSlThreads + threadIdx.x) * kChunk;
#pragma unroll 1
        for (int r0 = 0; r0 < kChunk; r0 += KPT) {
            ta.load(sl_smem, B, tile_meta + ((size_t)blockIdx.x * (kChunk / KPT) + r0 / KPT) * (B + 1), ans);
#pragma unroll
            for (int i = 0; i < KPT; ++i) if (pos0 + r0 + i < n_inst) sl_count_of_kmer<NJ>(pos, ta, pos0 + r0 + i, hd, hc, counts, out_base);
            __syncthreads();   // the next round overwrites the staged answers
        }
    }
}

// ---- I1: k-merise, tile-sort the base hash of every usable k-mer by key range -------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kSlThreads) ks_route_keys(const Ingest g, int k, int n_ranges, int range_shift, const SlArena arena, int* overflow) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    __shared__ RollLut lut;
    build_lut(&lut, k);
    TileSort<unsigned long long, kChunk, true> ts;
    ts.init(sl_smem, arena.B);
    const int64_t pos0 = ((int64_t)blockIdx.x * kSlThreads + threadIdx.x) * kChunk;
    const int n = pos0 < g.n_pos ? (int)min((int64_t)kChunk, g.n_pos - pos0) : 0;
    PositionWalker<MODE> pw;
    if (n) pw.start(g, pos0, k, lut);
    unsigned long long rec[kChunk];
    uint32_t slot[kChunk];
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        slot[i] = kNoSlot; rec[i] = 0;
        if (i < n) {
            pw.advance(g, k, lut);
            if (pw.wk.bad == 0) {
                const uint64_t b = pw.wk.base();
                rec[i] = b;
                slot[i] = n_ranges > 1 ? (uint32_t)(sl_mixkey(b) >> range_shift) : 0;
            }
        }
    }
    ts.run(arena, 0, slot, rec, overflow, nullptr);
}

// ---- I2: second-level split: the keys of every range are tile-sorted again by their next hash bits ------------------------------------------
// After it a sub-range holds ~1 Ki keys: small enough for a shared-memory hash table, so no global table is ever touched
// (the L2-sliced global table this replaces ran at 4-9 G keys/s: one CAS + one add per key against ~24 B of table per key).
__global__ void __launch_bounds__(kSlThreads, 3) ks_split_keys(const SlArena in, int* chunk_prefix, int sub_bits, int sub_shift, int region_div,
                                                           const SlArena out, int* overflow) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    TileSort<unsigned long long, kKeyE, true> ts;
    const int n_sub = 1 << sub_bits;
    ts.init(sl_smem, n_sub);
    int* pre = reinterpret_cast<int*>(sl_smem + TileSort<unsigned long long, kKeyE, true>::smem_bytes(n_sub));
    sl_load_prefix(pre, chunk_prefix, in.B);
    const int total = pre[in.B];
    __shared__ int s_c;
    for (int c = sl_next_chunk(chunk_prefix + in.B + 1, &s_c); c < total; c = sl_next_chunk(chunk_prefix + in.B + 1, &s_c)) {
        const SlWork w = sl_work_item(in, pre, c);   // in.chunk <= 256 * kKeyE keys
        const unsigned long long* rec_in = sl_region_records<unsigned long long>(in, w.b);
        unsigned long long rec[kKeyE];
        uint32_t slot[kKeyE];
#pragma unroll
        for (int i = 0; i < kKeyE; ++i) {
            const uint32_t idx = threadIdx.x + i * kSlThreads;
            slot[i] = kNoSlot; rec[i] = 0;
            if (idx < w.n) {
                rec[i] = __ldcs(rec_in + w.first + idx);
                slot[i] = n_sub > 1 ? (uint32_t)((sl_mixkey(rec[i]) >> sub_shift) & (uint64_t)(n_sub - 1)) : 0;
            }
        }
        ts.run(out, (w.b / region_div) * n_sub, slot, rec, overflow, nullptr);
    }
}

// ---- I3: one CTA per sub-range: (key -> multiplicity) in a shared-memory hash table, distinct keys appended to the dense arrays ----------------
// keys that did not fit their range / sub-range (heavy hitters: one k-mer with thousands of copies in a round): aggregated here first,
// then merged into the multiplicities ks_dedup finds, and what ks_dedup never saw is appended afterwards (ks_spill_append)
struct SpillTable {
    unsigned long long* keys;   // n_slots + 1; 0 = empty; slot n_slots stands for key 0
    unsigned int* counts;
    uint64_t n_slots;           // power of two; nullptr keys = no spill this round
    int shift;
};
__device__ __forceinline__ unsigned int spill_take(const SpillTable& h, unsigned long long key) {
    if (key == 0ULL) return atomicExch(&h.counts[h.n_slots], 0u);
    uint64_t s = sl_mixkey(key) >> h.shift;
    for (;;) {
        const unsigned long long k = h.keys[s];
        if (k == key) return atomicExch(&h.counts[s], 0u);
        if (k == 0ULL) return 0u;
        s = (s + 1) & (h.n_slots - 1);
    }
}
__global__ void __launch_bounds__(kSlThreads) ks_spill_aggregate(const unsigned long long* __restrict__ spill, const unsigned int* __restrict__ n_spill,
                                                                uint32_t cap, const SpillTable h) {
    const uint32_t n = min(*n_spill, cap);
    for (uint32_t i = blockIdx.x * kSlThreads + threadIdx.x; i < n; i += gridDim.x * kSlThreads) {
        const unsigned long long key = spill[i];
        if (key == 0ULL) { atomicAdd(&h.counts[h.n_slots], 1u); continue; }
        uint64_t s = sl_mixkey(key) >> h.shift;
        for (;;) {
            const unsigned long long old = atomicCAS(&h.keys[s], 0ULL, key);
            if (old == 0ULL || old == key) { atomicAdd(&h.counts[s], 1u); break; }
            s = (s + 1) & (h.n_slots - 1);
        }
    }
}
__global__ void __launch_bounds__(kSlThreads) ks_spill_append(const SpillTable h, unsigned long long* __restrict__ dkey, unsigned int* __restrict__ dmult,
                                                             unsigned int* n_distinct, unsigned int dense_cap, int* overflow) {
    for (uint64_t s = (uint64_t)blockIdx.x * kSlThreads + threadIdx.x; s <= h.n_slots; s += (uint64_t)gridDim.x * kSlThreads) {
        const unsigned int c = h.counts[s];
        if (c) {   // ks_dedup did not meet this key in its sub-range: every copy of it was spilled
            const unsigned int d = atomicAdd(n_distinct, 1u);
            if (d < dense_cap) { dkey[d] = s == h.n_slots ? 0ULL : h.keys[s]; dmult[d] = c; }
            else atomicOr(reinterpret_cast<unsigned int*>(overflow), 1u);
        }
    }
}
constexpr int kSlDedupSlots = 4096;   // 32 KiB of keys + 16 KiB of counters; a sub-range holds fewer keys than that (host: cap < slots)
__global__ void __launch_bounds__(kSlThreads) ks_dedup(const SlArena in, int n_regions, int hash_shift, unsigned long long* __restrict__ dkey,
                                                      unsigned int* __restrict__ dmult, unsigned int* n_distinct, unsigned int dense_cap, int* overflow,
                                                      const SpillTable spill) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    unsigned long long* tkeys = reinterpret_cast<unsigned long long*>(sl_smem);
    unsigned int* tcnt = reinterpret_cast<unsigned int*>(tkeys + kSlDedupSlots);
    __shared__ unsigned int n_occ, n_zero, out_base;
    const unsigned long long* rec = reinterpret_cast<const unsigned long long*>(in.data);
    for (int sr = blockIdx.x; sr < n_regions; sr += gridDim.x) {
        const uint32_t lo = sl_region_lo(in, sr);
        const uint32_t n = min(in.cursor[(size_t)sr * in.cursor_stride], sl_region_hi(in, sr) - lo);
        if (n == 0) continue;   // the whole CTA
        uint32_t T = 256;
        while (T < 2 * n && T < (uint32_t)kSlDedupSlots) T <<= 1;
        for (uint32_t i = threadIdx.x; i < T; i += kSlThreads) { tkeys[i] = 0ULL; tcnt[i] = 0u; }
        if (threadIdx.x == 0) { n_occ = 0; n_zero = 0; }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += kSlThreads) {
            const unsigned long long key = __ldcs(rec + lo + i);
            if (key == 0ULL) { atomicAdd(&n_zero, 1u); continue; }   // 0 marks an empty slot
            uint32_t s = (uint32_t)((sl_mixkey(key) << hash_shift) >> 40) & (T - 1);   // hash bits the two splits did not use
            for (;;) {   // the copy that claims the slot is counted by the slot being taken: one atomic per distinct key, two per duplicate
                const unsigned long long old = atomicCAS(&tkeys[s], 0ULL, key);
                if (old == 0ULL) break;
                if (old == key) { atomicAdd(&tcnt[s], 1u); break; }
                s = (s + 1) & (T - 1);
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < T; i += kSlThreads)
            if (tkeys[i]) tcnt[i] |= atomicAdd(&n_occ, 1u) << 16;   // extra copies (< 2^16: a sub-range holds < 4096 keys) | dense rank
        __syncthreads();
        if (threadIdx.x == 0) out_base = atomicAdd(n_distinct, n_occ + (n_zero ? 1u : 0u));
        __syncthreads();
        if (out_base + n_occ + 1u > dense_cap) {   // more distinct keys than the dense arrays hold (sharded graph: key ranges out of balance)
            if (threadIdx.x == 0) atomicOr(reinterpret_cast<unsigned int*>(overflow), 1u);
        } else {
            for (uint32_t i = threadIdx.x; i < T; i += kSlThreads)
                if (tkeys[i]) {
                    const uint32_t d = out_base + (tcnt[i] >> 16);
                    dkey[d] = tkeys[i];
                    dmult[d] = (tcnt[i] & 0xFFFFu) + 1u + (spill.keys ? spill_take(spill, tkeys[i]) : 0u);
                }
            if (threadIdx.x == 0 && n_zero) { dkey[out_base + n_occ] = 0ULL; dmult[out_base + n_occ] = n_zero + (spill.keys ? spill_take(spill, 0ULL) : 0u); }
        }
        __syncthreads();
    }
}

// ---- I4: the probes of every distinct key, tile-sorted by filter slice ---------------------------------------------------------------------------
template <int NJ>
__global__ void __launch_bounds__(kSlThreads) ks_emit_probes(const unsigned long long* __restrict__ dkey, const unsigned int* __restrict__ n_distinct,
                                                            const HashMults hm, const SlGeom sg, int with_cbf, const SlArena arena,
                                                            uint32_t* __restrict__ pos, uint2* __restrict__ tile_meta, int* overflow) {
    constexpr int KPT = SlShape<NJ>::KPT, TILE = SlShape<NJ>::TILE;
    const int64_t nd = (int64_t)*n_distinct;
    if ((int64_t)blockIdx.x * TILE >= nd) return;   // whole CTA
    RB_DYN_SMEM(unsigned char, sl_smem);
    TileSort<uint32_t, KPT * NJ> ts;
    ts.init(sl_smem, arena.B);
    const int64_t d0 = (int64_t)blockIdx.x * TILE + threadIdx.x;   // item i of the thread = distinct key d0 + i * 256
    uint32_t rec[KPT * NJ], slot[KPT * NJ];
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) { slot[i * NJ + j] = kNoSlot; rec[i * NJ + j] = 0; }
        if (d0 + i * kSlThreads < nd) sl_probes<NJ>(sg, hm, (uint64_t)dkey[d0 + i * kSlThreads], with_cbf != 0, &slot[i * NJ], &rec[i * NJ]);
    }
    ts.run(arena, 0, slot, rec, overflow, tile_meta + (size_t)blockIdx.x * (arena.B + 1));
#pragma unroll
    for (int i = 0; i < KPT; ++i) if (d0 + i * kSlThreads < nd) sl_store_places<NJ>(pos, d0 + i * kSlThreads, &slot[i * NJ]);
}

// ---- I6: per distinct key: present?, replay the increments, write the new value of every counter that grew over the probe's answer byte ----
// The raise goes back the way the answer came: the probe record of that counter still sits in the region of its slice (where the tile
// sort put it), so "raise counter c to v" is one byte at the record's position -- no raise records, no second sort, no region that
// could overflow.  0 = no raise (a counter that grew is >= 1).
template <int NJ>
__global__ void __launch_bounds__(kSlThreads) ks_combine_insert(const unsigned long long* __restrict__ dkey, const unsigned int* __restrict__ dmult,
                                                               const unsigned int* __restrict__ n_distinct, const uint32_t* __restrict__ pos,
                                                               const uint2* __restrict__ tile_meta, int probe_B, uint8_t* ans,
                                                               const SlGeom sg, int policy, uint64_t rng_seed, const int* abort) {
    constexpr int KPT = SlShape<NJ>::KPT, TILE = SlShape<NJ>::TILE;
    if (abort && *abort) return;
    const int64_t nd = (int64_t)*n_distinct;
    if ((int64_t)blockIdx.x * TILE >= nd) return;   // whole CTA
    RB_DYN_SMEM(unsigned char, sl_smem);
    const int64_t d0 = (int64_t)blockIdx.x * TILE + threadIdx.x;   // item i of the thread = distinct key d0 + i * 256 (as in ks_emit_probes)
    TileAnswers ta;   // the tile of ks_emit_probes with the same block index
    ta.load<false>(sl_smem, probe_B, tile_meta + (size_t)blockIdx.x * (probe_B + 1), ans);
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        if (d0 + i * kSlThreads < nd) {
            uint32_t place[NJ], a[NJ];
            sl_load_places<NJ>(pos, d0 + i * kSlThreads, place);
#pragma unroll
            for (int j = 0; j < NJ; ++j) a[j] = ta.get(place[j]);
            const uint64_t key = (uint64_t)dkey[d0 + i * kSlThreads];
            const unsigned int m = dmult[d0 + i * kSlThreads];
            bool present = true;
#pragma unroll
            for (int h = 0; h < kSlMaxH; ++h) if (h < sg.hd) present = present && sl_ans_bit(a[h]);
            // graph.add :405-412 -- the first sighting of an absent k-mer only sets bits; addCountIfPresent :424-428 needs presence
            unsigned int n_inc = (policy == POLICY_COUNT_IF_PRESENT) ? (present ? m : 0u) : (m - 1u + (present ? 1u : 0u));
            int v0[kSlMaxH], v[kSlMaxH];
            int mn0 = 127;
#pragma unroll
            for (int h = 0; h < kSlMaxH; ++h) {
                v0[h] = 127;
                if (h < sg.hc) { v0[h] = sl_ans_counter<NJ>(a, h); mn0 = min(mn0, v0[h]); }
                v[h] = v0[h];
            }
            if (policy == POLICY_COUNT_IF_PRESENT && mn0 == 0) n_inc = 0;   // "&& cbf.getCount(hashVals) > 0" (graph :425)
            uint64_t rr = mix64(key ^ rng_seed);
            for (unsigned int it = 0; it < n_inc; ++it) {   // CountingBloomFilter.increment :170-194, n_inc times
                int mn = 127;
#pragma unroll
                for (int h = 0; h < kSlMaxH; ++h) if (h < sg.hc) mn = min(mn, v[h]);
                if (mn >= 127) break;
                rr = mix64(rr + it);
                const int u = minifloat_increment(mn, rr);
                if (u != mn) {
#pragma unroll
                    for (int h = 0; h < kSlMaxH; ++h) if (h < sg.hc && v[h] == mn) v[h] = u;
                }
            }
            // two hashes of one key on the same counter read the same value and replay to the same value: both records carry the same raise
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int h = NJ == 3 ? j : j - kSlMaxH;   // counter hash of slot j (separate records: slots 0..2 are dbgbf probes)
                if (h >= 0) ta.put(place[j], (h < sg.hc && v[h] > v0[h]) ? (uint32_t)v[h] : 0u);
            }
        }
    }
    __syncthreads();
    ta.store(probe_B, ans, sg);
}

// ---- I7: raise the counters slice by slice: the second sweep over the probe regions, this time reading the raise bytes ------------------------
template <bool STAGED>
__global__ void __launch_bounds__(kSlThreads) ks_apply_raises(const SlArena arena, int* chunk_prefix, const SlGeom sg,
                                                             uint32_t* __restrict__ cbf_words, const uint8_t* __restrict__ raise, const int* abort) {
    if (abort && *abort) return;   // a region overflowed while the round was routed (on this or another rank): nothing may be modified
    RB_DYN_SMEM(unsigned char, sl_smem);
    int* pre = reinterpret_cast<int*>(sl_smem);
    sl_load_prefix(pre, chunk_prefix, arena.B);
    const int total = pre[arena.B];
    constexpr int U = 4;
    const L2Keep keep = l2_keep_policy();
    __shared__ int s_c;
    // the same cp.async pipeline as ks_apply_probes, for the records and their raise bytes (local, or a peer's over NVLink)
    constexpr bool staged = STAGED;
    uint32_t* sbuf = reinterpret_cast<uint32_t*>(sl_smem + (((size_t)(arena.B + 1) * 4 + 15) & ~(size_t)15));
    uint8_t* vbuf = reinterpret_cast<uint8_t*>(sbuf + 2 * arena.chunk);
    int buf = 0;
    SlWork w_ahead;
    int c = sl_next_chunk(chunk_prefix + arena.B + 1, &s_c);
    if (staged && c < total) {
        w_ahead = sl_work_item(arena, pre, c);
        sl_stage(sbuf, sl_region_records<uint32_t>(arena, w_ahead.b) + w_ahead.first, w_ahead.n * 4u);
        sl_stage(vbuf, (arena.peer_ans ? arena.peer_ans[w_ahead.b % arena.n_peers] : raise) + w_ahead.first, w_ahead.n);
        sl_cp_commit();
    }
    for (; c < total; c = staged ? c : sl_next_chunk(chunk_prefix + arena.B + 1, &s_c)) {   // as in ks_apply_probes
        SlWork w;
        if (staged) {
            w = w_ahead;
            c = sl_next_chunk(chunk_prefix + arena.B + 1, &s_c);   // its barrier: everybody is done with the buffers the next copy overwrites
            if (c < total) {
                w_ahead = sl_work_item(arena, pre, c);
                sl_stage(sbuf + (buf ^ 1) * arena.chunk, sl_region_records<uint32_t>(arena, w_ahead.b) + w_ahead.first, w_ahead.n * 4u);
                sl_stage(vbuf + (buf ^ 1) * arena.chunk, (arena.peer_ans ? arena.peer_ans[w_ahead.b % arena.n_peers] : raise) + w_ahead.first, w_ahead.n);
                sl_cp_commit();
                sl_cp_wait<1>();
            } else sl_cp_wait<0>();
            __syncthreads();
        } else w = sl_work_item(arena, pre, c);
        const uint32_t* cur_r = sbuf + buf * arena.chunk;
        const uint8_t* cur_v = vbuf + buf * arena.chunk;
        buf ^= 1;
        const int lr = w.b / sg.region_div;   // local region
        if (!sg.paired && lr < sg.n_dbg) continue;   // dbgbf probes (the whole CTA)
        const uint32_t* rec = sl_region_records<uint32_t>(arena, w.b);                              // local, or the source rank's arena over NVLink
        const uint8_t* rb = arena.peer_ans ? arena.peer_ans[w.b % arena.n_peers] : raise;   // ... and the source rank's raise bytes
        const uint64_t byte0 = sg.paired ? (uint64_t)lr << sg.pair_log2 : (uint64_t)(lr - sg.n_dbg) << sg.cbf_log2;
        const uint32_t off_mask = sg.paired ? (1u << sg.pair_log2) - 1u : 0xFFFFFFFFu;
        const int sub_shift = sg.paired ? sg.pair_log2 - sg.pair_sub_log2 : 31;
        // cells: cbf_words is the cell array, two 16-bit cells per word, the counter in the low byte of its cell
        const int wsh = sg.cells ? 1 : 2, bsh = sg.cells ? 16 : 8;
        for (uint32_t i0 = threadIdx.x; i0 < w.n; i0 += kSlThreads * U) {
            uint32_t v[U], li[U], wd[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                v[u] = !(i0 + u * kSlThreads < w.n) ? 0u : staged ? (uint32_t)cur_v[i0 + u * kSlThreads] : (uint32_t)__ldcs(rb + w.first + i0 + u * kSlThreads);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                li[u] = (!v[u] ? 0u : staged ? cur_r[i0 + u * kSlThreads] : __ldcs(rec + w.first + i0 + u * kSlThreads)) & off_mask;
                if (arena.passes > 1 && (int)(li[u] >> sub_shift) != w.pass) v[u] = 0;   // another pass handles this sub-slice
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { wd[u] = 0; if (v[u]) wd[u] = ld_cg_keep(cbf_words + ((byte0 + li[u]) >> wsh), keep); }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (v[u]) byte_raise_keep(cbf_words + ((byte0 + li[u]) >> wsh), (int)((byte0 + li[u]) & (sg.cells ? 1 : 3)) * bsh, v[u], wd[u], keep);
        }
    }
}

// ---- cells <-> logical arrays (SlGeom::cells).  One thread = 32 consecutive counter indices: 8 words of counters, one word of bits per
// chunk, 16 words of cells.  C (counters of the share) is a multiple of 32; q <= 8 chunks.
__global__ void __launch_bounds__(kSlThreads) k_cells_pack(const uint32_t* __restrict__ dbg, const uint32_t* __restrict__ cbf, uint32_t* __restrict__ cells,
                                                          int64_t C, int q) {
  for (int64_t t = (int64_t)blockIdx.x * kSlThreads + threadIdx.x; t < (C >> 5); t += (int64_t)gridDim.x * kSlThreads) {
    uint32_t bits[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bits[c] = c < q ? __ldg(dbg + ((c * C) >> 5) + t) : 0u;
    const uint4* src = reinterpret_cast<const uint4*>(cbf + t * 8);
    uint4* dst = reinterpret_cast<uint4*>(cells + t * 16);
    uint32_t any_bit = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) any_bit |= bits[c];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint4 cw = __ldg(src + h);
        if (!(any_bit | cw.x | cw.y | cw.z | cw.w)) {   // an untouched stretch of the filters
            dst[2 * h] = make_uint4(0u, 0u, 0u, 0u); dst[2 * h + 1] = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        const uint32_t w4[4] = {cw.x, cw.y, cw.z, cw.w};
        uint32_t out[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // counter word j of this half: counters i0 .. i0 + 3
            const int i0 = h * 16 + j * 4;
            uint32_t cell[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint32_t b = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) b |= ((bits[c] >> (i0 + e)) & 1u) << c;
                cell[e] = ((w4[j] >> (8 * e)) & 0xFFu) | (b << 8);
            }
            out[2 * j] = cell[0] | (cell[1] << 16);
            out[2 * j + 1] = cell[2] | (cell[3] << 16);
        }
        dst[2 * h] = make_uint4(out[0], out[1], out[2], out[3]);
        dst[2 * h + 1] = make_uint4(out[4], out[5], out[6], out[7]);
    }
  }
}
__global__ void __launch_bounds__(kSlThreads) k_cells_unpack(const uint32_t* __restrict__ cells, uint32_t* __restrict__ dbg, uint32_t* __restrict__ cbf,
                                                            int64_t C, int q) {
  for (int64_t t = (int64_t)blockIdx.x * kSlThreads + threadIdx.x; t < (C >> 5); t += (int64_t)gridDim.x * kSlThreads) {
    uint32_t bits[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    const uint4* src = reinterpret_cast<const uint4*>(cells + t * 16);
    uint4* dst = reinterpret_cast<uint4*>(cbf + t * 8);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint4 a = __ldg(src + 2 * h), b4 = __ldg(src + 2 * h + 1);
        if (!(a.x | a.y | a.z | a.w | b4.x | b4.y | b4.z | b4.w)) { dst[h] = make_uint4(0u, 0u, 0u, 0u); continue; }
        const uint32_t in[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
        uint32_t cw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i0 = h * 16 + j * 4;
            const uint32_t cell[4] = {in[2 * j] & 0xFFFFu, in[2 * j] >> 16, in[2 * j + 1] & 0xFFFFu, in[2 * j + 1] >> 16};
            cw[j] = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                cw[j] |= (cell[e] & 0xFFu) << (8 * e);
#pragma unroll
                for (int c = 0; c < 8; ++c) bits[c] |= ((cell[e] >> (8 + c)) & 1u) << (i0 + e);
            }
        }
        dst[h] = make_uint4(cw[0], cw[1], cw[2], cw[3]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c < q) dbg[((c * C) >> 5) + t] = bits[c];
  }
}

// ---- exchange helpers of the hash-sharded graph ----------------------------------------------------------------------------------------------
// padded cursors of a producer arena -> dense counts (what travels with the regions)
__global__ void __launch_bounds__(kSlThreads) ks_pack_counts(const unsigned int* __restrict__ cursor, int stride, uint32_t cap, int n, uint32_t* __restrict__ dense) {
    const int i = blockIdx.x * kSlThreads + threadIdx.x;
    if (i < n) dense[i] = min(cursor[(size_t)i * stride], cap);
}
// received counts [source rank][local region] -> consumer order [local region][source rank], and the region starts in that order
__global__ void __launch_bounds__(kSlThreads) ks_order_counts(const uint32_t* __restrict__ recv, int n_ranks, int per_rank, uint32_t cap,
                                                             unsigned int* __restrict__ cursor, uint32_t* __restrict__ rlo) {
    const int i = blockIdx.x * kSlThreads + threadIdx.x;   // consumer region
    if (i < n_ranks * per_rank) {
        const int lr = i / n_ranks, src = i % n_ranks;
        cursor[i] = recv[src * per_rank + lr];
        rlo[i] = (uint32_t)(src * per_rank + lr) * cap;
    }
}

// the same in peer-to-peer mode: the counts are read from the source ranks' packed count arrays, the regions stay where they are
// (source src's arena holds the regions for this rank at (me * per_rank + local region) * cap)
__global__ void __launch_bounds__(kSlThreads) ks_order_counts_p2p(const uint32_t* const* __restrict__ peer_cnt, int n_ranks, int me, int per_rank, uint32_t cap,
                                                                 unsigned int* __restrict__ cursor, uint32_t* __restrict__ rlo) {
    const int i = blockIdx.x * kSlThreads + threadIdx.x;   // consumer region
    if (i < n_ranks * per_rank) {
        const int lr = i / n_ranks, src = i % n_ranks;
        cursor[i] = peer_cnt[src][me * per_rank + lr];
        rlo[i] = (uint32_t)(me * per_rank + lr) * cap;
    }
}

}  // namespace rb
