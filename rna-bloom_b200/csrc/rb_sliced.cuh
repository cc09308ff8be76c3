// rb_sliced.cuh -- the sliced engine: graph.add / graph.getKmers without isolated random HBM probes.
//
// Why (profiles/r01_notes.md): an isolated probe costs one 128 B DRAM line request whatever it uses of it, and the part serves
// ~47.5 G of those per second -- 23 % of the 32 B-per-probe roofline.  The same probe against an L2-resident slice of the filter
// costs one L1TEX wavefront (~280 G/s), and streaming runs at ~6 TB/s.  So every probe is first written, as a 4-byte slice-local
// index, into the region of the filter slice it falls into; then the regions are visited in order by all CTAs together, so
// that only a few slices are live in L2 at any time; the answer to a probe is one byte at the probe's own position in a
// parallel array, and the k-mer picks its answers up through the positions it remembered (a gather, no second sort).
// The logical bit / byte arrays are untouched: this is a schedule, not a blocked Bloom filter.
//
//   lookup (graph.getKmers / getCount, graph/BloomFilterDeBruijnGraph.java:562-570)
//     S1 ks_route_lookup    hash every k-mer, tile-sort its h_d + h_c probes by filter slice, remember the positions
//     S2 ks_apply_probes<0> slice by slice: read bit / counter, write the answer byte
//     S3 ks_combine_lookup  gather the answers: count = MiniFloat(min counter) + 1 if all bits are set
//   insert (graph.add, :405-412; addCountIfPresent :424-428; addDbgOnly :430-436)
//     I1 ks_route_keys      tile-sort the base hashes by key range
//     I2 ks_aggregate       key range by key range: (key -> multiplicity) in an L2-resident slice of a hash table
//     I3 ks_compact_table   occupied slots -> dense (key, multiplicity) arrays
//     I4 ks_emit_probes     per distinct key: probes tile-sorted by filter slice, positions remembered
//     I5 ks_apply_probes<1> test-and-set the dbgbf bits (old bit is the answer), read the counters
//     I6 ks_combine_insert  present = AND(old bits); replay m-1+present min-increments on the counter values
//                           (bloom/CountingBloomFilter.java:170-194); raises tile-sorted by counter slice
//     I7 ks_apply_raises    counter = max(counter, value), slice by slice
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
constexpr int kSlNJ = 2 * kSlMaxH;     // probe slots per k-mer: dbgbf hashes at 0..2, cbf hashes at 3..5
constexpr int kSlRoundKmers = 4;       // k-mers per thread and sort round
constexpr int kSlPad = 32;             // one cursor per 128 B line (atomics to one line serialise in L2)
constexpr int kSlMaxRegions = 2048;    // bucket ids are kept in 12 bits, 0xFFF = no record
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;

struct SlArena {
    void* data;               // records; region b = [roff[b], roff[b+1])
    unsigned int* cursor;     // [B * kSlPad] records appended to region b so far (may pass the capacity: overflow)
    const uint32_t* roff;     // [B + 1] region offsets in records (the whole arena holds < 2^32 records)
    int B;
    int chunk;                // records per work item of the kernels that consume the arena region by region
};
struct SlGeom {
    FastMod dbg_fm, cbf_fm;   // global index arithmetic (reference semantics)
    int hd, hc;
    int dbg_log2, cbf_log2;   // slice sizes: 2^dbg_log2 bits, 2^cbf_log2 bytes
    int n_dbg, n_cbf;         // probe region = dbgbf slice, or n_dbg + cbf slice
    int raise_log2, n_raise;  // counter raises: record = slice-local byte index | value << raise_log2 (raise_log2 <= 25)
};
struct SlTable {
    unsigned long long* keys;   // T + 1 slots, 0 = empty; slot T stands for key 0
    unsigned int* counts;
    uint64_t n_slots;           // T (power of two)
    int shift;                  // slot = mixkey >> shift
};
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
template <typename REC, int E>
struct TileSort {
    uint32_t *start, *gdst, *glim, *scratch;   // [B] [B] [B] [296]
    REC* stage;                                // [256 * E] records in bucket order
    uint16_t* tag;                             // [256 * E] bucket of each staged record
    int B;
    static size_t smem_bytes(int B) {
        const size_t words = ((size_t)3 * B + 296 + 3) & ~(size_t)3;
        return words * 4 + (size_t)kSlThreads * E * sizeof(REC) + (size_t)kSlThreads * E * 2;
    }
    __device__ __forceinline__ void init(unsigned char* smem, int B_) {
        B = B_;
        start = reinterpret_cast<uint32_t*>(smem);
        gdst = start + B;
        glim = gdst + B;
        scratch = glim + B;
        const size_t words = ((size_t)3 * B + 296 + 3) & ~(size_t)3;
        stage = reinterpret_cast<REC*>(smem + words * 4);
        tag = reinterpret_cast<uint16_t*>(smem + words * 4 + (size_t)kSlThreads * E * sizeof(REC));
    }
    // bkt[e] < 0: no record.  slot[e] receives the arena position the record was written to (kNoSlot: none, or dropped because its
    // region is full -- *overflow is set then).  Every thread of the CTA must call it.
    __device__ __forceinline__ void run(const SlArena& out, const int (&bkt)[E], const REC (&rec)[E], uint32_t (&slot)[E], int* overflow) {
        const int t = threadIdx.x;
        for (int b = t; b < B; b += kSlThreads) start[b] = 0;
        __syncthreads();
        uint32_t br[E];   // bucket | rank inside (tile, bucket) << 12
#pragma unroll
        for (int e = 0; e < E; ++e) br[e] = bkt[e] >= 0 ? ((uint32_t)bkt[e] | (atomicAdd(&start[bkt[e]], 1u) << 12)) : 0xFFFu;
        __syncthreads();
        for (int b = t; b < B; b += kSlThreads) {   // one cursor bump per (tile, bucket)
            const uint32_t c = start[b];
            uint32_t lo = 0, hi = 0, at = 0;
            if (c) {
                lo = __ldg(&out.roff[b]);
                hi = __ldg(&out.roff[b + 1]);
                at = atomicAdd(&out.cursor[(size_t)b * kSlPad], c);
                if (at > hi - lo) at = hi - lo;
                if (at + c > hi - lo) *overflow = 1;
            }
            gdst[b] = lo + at;
            glim[b] = hi;
        }
        const uint32_t total = cta_exclusive_scan(start, B, scratch);   // start[b] = staging position of the bucket's first record
#pragma unroll
        for (int e = 0; e < E; ++e) {
            slot[e] = kNoSlot;
            if ((br[e] & 0xFFFu) != 0xFFFu) {
                const uint32_t b = br[e] & 0xFFFu, r = br[e] >> 12;
                const uint32_t p = start[b] + r, d = gdst[b] + r;
                stage[p] = rec[e];
                tag[p] = (uint16_t)b;
                if (d < glim[b]) slot[e] = d;
            }
        }
        __syncthreads();
        REC* data = reinterpret_cast<REC*>(out.data);
        for (uint32_t p = t; p < total; p += kSlThreads) {   // consecutive threads -> consecutive addresses inside a bucket's run
            const uint32_t b = tag[p];
            const uint32_t d = gdst[b] + (p - start[b]);
            if (d < glim[b]) data[d] = stage[p];
        }
        __syncthreads();
    }
};

// the probes of one k-mer: slots 0..2 dbgbf, 3..5 cbf (bloom/hash/NTHash.java:518-527 + bloom/BloomFilter.java:108-111)
__device__ __forceinline__ void sl_probes(const SlGeom& sg, const HashMults& hm, uint64_t base, bool with_cbf, int* bkt, uint32_t* rec) {
#pragma unroll
    for (int j = 0; j < kSlMaxH; ++j) {
        if (j < sg.hd) {
            const uint64_t gi = fm_index(expand_hash(base, j, hm), sg.dbg_fm);
            bkt[j] = (int)(gi >> sg.dbg_log2);
            rec[j] = (uint32_t)(gi & ((1ULL << sg.dbg_log2) - 1));
        }
        if (with_cbf && j < sg.hc) {
            const uint64_t gi = fm_index(expand_hash(base, j, hm), sg.cbf_fm);
            bkt[kSlMaxH + j] = sg.n_dbg + (int)(gi >> sg.cbf_log2);
            rec[kSlMaxH + j] = (uint32_t)(gi & ((1ULL << sg.cbf_log2) - 1));
        }
    }
}
// a thread's 4 k-mers x 6 positions are 96 contiguous, 16-byte aligned bytes of the position array
__device__ __forceinline__ void sl_store_positions(uint32_t* pos, int64_t first, const uint32_t (&slot)[kSlRoundKmers * kSlNJ]) {
    uint4* dst = reinterpret_cast<uint4*>(pos + first * kSlNJ);
#pragma unroll
    for (int q = 0; q < kSlNJ; ++q) dst[q] = make_uint4(slot[4 * q], slot[4 * q + 1], slot[4 * q + 2], slot[4 * q + 3]);
}
__device__ __forceinline__ void sl_load_answers(const uint32_t* pos, int64_t first, const uint8_t* __restrict__ ans, uint32_t (&slot)[kSlRoundKmers * kSlNJ],
                                                uint32_t (&a)[kSlRoundKmers * kSlNJ]) {
    const uint4* src = reinterpret_cast<const uint4*>(pos + first * kSlNJ);
#pragma unroll
    for (int q = 0; q < kSlNJ; ++q) {
        const uint4 v = __ldg(src + q);
        slot[4 * q] = v.x; slot[4 * q + 1] = v.y; slot[4 * q + 2] = v.z; slot[4 * q + 3] = v.w;
    }
#pragma unroll
    for (int e = 0; e < kSlRoundKmers * kSlNJ; ++e) a[e] = slot[e] != kNoSlot ? (uint32_t)__ldg(ans + slot[e]) : 0u;
}

// ---- S1: k-merise, tile-sort the probes of every usable k-mer instance by filter slice --------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kSlThreads) ks_route_lookup(const Ingest g, int k, const HashMults hm, const SlGeom sg, const SlArena arena,
                                                             uint32_t* __restrict__ pos, int64_t* __restrict__ fhash, int64_t* __restrict__ rhash,
                                                             int* overflow) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    __shared__ RollLut lut;
    build_lut(&lut, k);
    TileSort<uint32_t, kSlRoundKmers * kSlNJ> ts;
    ts.init(sl_smem, arena.B);
    const int64_t pos0 = ((int64_t)blockIdx.x * kSlThreads + threadIdx.x) * kChunk;
    const int n = pos0 < g.n_pos ? (int)min((int64_t)kChunk, g.n_pos - pos0) : 0;
    PositionWalker<MODE> pw;
    if (n) pw.start(g, pos0, k, lut);
#pragma unroll 1
    for (int r0 = 0; r0 < kChunk; r0 += kSlRoundKmers) {
        int bkt[kSlRoundKmers * kSlNJ];
        uint32_t rec[kSlRoundKmers * kSlNJ], slot[kSlRoundKmers * kSlNJ];
#pragma unroll
        for (int i = 0; i < kSlRoundKmers; ++i) {
#pragma unroll
            for (int j = 0; j < kSlNJ; ++j) { bkt[i * kSlNJ + j] = -1; rec[i * kSlNJ + j] = 0; }
            if (r0 + i < n) {
                pw.advance(g, k, lut);
                const int64_t o = g.out_base + pos0 + r0 + i;
                if (fhash) fhash[o] = (int64_t)pw.wk.f;
                if (rhash) rhash[o] = (int64_t)pw.wk.r;
                if (pw.wk.bad == 0) sl_probes(sg, hm, pw.wk.base(), true, &bkt[i * kSlNJ], &rec[i * kSlNJ]);
            }
        }
        ts.run(arena, bkt, rec, slot, overflow);
        if (r0 < n) sl_store_positions(pos, pos0 + r0, slot);
    }
}

// ---- work list of an arena: chunk_prefix[b] = number of arena.chunk-record work items in the regions before b --------------------------------
__global__ void __launch_bounds__(kSlThreads) ks_chunk_prefix(const SlArena arena, int* __restrict__ chunk_prefix) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    uint32_t* v = reinterpret_cast<uint32_t*>(sl_smem);
    uint32_t* scratch = v + ((arena.B + 3) & ~3);
    for (int b = threadIdx.x; b < arena.B; b += kSlThreads) {
        const uint32_t cap = arena.roff[b + 1] - arena.roff[b];
        const uint32_t cnt = min(arena.cursor[(size_t)b * kSlPad], cap);
        v[b] = (cnt + (uint32_t)arena.chunk - 1) / (uint32_t)arena.chunk;
    }
    __syncthreads();
    const uint32_t total = cta_exclusive_scan(v, arena.B, scratch);
    for (int b = threadIdx.x; b < arena.B; b += kSlThreads) chunk_prefix[b] = (int)v[b];
    if (threadIdx.x == 0) chunk_prefix[arena.B] = (int)total;
}
// Work items are dealt round-robin in region order, so at any moment the whole grid works inside a window of gridDim.x * chunk
// records.  The window must be a small fraction of a region (the host sizes grid and chunk for that): then one, at region
// boundaries two, filter slices are live and they stay L2-resident without any grid barrier.  (Measured with a window as large
// as a region: 109 B of DRAM reads per probe, i.e. no residency at all -- profiles/r01_notes.md section 6.)
struct SlWork { int b; uint32_t first, n; };
__device__ __forceinline__ void sl_load_prefix(int* pre, const int* __restrict__ chunk_prefix, int B) {
    for (int i = threadIdx.x; i <= B; i += kSlThreads) pre[i] = chunk_prefix[i];
    __syncthreads();
}
__device__ __forceinline__ SlWork sl_work_item(const SlArena& arena, const int* pre, int c) {
    int lo = 0, hi = arena.B;   // pre[lo] <= c < pre[hi]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= c) lo = mid; else hi = mid; }
    SlWork w;
    w.b = lo;
    const uint32_t r_lo = __ldg(&arena.roff[lo]), cap = __ldg(&arena.roff[lo + 1]) - r_lo;
    const uint32_t cnt = min(arena.cursor[(size_t)lo * kSlPad], cap);
    const uint32_t off = (uint32_t)(c - pre[lo]) * (uint32_t)arena.chunk;
    w.first = r_lo + off;
    w.n = min((uint32_t)arena.chunk, cnt - off);
    return w;
}

// ---- S2 / I5: apply the probes.  SET = 1: dbgbf probes are test-and-set (graph.add / addDbgOnly) -------------------------------------
template <int SET>
__global__ void __launch_bounds__(kSlThreads) ks_apply_probes(const SlArena arena, const int* __restrict__ chunk_prefix, const SlGeom sg,
                                                             uint32_t* __restrict__ dbg_words, const uint32_t* __restrict__ cbf_words,
                                                             uint8_t* __restrict__ ans) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    int* pre = reinterpret_cast<int*>(sl_smem);
    sl_load_prefix(pre, chunk_prefix, arena.B);
    const int total = pre[arena.B];
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(arena.data);
    constexpr int U = 8;   // probes in flight per thread
    for (int c = blockIdx.x; c < total; c += gridDim.x) {
        const SlWork w = sl_work_item(arena, pre, c);
        const bool is_dbg = w.b < sg.n_dbg;
        const int64_t word0 = is_dbg ? ((int64_t)w.b << (sg.dbg_log2 - 5)) : ((int64_t)(w.b - sg.n_dbg) << (sg.cbf_log2 - 2));
        for (uint32_t i0 = threadIdx.x; i0 < w.n; i0 += kSlThreads * U) {
            uint32_t li[U], wd[U];
#pragma unroll
            for (int u = 0; u < U; ++u) li[u] = (i0 + u * kSlThreads < w.n) ? __ldcs(rec + w.first + i0 + u * kSlThreads) : 0u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                wd[u] = 0;
                if (i0 + u * kSlThreads < w.n) wd[u] = is_dbg ? ld_cg(dbg_words + word0 + (li[u] >> 5)) : ld_cg(cbf_words + word0 + (li[u] >> 2));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u * kSlThreads < w.n) {
                    uint32_t value;
                    if (is_dbg) {
                        const uint32_t bit = 1u << (li[u] & 31);
                        if (SET && !(wd[u] & bit)) wd[u] = atomicOr(dbg_words + word0 + (li[u] >> 5), bit);
                        value = (wd[u] & bit) ? 1u : 0u;
                    } else {
                        value = (wd[u] >> ((li[u] & 3) * 8)) & 0xFFu;
                    }
                    ans[w.first + i0 + u * kSlThreads] = (uint8_t)value;
                }
            }
        }
    }
}

// ---- S3: gather the answers of 4 k-mer instances per thread, write the counts (graph :562-570) ----------------------------------------
__global__ void __launch_bounds__(kSlThreads) ks_combine_lookup(const uint32_t* __restrict__ pos, const uint8_t* __restrict__ ans, int64_t n_inst, int hd,
                                                               int hc, float* __restrict__ counts, int64_t out_base) {
    const int64_t i0 = ((int64_t)blockIdx.x * kSlThreads + threadIdx.x) * kSlRoundKmers;
    if (i0 >= n_inst) return;
    uint32_t slot[kSlRoundKmers * kSlNJ], a[kSlRoundKmers * kSlNJ];
    sl_load_answers(pos, i0, ans, slot, a);
#pragma unroll
    for (int i = 0; i < kSlRoundKmers; ++i) {
        if (i0 + i < n_inst) {
            float c = 0.f;
            bool all = slot[i * kSlNJ] != kNoSlot;   // unusable k-mers (masked base in the window) made no probes
#pragma unroll
            for (int h = 0; h < kSlMaxH; ++h) if (h < hd) all = all && (a[i * kSlNJ + h] & 1u);
            if (all) {
                int mn = 127;
#pragma unroll
                for (int h = 0; h < kSlMaxH; ++h) if (h < hc) { const int v = (int)(int8_t)a[i * kSlNJ + kSlMaxH + h]; mn = v < mn ? v : mn; }
                c = minifloat_to_float(mn) + 1.f;
            }
            counts[out_base + i0 + i] = c;
        }
    }
}

// ---- I1: k-merise, tile-sort the base hash of every usable k-mer by key range -------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kSlThreads) ks_route_keys(const Ingest g, int k, int n_ranges, int range_shift, const SlArena arena, int* overflow) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    __shared__ RollLut lut;
    build_lut(&lut, k);
    TileSort<unsigned long long, kChunk> ts;
    ts.init(sl_smem, arena.B);
    const int64_t pos0 = ((int64_t)blockIdx.x * kSlThreads + threadIdx.x) * kChunk;
    const int n = pos0 < g.n_pos ? (int)min((int64_t)kChunk, g.n_pos - pos0) : 0;
    PositionWalker<MODE> pw;
    if (n) pw.start(g, pos0, k, lut);
    int bkt[kChunk];
    unsigned long long rec[kChunk];
    uint32_t slot[kChunk];
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        bkt[i] = -1; rec[i] = 0;
        if (i < n) {
            pw.advance(g, k, lut);
            if (pw.wk.bad == 0) {
                const uint64_t b = pw.wk.base();
                rec[i] = b;
                bkt[i] = n_ranges > 1 ? (int)(sl_mixkey(b) >> range_shift) : 0;
            }
        }
    }
    ts.run(arena, bkt, rec, slot, overflow);
}

// ---- I2: aggregate the keys range by range (the slice of the table a range maps to stays in L2 while its keys stream by) ----------------
__global__ void __launch_bounds__(kSlThreads) ks_aggregate(const SlArena arena, const int* __restrict__ chunk_prefix, const SlTable t) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    int* pre = reinterpret_cast<int*>(sl_smem);
    sl_load_prefix(pre, chunk_prefix, arena.B);
    const int total = pre[arena.B];
    const unsigned long long* rec = reinterpret_cast<const unsigned long long*>(arena.data);
    for (int c = blockIdx.x; c < total; c += gridDim.x) {
        const SlWork w = sl_work_item(arena, pre, c);
        constexpr int U = 4;   // keys in flight per thread: the CAS round trip to L2 is the cost
        for (uint32_t i0 = threadIdx.x; i0 < w.n; i0 += kSlThreads * U) {
            unsigned long long key[U], old[U];
            uint64_t s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) key[u] = (i0 + u * kSlThreads < w.n) ? __ldcs(rec + w.first + i0 + u * kSlThreads) : 0ULL;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                s[u] = sl_mixkey(key[u]) >> t.shift;
                old[u] = 0ULL;
                if (key[u] != 0ULL) old[u] = atomicCAS(&t.keys[s[u]], 0ULL, key[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u * kSlThreads >= w.n) continue;
                if (key[u] == 0ULL) { atomicAdd(&t.counts[t.n_slots], 1u); continue; }   // key 0 lives in the extra slot
                uint64_t at = s[u];
                unsigned long long o = old[u];
                while (o != 0ULL && o != key[u]) {   // linear probing
                    if (++at == t.n_slots) at = 0;
                    o = atomicCAS(&t.keys[at], 0ULL, key[u]);
                }
                atomicAdd(&t.counts[at], 1u);
            }
        }
    }
}

// ---- I3: occupied table slots -> dense (key, multiplicity) arrays ---------------------------------------------------------------------------
constexpr int kSlCompactPer = 8;
__global__ void __launch_bounds__(kSlThreads) ks_compact_table(const SlTable t, unsigned long long* __restrict__ dkey, unsigned int* __restrict__ dmult,
                                                              unsigned int* n_distinct) {
    __shared__ unsigned int tile_count, tile_base;
    const int64_t total = (int64_t)t.n_slots + 1;
    const int64_t tile0 = (int64_t)blockIdx.x * (kSlThreads * kSlCompactPer);
    if (threadIdx.x == 0) tile_count = 0;
    __syncthreads();
    unsigned int m[kSlCompactPer], rank[kSlCompactPer];
    unsigned long long key[kSlCompactPer];
#pragma unroll
    for (int i = 0; i < kSlCompactPer; ++i) {
        const int64_t s = tile0 + (int64_t)i * kSlThreads + threadIdx.x;
        m[i] = s < total ? t.counts[s] : 0u;
        rank[i] = 0; key[i] = 0;
        if (m[i]) {
            key[i] = (s == (int64_t)t.n_slots) ? 0ULL : t.keys[s];
            rank[i] = atomicAdd(&tile_count, 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) tile_base = tile_count ? atomicAdd(n_distinct, tile_count) : 0u;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSlCompactPer; ++i)
        if (m[i]) { dkey[tile_base + rank[i]] = key[i]; dmult[tile_base + rank[i]] = m[i]; }
}

// ---- I4: the probes of every distinct key, tile-sorted by filter slice ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kSlThreads) ks_emit_probes(const unsigned long long* __restrict__ dkey, const unsigned int* __restrict__ n_distinct,
                                                            const HashMults hm, const SlGeom sg, int with_cbf, const SlArena arena,
                                                            uint32_t* __restrict__ pos, int* overflow) {
    const int64_t nd = (int64_t)*n_distinct;
    if ((int64_t)blockIdx.x * (kSlThreads * kSlRoundKmers) >= nd) return;   // whole CTA
    RB_DYN_SMEM(unsigned char, sl_smem);
    TileSort<uint32_t, kSlRoundKmers * kSlNJ> ts;
    ts.init(sl_smem, arena.B);
    const int64_t d0 = ((int64_t)blockIdx.x * kSlThreads + threadIdx.x) * kSlRoundKmers;
    int bkt[kSlRoundKmers * kSlNJ];
    uint32_t rec[kSlRoundKmers * kSlNJ], slot[kSlRoundKmers * kSlNJ];
#pragma unroll
    for (int i = 0; i < kSlRoundKmers; ++i) {
#pragma unroll
        for (int j = 0; j < kSlNJ; ++j) { bkt[i * kSlNJ + j] = -1; rec[i * kSlNJ + j] = 0; }
        if (d0 + i < nd) sl_probes(sg, hm, (uint64_t)dkey[d0 + i], with_cbf != 0, &bkt[i * kSlNJ], &rec[i * kSlNJ]);
    }
    ts.run(arena, bkt, rec, slot, overflow);
    if (d0 < nd) sl_store_positions(pos, d0, slot);
}

// ---- I6: per distinct key: present?, replay the increments, emit one raise per counter that grew ----------------------------------------------
__global__ void __launch_bounds__(kSlThreads) ks_combine_insert(const unsigned long long* __restrict__ dkey, const unsigned int* __restrict__ dmult,
                                                               const unsigned int* __restrict__ n_distinct, const uint32_t* __restrict__ pos,
                                                               const uint8_t* __restrict__ ans, const HashMults hm, const SlGeom sg, int policy,
                                                               uint64_t rng_seed, const SlArena raises, int* overflow) {
    const int64_t nd = (int64_t)*n_distinct;
    if ((int64_t)blockIdx.x * (kSlThreads * kSlRoundKmers) >= nd) return;   // whole CTA
    RB_DYN_SMEM(unsigned char, sl_smem);
    TileSort<uint32_t, kSlRoundKmers * kSlMaxH> ts;
    ts.init(sl_smem, raises.B);
    const int64_t d0 = ((int64_t)blockIdx.x * kSlThreads + threadIdx.x) * kSlRoundKmers;
    int bkt[kSlRoundKmers * kSlMaxH];
    uint32_t rec[kSlRoundKmers * kSlMaxH], rslot[kSlRoundKmers * kSlMaxH];
#pragma unroll
    for (int e = 0; e < kSlRoundKmers * kSlMaxH; ++e) { bkt[e] = -1; rec[e] = 0; }
    if (d0 < nd) {
        uint32_t slot[kSlRoundKmers * kSlNJ], a[kSlRoundKmers * kSlNJ];
        sl_load_answers(pos, d0, ans, slot, a);
#pragma unroll
        for (int i = 0; i < kSlRoundKmers; ++i) {
            if (d0 + i < nd) {
                const uint64_t key = (uint64_t)dkey[d0 + i];
                const unsigned int m = dmult[d0 + i];
                bool present = true;
#pragma unroll
                for (int h = 0; h < kSlMaxH; ++h) if (h < sg.hd) present = present && (a[i * kSlNJ + h] & 1u);
                // graph.add :405-412 -- the first sighting of an absent k-mer only sets bits; addCountIfPresent :424-428 needs presence
                unsigned int n_inc = (policy == POLICY_COUNT_IF_PRESENT) ? (present ? m : 0u) : (m - 1u + (present ? 1u : 0u));
                int v0[kSlMaxH], v[kSlMaxH];
                uint64_t gi[kSlMaxH];
                int mn0 = 127;
#pragma unroll
                for (int h = 0; h < kSlMaxH; ++h) {
                    v0[h] = 127; gi[h] = ~0ULL;
                    if (h < sg.hc) {
                        v0[h] = (int)(a[i * kSlNJ + kSlMaxH + h] & 0x7Fu);
                        gi[h] = fm_index(expand_hash(key, h, hm), sg.cbf_fm);
                        mn0 = min(mn0, v0[h]);
                    }
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
#pragma unroll
                for (int h = 0; h < kSlMaxH; ++h) {
                    bool dup = false;
#pragma unroll
                    for (int h2 = 0; h2 < kSlMaxH; ++h2) if (h2 < h && gi[h2] == gi[h]) dup = true;   // one raise per distinct counter
                    if (h < sg.hc && !dup && v[h] > v0[h]) {
                        bkt[i * kSlMaxH + h] = (int)(gi[h] >> sg.raise_log2);
                        rec[i * kSlMaxH + h] = (uint32_t)(gi[h] & ((1ULL << sg.raise_log2) - 1)) | ((uint32_t)v[h] << sg.raise_log2);
                    }
                }
            }
        }
    }
    ts.run(raises, bkt, rec, rslot, overflow);
}

// ---- I7: raise the counters slice by slice ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSlThreads) ks_apply_raises(const SlArena arena, const int* __restrict__ chunk_prefix, const SlGeom sg,
                                                             uint32_t* __restrict__ cbf_words) {
    RB_DYN_SMEM(unsigned char, sl_smem);
    int* pre = reinterpret_cast<int*>(sl_smem);
    sl_load_prefix(pre, chunk_prefix, arena.B);
    const int total = pre[arena.B];
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(arena.data);
    for (int c = blockIdx.x; c < total; c += gridDim.x) {
        const SlWork w = sl_work_item(arena, pre, c);
        const int64_t word0 = (int64_t)w.b << (sg.raise_log2 - 2);
        for (uint32_t i = threadIdx.x; i < w.n; i += kSlThreads) {
            const uint32_t a = __ldcs(rec + w.first + i);
            const uint32_t li = a & ((1u << sg.raise_log2) - 1u);
            uint32_t* wp = cbf_words + word0 + (li >> 2);
            byte_raise(wp, (int)(li & 3) * 8, a >> sg.raise_log2, ld_cg(wp));
        }
    }
}

}  // namespace rb
