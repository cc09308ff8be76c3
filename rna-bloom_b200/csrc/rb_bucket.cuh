// rb_bucket.cuh -- the bucketed engine: same operators, but no isolated random HBM probes.
//
// Measured on B200 (profiles/r01_notes.md): isolated random 32 B probes top out at 47.5 G/s (21.6 G/s for atomics) because every probe
// costs a DRAM row activation, while the same operations confined to an L2-resident slice run at 280 G/s (loads) / 126-190 G/s (atomics)
// and streaming runs at ~6 TB/s.  So every probe is first *partitioned by the 16 MiB filter slice it falls into* (streaming writes),
// then the slices are visited one after the other: a slice is prefetched sequentially, all of its probes hit L2, and the answers are
// partitioned back by k-mer id the same way.  The logical bit/byte arrays are untouched -- this is a schedule, not a blocked Bloom filter.
//
//   insert (graph.add, graph/BloomFilterDeBruijnGraph.java:405-412):
//     B1 route keys by key range           -> B2 aggregate duplicates per range into an L2-resident table slice (key -> multiplicity)
//     B3 emit h_d test-and-set + h_c counter-read probes per distinct key, partitioned by filter slice
//     B4 apply slice by slice, partition the answers by table-slot range (16 Ki slots)
//     B5 per slot range: gather the answers in shared memory; present = AND(old bits); replay m-1+present min-increments on the h_c
//        counter values (CountingBloomFilter.java:170-194); emit raises partitioned by counter slice -> B6 apply raises slice by slice
//   lookup (graph.getKmers / getCount, :562-570): B7 emit probes per k-mer instance -> B4 -> B8 per instance range: gather, count
//
// Partitioning is a counting sort with a persistent grid: pass 0 histograms the records every CTA will produce (shared memory), a scan
// turns the (region, CTA) histogram into private write cursors, pass 1 recomputes the records and writes them -- no global cursor
// atomics, exact sizes, dense regions.  Record formats (64-bit): probe = slice-local index [0,28) | id [28,60) | j [60,63);
// answer (32-bit) = id inside its range [0,14) | j [14,17) | value [17,25); raise (32-bit) = slice-local byte index [0,24) | value [24,31).
// j = hash number (dbgbf: 0..h_d-1, cbf: h_d..h_d+h_c-1).
#pragma once
#include "rb_kernels.cuh"

namespace rb {

constexpr int kSliceBitsLog2 = 27;        // dbgbf slice: 2^27 bits  = 16 MiB (two slices in flight + the next one prefetched + the record
constexpr int kSliceBytesLog2 = 24;       // cbf slice:   2^24 bytes = 16 MiB  streams must fit the ~63 MB one L2 partition keeps)
constexpr int kIdRangeLog2 = 14;          // ids per answer range: 16 Ki ids * 8 answer bytes = 128 KiB of shared memory
constexpr int kTableRangeLog2 = 20;       // table slots per key range: 1 Mi * 12 B = 12 MiB
constexpr int kSub = 8;                   // cursor-scatter only (keys, raises): sub-regions per region
constexpr int kMaxCursorRegions = 2048;

// ---- dense regions produced by the counting sort ---------------------------------------------------------------------------------
// Writers are grouped (group = blockIdx % C): all CTAs of a group append to the same tail of a region through one global cursor.
// Few open write streams (R*C, chosen around 16-32 Ki) let L2 merge the 4-8 byte stores into full sectors before they reach HBM,
// and that many cursors spread the atomics enough to run at L2 speed (measured 126 G atomics/s when spread).
constexpr int kCursorPad = 32;  // one cursor per 128 B line: atomics to one line serialise in L2 (measured: 16 counters in one line = 1.3 G/s total)
struct SortPlan {
    unsigned int* hist;      // [R * C] pass 0: records group c produces for region r; after the scan: that group's offset inside the region
    unsigned int* cursor;    // [R * C * kCursorPad] pass 1: the live cursors (copied from hist, padded)
    int64_t* roff;           // [R + 1] region offsets (records) into data
    void* data;
    int R, C;
};
__global__ void __launch_bounds__(256) kb_spread_cursors(const unsigned int* __restrict__ hist, int n, unsigned int* __restrict__ cursor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cursor[(int64_t)i * kCursorPad] = hist[i];
}
template <typename REC, int PASS>
struct SortWriter {
    unsigned int* bins;      // pass 0: shared-memory histogram of one CTA [R]
    int group;
    __device__ __forceinline__ void begin(unsigned int* smem, const SortPlan& p) {
        bins = smem;
        group = blockIdx.x % p.C;
        if (!PASS) {
            for (int r = threadIdx.x; r < p.R; r += blockDim.x) bins[r] = 0u;
            __syncthreads();
        }
    }
    __device__ __forceinline__ void put(const SortPlan& p, int region, REC rec) {
        if (!PASS) atomicAdd(&bins[region], 1u);
        else reinterpret_cast<REC*>(p.data)[p.roff[region] + atomicAdd(&p.cursor[((int64_t)region * p.C + group) * kCursorPad], 1u)] = rec;
    }
    // N records at once: all cursor atomics are issued before the first dependent store (the atomic round trip is the cost)
    template <int N>
    __device__ __forceinline__ void put_batch(const SortPlan& p, const int* region, const REC* rec) {   // region < 0: no record
        if (!PASS) {
#pragma unroll
            for (int i = 0; i < N; ++i) if (region[i] >= 0) atomicAdd(&bins[region[i]], 1u);
        } else {
            int64_t at[N];
#pragma unroll
            for (int i = 0; i < N; ++i)
                if (region[i] >= 0) at[i] = p.roff[region[i]] + atomicAdd(&p.cursor[((int64_t)region[i] * p.C + group) * kCursorPad], 1u);
#pragma unroll
            for (int i = 0; i < N; ++i) if (region[i] >= 0) reinterpret_cast<REC*>(p.data)[at[i]] = rec[i];
        }
    }
    __device__ __forceinline__ void end(const SortPlan& p) {
        if (!PASS) {
            __syncthreads();
            for (int r = threadIdx.x; r < p.R; r += blockDim.x) if (bins[r]) atomicAdd(&p.hist[(int64_t)r * p.C + group], bins[r]);
        }
    }
};
// scan step 1: one CTA per region turns its C counts into exclusive offsets and leaves the region total in tot[r]
__global__ void __launch_bounds__(1024) kb_scan_regions(unsigned int* __restrict__ hist, int C, int64_t* __restrict__ tot) {
    __shared__ unsigned int warp_sum[32];
    const int r = blockIdx.x, t = threadIdx.x;
    unsigned int v = t < C ? hist[(int64_t)r * C + t] : 0u, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if ((t & 31) >= o) x += y; }
    if ((t & 31) == 31) warp_sum[t >> 5] = x;
    __syncthreads();
    if (t < 32) {
        unsigned int w = warp_sum[t], z = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, z, o); if (t >= o) z += y; }
        warp_sum[t] = z - w;
        if (t == 31) tot[r] = (int64_t)z;
    }
    __syncthreads();
    if (t < C) hist[(int64_t)r * C + t] = x - v + warp_sum[t >> 5];
}
// scan step 2: region totals -> region offsets (single CTA; R <= 64 Ki)
__global__ void __launch_bounds__(1024) kb_scan_totals(const int64_t* __restrict__ tot, int R, int64_t* __restrict__ roff) {
    __shared__ int64_t part[1024];
    const int t = threadIdx.x;
    const int per = (R + 1023) / 1024;
    int64_t s = 0;
    for (int i = t * per; i < min(R, (t + 1) * per); ++i) s += tot[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) { int64_t acc = 0; for (int i = 0; i < 1024; ++i) { const int64_t v = part[i]; part[i] = acc; acc += v; } roff[R] = acc; }
    __syncthreads();
    int64_t acc = part[t];
    for (int i = t * per; i < min(R, (t + 1) * per); ++i) { roff[i] = acc; acc += tot[i]; }
}

// ---- loose grid barrier: the persistent CTAs may be at most two regions apart (keeps the L2 working set to a few slices) ------------
__device__ __forceinline__ void region_arrive(unsigned int* done, int r) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(&done[r], 1u); }
}
__device__ __forceinline__ void region_wait(unsigned int* done, int r) {
    if (r >= 0) {
        if (threadIdx.x == 0) { while (*(volatile unsigned int*)&done[r] < gridDim.x) __nanosleep(100); }
        __syncthreads();
    }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// every CTA pulls its share of [ptr, ptr+bytes) into L2 (128 B lines)
__device__ __forceinline__ void cta_prefetch(const char* ptr, int64_t bytes) {
    const int64_t lines = (bytes + 127) >> 7;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < lines; l += (int64_t)gridDim.x * blockDim.x) prefetch_l2(ptr + (l << 7));
}

struct BucketGeom {
    FastMod dbg_fm, cbf_fm;      // global index arithmetic (reference semantics)
    int hd, hc;
    int n_dbg_slices, n_cbf_slices;
    int n_key_ranges, key_range_shift;      // key range = mixed key >> key_range_shift (n_key_ranges == 1: range 0)
    int n_id_ranges;                        // id range = id >> kIdRangeLog2 (ids: table slots for insert, instances for lookup)
};
__device__ __forceinline__ uint64_t mixkey(uint64_t key) { return key * 0x9E3779B97F4A7C15ULL; }
__device__ __forceinline__ uint64_t make_probe(uint64_t local_idx, uint64_t id, int j) { return local_idx | (id << 28) | ((uint64_t)j << 60); }
__device__ __forceinline__ uint32_t make_answer(uint64_t id, int j, uint32_t value) { return (uint32_t)(id & ((1u << kIdRangeLog2) - 1)) | ((uint32_t)j << kIdRangeLog2) | (value << (kIdRangeLog2 + 3)); }
// region of a probe: dbgbf slices first, then cbf slices
__device__ __forceinline__ void probe_of(const BucketGeom& bg, const HashMults& hm, uint64_t key, int j, uint64_t id, int* region, uint64_t* rec) {
    if (j < bg.hd) {
        const uint64_t gi = fm_index(expand_hash(key, j, hm), bg.dbg_fm);
        *region = (int)(gi >> kSliceBitsLog2);
        *rec = make_probe(gi & ((1ULL << kSliceBitsLog2) - 1), id, j);
    } else {
        const uint64_t gi = fm_index(expand_hash(key, j - bg.hd, hm), bg.cbf_fm);
        *region = bg.n_dbg_slices + (int)(gi >> kSliceBytesLog2);
        *rec = make_probe(gi & ((1ULL << kSliceBytesLog2) - 1), id, j);
    }
}

// ---- cursor scatter (keys and raises only: few records per item, 8 sub-regions per region spread the cursor atomics) --------------
struct Regions {
    void* data;
    unsigned int* count;    // [R * kSub * kCursorPad] one live counter per 128 B line
    int64_t cap;            // records per sub-region
    int n;                  // R
};
struct RegionView {          // flat view of one region
    int64_t pre[kSub + 1];
    __device__ __forceinline__ void open(const Regions& in, int r) {
        pre[0] = 0;
#pragma unroll
        for (int s = 0; s < kSub; ++s) pre[s + 1] = pre[s] + min((int64_t)in.count[(r * kSub + s) * kCursorPad], in.cap);
    }
    __device__ __forceinline__ int64_t size() const { return pre[kSub]; }
    __device__ __forceinline__ int64_t at(const Regions& in, int r, int64_t i) const {
        int s = 0;
#pragma unroll
        for (int q = 1; q < kSub; ++q) s += (i >= pre[q]) ? 1 : 0;
        return ((int64_t)r * kSub + s) * in.cap + (i - pre[s]);
    }
};
template <typename REC, int E>
struct CtaScatter {
    unsigned int* hist;     // smem [R]
    unsigned int* base;     // smem [R]
    REC rec[E];
    unsigned int meta[E];   // region (low 12 bits, 0xFFF = none) | rank inside the CTA's share << 12
    __device__ __forceinline__ void begin(unsigned int* smem, int R) {
        hist = smem; base = smem + R;
        for (int r = threadIdx.x; r < R; r += blockDim.x) hist[r] = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) meta[e] = 0xFFFu;
        __syncthreads();
    }
    __device__ __forceinline__ void put(int e, int reg, REC r) { rec[e] = r; meta[e] = (unsigned int)reg | (atomicAdd(&hist[reg], 1u) << 12); }
    __device__ __forceinline__ void flush(const Regions& out, int* overflow) {   // all threads of the CTA must call it
        __syncthreads();
        const int sub = blockIdx.x % kSub;
        for (int r = threadIdx.x; r < out.n; r += blockDim.x) { const unsigned int c = hist[r]; base[r] = c ? atomicAdd(&out.count[(r * kSub + sub) * kCursorPad], c) : 0u; }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e)
            if ((meta[e] & 0xFFFu) != 0xFFFu) {
                const int reg = (int)(meta[e] & 0xFFFu);
                const int64_t p = (int64_t)base[reg] + (meta[e] >> 12);
                if (p < out.cap) __stcs(reinterpret_cast<REC*>(out.data) + ((int64_t)reg * kSub + sub) * out.cap + p, rec[e]);
                else *overflow = 1;
            }
        __syncthreads();
    }
};

// ---- B1: k-merise, scatter every usable k-mer's base hash by key range ---------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads) kb_route_keys(const Ingest g, int k, const BucketGeom bg, const Regions out, int* overflow) {
    RB_DYN_SMEM(unsigned int, smem);
    __shared__ RollLut lut;
    build_lut(&lut, k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    const int n = pos < g.n_pos ? (int)min((int64_t)kChunk, g.n_pos - pos) : 0;
    PositionWalker<MODE> pw;
    if (n) pw.start(g, pos, k, lut);
    CtaScatter<uint64_t, kChunk> sc;
    sc.begin(smem, out.n);
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        if (i < n) {
            pw.advance(g, k, lut);
            if (pw.wk.bad == 0) {
                const uint64_t b = pw.wk.base();
                sc.put(i, bg.n_key_ranges > 1 ? (int)(mixkey(b) >> bg.key_range_shift) : 0, b);
            }
        }
    }
    sc.flush(out, overflow);
}

// ---- B2: aggregate the keys of each range into its slice of the table (persistent grid, ranges in lock-step) -------------------------
struct AggTable2 {
    unsigned long long* keys;   // T + 1 slots, 0 = empty; slot T stands for key 0
    unsigned int* counts;
    uint64_t n_slots;           // T (power of two)
    int shift;                  // slot = mixkey >> shift
    uint64_t zero_slot;         // = T
};
__global__ void __launch_bounds__(kThreads) kb_aggregate(const Regions in, const AggTable2 t, unsigned int* done) {
    const int64_t slots_per_range = (int64_t)(t.n_slots / (uint64_t)in.n);
    for (int r = 0; r < in.n; ++r) {
        region_wait(done, r - 2);
        if (r + 1 < in.n) {   // next range's table slice (performance only; the table was zeroed beforehand)
            cta_prefetch((const char*)(t.keys + (int64_t)(r + 1) * slots_per_range), slots_per_range * 8);
            cta_prefetch((const char*)(t.counts + (int64_t)(r + 1) * slots_per_range), slots_per_range * 4);
        }
        RegionView rv;
        rv.open(in, r);
        const int64_t cnt = rv.size();
        const unsigned long long* rec = reinterpret_cast<const unsigned long long*>(in.data);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
            const uint64_t key = __ldcs(rec + rv.at(in, r, i));
            if (key == 0) { atomicAdd(&t.counts[t.zero_slot], 1u); continue; }
            uint64_t s = mixkey(key) >> t.shift;
            for (;;) {
                const unsigned long long old = atomicCAS(&t.keys[s], 0ULL, (unsigned long long)key);
                if (old == 0ULL || old == key) { atomicAdd(&t.counts[s], 1u); break; }
                if (++s == t.zero_slot) s = 0;
            }
        }
        region_arrive(done, r);
    }
}

// ---- B3: one probe per (distinct key, hash); counting sort by filter slice -------------------------------------------------------------
template <int MAXJ, int PASS>   // MAXJ >= hd + hc
__global__ void __launch_bounds__(kThreads) kb_emit_probes(const AggTable2 t, const HashMults hm, const BucketGeom bg, int with_cbf, const SortPlan plan) {
    RB_DYN_SMEM(unsigned int, smem);
    SortWriter<uint64_t, PASS> sw;
    sw.begin(smem, plan);
    const int64_t total = (int64_t)t.zero_slot + 1;
    const int nj = bg.hd + (with_cbf ? bg.hc : 0);
    for (int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x; s < total; s += (int64_t)gridDim.x * kThreads) {
        if (t.counts[s] != 0) {
            const uint64_t key = (s == (int64_t)t.zero_slot) ? 0ULL : (uint64_t)t.keys[s];
            int reg[MAXJ]; uint64_t rec[MAXJ];
#pragma unroll
            for (int j = 0; j < MAXJ; ++j) { reg[j] = -1; rec[j] = 0; if (j < nj) probe_of(bg, hm, key, j, (uint64_t)s, &reg[j], &rec[j]); }
            sw.template put_batch<MAXJ>(plan, reg, rec);
        }
    }
    sw.end(plan);
}

// ---- B7: one probe per (usable k-mer instance, hash); pass 1 also writes the hashes getKmers returns ----------------------------------
template <int MODE, int MAXJ, int PASS>
__global__ void __launch_bounds__(kThreads) kb_route_lookup(const Ingest g, int k, const HashMults hm, const BucketGeom bg, const SortPlan plan,
                                                           uint8_t* __restrict__ usable, int64_t* __restrict__ fhash, int64_t* __restrict__ rhash) {
    RB_DYN_SMEM(unsigned int, smem);
    __shared__ RollLut lut;
    build_lut(&lut, k);
    SortWriter<uint64_t, PASS> sw;
    sw.begin(smem, plan);
    const int nj = bg.hd + bg.hc;
    const int64_t tile = (int64_t)kThreads * kChunk;
    for (int64_t t0 = (int64_t)blockIdx.x * tile; t0 < g.n_pos; t0 += (int64_t)gridDim.x * tile) {
        const int64_t pos = t0 + (int64_t)threadIdx.x * kChunk;
        const int n = pos < g.n_pos ? (int)min((int64_t)kChunk, g.n_pos - pos) : 0;
        if (n) {
            PositionWalker<MODE> pw;
            pw.start(g, pos, k, lut);
            for (int i = 0; i < n; ++i) {
                pw.advance(g, k, lut);
                const int64_t inst = pos + i;
                const bool ok = pw.wk.bad == 0;
                if (PASS) {
                    if (fhash) fhash[g.out_base + inst] = (int64_t)pw.wk.f;
                    if (rhash) rhash[g.out_base + inst] = (int64_t)pw.wk.r;
                    usable[inst] = ok ? 1 : 0;
                }
                if (ok) {
                    const uint64_t b = pw.wk.base();
                    int reg[MAXJ]; uint64_t rec[MAXJ];
#pragma unroll
                    for (int j = 0; j < MAXJ; ++j) { reg[j] = -1; rec[j] = 0; if (j < nj) probe_of(bg, hm, b, j, (uint64_t)inst, &reg[j], &rec[j]); }
                    sw.template put_batch<MAXJ>(plan, reg, rec);
                }
            }
        }
    }
    sw.end(plan);
}

// ---- B4: apply the probes slice by slice; answers are counting-sorted by id range --------------------------------------------------------
// PASS 0 only histograms (id range of every probe a CTA will answer); PASS 1 probes the filters and writes the answers.
// SET = 1: dbgbf probes are test-and-set (graph.add / addDbgOnly), 0: read only (lookup, addCountIfPresent)
template <int PASS, int SET>
__global__ void __launch_bounds__(kThreads) kb_apply_probes(const SortPlan in, const BucketGeom bg, uint32_t* __restrict__ dbg_words,
                                                           const uint32_t* __restrict__ cbf_words, int64_t dbg_bytes, int64_t cbf_bytes,
                                                           int want_answers, const SortPlan out, unsigned int* done) {
    RB_DYN_SMEM(unsigned int, smem);
    constexpr int U = 4;   // probes in flight per thread
    SortWriter<uint32_t, PASS> sw;
    if (want_answers) sw.begin(smem, out);
    const unsigned long long* rec = reinterpret_cast<const unsigned long long*>(in.data);
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int r = 0; r < in.R; ++r) {
        const int64_t lo = in.roff[r], hi = in.roff[r + 1];
        if (PASS) {
            region_wait(done, r - 2);
            if (r + 1 < in.R) {   // pull the next slice into L2 while this one is being probed
                const bool nd = r + 1 < bg.n_dbg_slices;
                const int64_t off = nd ? ((int64_t)(r + 1) << (kSliceBitsLog2 - 3)) : ((int64_t)(r + 1 - bg.n_dbg_slices) << kSliceBytesLog2);
                const int64_t len = min((int64_t)1 << kSliceBytesLog2, (nd ? dbg_bytes : cbf_bytes) - off);
                if (len > 0 && (in.roff[r + 2] - hi) > (len >> 7)) cta_prefetch((nd ? (const char*)dbg_words : (const char*)cbf_words) + off, len);
            }
        }
        const bool is_dbg = r < bg.n_dbg_slices;
        const int64_t word0 = is_dbg ? ((int64_t)r << (kSliceBitsLog2 - 5)) : ((int64_t)(r - bg.n_dbg_slices) << (kSliceBytesLog2 - 2));
        for (int64_t i0 = lo + (int64_t)blockIdx.x * kThreads + threadIdx.x; i0 < hi; i0 += stride * U) {
            uint64_t p[U];
#pragma unroll
            for (int u = 0; u < U; ++u) p[u] = (i0 + u * stride < hi) ? __ldcs(rec + i0 + u * stride) : ~0ULL;
            int areg[U]; uint32_t arec[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { areg[u] = (p[u] != ~0ULL && want_answers) ? (int)(((p[u] >> 28) & 0xFFFFFFFFULL) >> kIdRangeLog2) : -1; arec[u] = 0; }
            if (!PASS) { sw.template put_batch<U>(out, areg, arec); continue; }
            uint32_t w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                w[u] = 0;
                if (p[u] != ~0ULL) {
                    const uint64_t li = p[u] & ((1ULL << 28) - 1);
                    w[u] = is_dbg ? ld_cg(dbg_words + word0 + (li >> 5)) : ld_cg(cbf_words + word0 + (li >> 2));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (p[u] != ~0ULL) {
                    const uint64_t li = p[u] & ((1ULL << 28) - 1), id = (p[u] >> 28) & 0xFFFFFFFFULL;
                    const int j = (int)(p[u] >> 60) & 7;
                    uint32_t value;
                    if (is_dbg) {
                        const uint32_t bit = 1u << (li & 31);
                        if (SET && !(w[u] & bit)) w[u] = atomicOr(dbg_words + word0 + (li >> 5), bit);
                        value = (w[u] & bit) ? 1u : 0u;
                    } else {
                        value = (w[u] >> ((li & 3) * 8)) & 0xFFu;
                    }
                    arec[u] = make_answer(id, j, value);
                }
            }
            sw.template put_batch<U>(out, areg, arec);
        }
        if (PASS) region_arrive(done, r);
    }
    if (want_answers) sw.end(out);
}

// gathers the answers of one id range into shared memory: ans[(id - first)*8 + j]
__device__ __forceinline__ void gather_answers(const SortPlan& in, int r, uint8_t* ans_smem) {
    const int n_words = (1 << kIdRangeLog2) * 2;
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) reinterpret_cast<uint32_t*>(ans_smem)[i] = 0u;
    __syncthreads();
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(in.data);
    const int64_t lo = in.roff[r], hi = in.roff[r + 1];
#pragma unroll 4
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint32_t a = __ldcs(rec + i);
        ans_smem[((a & ((1u << kIdRangeLog2) - 1)) << 3) + ((a >> kIdRangeLog2) & 7)] = (uint8_t)(a >> (kIdRangeLog2 + 3));
    }
    __syncthreads();
}

// ---- B5: per table-slot range: present?, replay the increments, emit raises (cursor scatter by counter slice) -------------------------------
constexpr int kCombineThreads = 1024;
template <int MAXH>
__global__ void __launch_bounds__(kCombineThreads) kb_combine_insert(const SortPlan in, const AggTable2 t, const HashMults hm, const BucketGeom bg, int policy,
                                                             uint64_t rng_seed, const Regions out, int* overflow) {
    RB_DYN_SMEM(unsigned int, smem);
    uint8_t* ans = reinterpret_cast<uint8_t*>(smem);                         // 128 KiB
    unsigned int* sc_mem = smem + (1 << kIdRangeLog2) * 2;                   // scatter bins behind it
    const int64_t total = (int64_t)t.zero_slot + 1;
    for (int r = blockIdx.x; r < in.R; r += gridDim.x) {
        gather_answers(in, r, ans);
        for (int s0 = 0; s0 < (1 << kIdRangeLog2); s0 += kCombineThreads) {
            CtaScatter<uint32_t, MAXH> sc;
            sc.begin(sc_mem, out.n);
            const int sl = s0 + threadIdx.x;
            const int64_t s = ((int64_t)r << kIdRangeLog2) + sl;
            const unsigned int m = s < total ? t.counts[s] : 0u;
            if (m) {
                const uint64_t key = (s == (int64_t)t.zero_slot) ? 0ULL : (uint64_t)t.keys[s];
                const uint2 a2 = *reinterpret_cast<const uint2*>(ans + (sl << 3));
                const uint64_t a = (uint64_t)a2.x | ((uint64_t)a2.y << 32);
                bool present = true;
                for (int h = 0; h < bg.hd; ++h) present = present && ((a >> (8 * h)) & 1);
                // graph.add :405-412 -- the first sighting of an absent k-mer only sets bits; addCountIfPresent :424-428 needs presence
                unsigned int n_inc = (policy == POLICY_COUNT_IF_PRESENT) ? (present ? m : 0u) : (m - 1u + (present ? 1u : 0u));
                int v0[MAXH], v[MAXH];
                uint64_t gi[MAXH];
                int mn0 = 127;
#pragma unroll
                for (int h = 0; h < MAXH; ++h) {
                    v0[h] = 127; gi[h] = ~0ULL;
                    if (h < bg.hc) { v0[h] = (int)((a >> (8 * (bg.hd + h))) & 0x7F); gi[h] = fm_index(expand_hash(key, h, hm), bg.cbf_fm); mn0 = min(mn0, v0[h]); }
                    v[h] = v0[h];
                }
                if (policy == POLICY_COUNT_IF_PRESENT && mn0 == 0) n_inc = 0;   // "&& cbf.getCount(hashVals) > 0" (graph :425)
                uint64_t rr = mix64(key ^ rng_seed);
                for (unsigned int it = 0; it < n_inc; ++it) {   // CountingBloomFilter.increment :170-194, n_inc times
                    int mn = 127;
#pragma unroll
                    for (int h = 0; h < MAXH; ++h) if (h < bg.hc) mn = min(mn, v[h]);
                    if (mn >= 127) break;
                    rr = mix64(rr + it);
                    const int u = minifloat_increment(mn, rr);
                    if (u != mn) {
#pragma unroll
                        for (int h = 0; h < MAXH; ++h) if (h < bg.hc && v[h] == mn) v[h] = u;
                    }
                }
#pragma unroll
                for (int h = 0; h < MAXH; ++h) {
                    bool dup = false;
#pragma unroll
                    for (int h2 = 0; h2 < MAXH; ++h2) if (h2 < h && gi[h2] == gi[h]) dup = true;   // one raise per distinct counter
                    if (h < bg.hc && !dup && v[h] > v0[h])
                        sc.put(h, (int)(gi[h] >> kSliceBytesLog2), (uint32_t)(gi[h] & ((1ULL << kSliceBytesLog2) - 1)) | ((uint32_t)v[h] << kSliceBytesLog2));
                }
            }
            sc.flush(out, overflow);
        }
        __syncthreads();
    }
}

// ---- B6: raise counters slice by slice --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) kb_apply_raises(const Regions in, uint32_t* __restrict__ cbf_words, int64_t cbf_bytes, unsigned int* done) {
    for (int r = 0; r < in.n; ++r) {
        region_wait(done, r - 2);
        RegionView rv;
        rv.open(in, r);
        const int64_t cnt = rv.size();
        if (r + 1 < in.n) {
            const int64_t off = (int64_t)(r + 1) << kSliceBytesLog2;
            const int64_t len = min((int64_t)1 << kSliceBytesLog2, cbf_bytes - off);
            int64_t next = 0;
#pragma unroll
            for (int q = 0; q < kSub; ++q) next += in.count[((r + 1) * kSub + q) * kCursorPad];
            if (len > 0 && next > (len >> 7)) cta_prefetch((const char*)cbf_words + off, len);
        }
        const uint32_t* rec = reinterpret_cast<const uint32_t*>(in.data);
        const int64_t word0 = (int64_t)r << (kSliceBytesLog2 - 2);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
            const uint32_t a = __ldcs(rec + rv.at(in, r, i));
            const uint32_t li = a & ((1u << kSliceBytesLog2) - 1);
            uint32_t* wp = cbf_words + word0 + (li >> 2);
            byte_raise(wp, (int)(li & 3) * 8, a >> kSliceBytesLog2, ld_cg(wp));
        }
        region_arrive(done, r);
    }
}

// ---- B8: per instance range: gather the answers, write the counts (graph :562-570) ---------------------------------------------------------------
__global__ void __launch_bounds__(kCombineThreads) kb_combine_lookup(const SortPlan in, int64_t n_inst, int hd, int hc, const uint8_t* __restrict__ usable,
                                                             float* __restrict__ counts, int64_t out_base) {
    RB_DYN_SMEM(unsigned int, smem);
    uint8_t* ans = reinterpret_cast<uint8_t*>(smem);
    for (int r = blockIdx.x; r < in.R; r += gridDim.x) {
        gather_answers(in, r, ans);
        for (int sl = threadIdx.x; sl < (1 << kIdRangeLog2); sl += kCombineThreads) {
            const int64_t i = ((int64_t)r << kIdRangeLog2) + sl;
            if (i >= n_inst) break;
            float c = 0.f;
            if (usable[i]) {
                const uint2 a2 = *reinterpret_cast<const uint2*>(ans + (sl << 3));
                const uint64_t a = (uint64_t)a2.x | ((uint64_t)a2.y << 32);
                bool all = true;
                for (int h = 0; h < hd; ++h) all = all && ((a >> (8 * h)) & 1);
                if (all) {
                    int mn = 127;
                    for (int h = 0; h < hc; ++h) { const int v = (int)(int8_t)(a >> (8 * (hd + h))); mn = v < mn ? v : mn; }
                    c = minifloat_to_float(mn) + 1.f;
                }
            }
            counts[out_base + i] = c;
        }
        __syncthreads();
    }
}

}  // namespace rb
