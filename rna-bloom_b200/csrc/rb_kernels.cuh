// rb_kernels.cuh -- the CUDA kernels of the hot path (sm_100a).  One thread owns kChunk consecutive k-mer positions:
// it seeds the rolling ntHash once, then per group of kGroup k-mers issues every filter probe before consuming any
// (memory-level parallelism is what a random-sector workload lives on).  Citations: /root/reference/src/rnabloom/.
#pragma once
#include "rb_device.cuh"

namespace rb {

constexpr int kThreads = 256;

struct GraphDev {
    BitFilter dbg;
    ByteFilter cbf;
    HashMults hm;
    ClaimTable ct;
    uint64_t rng_seed;
    int k;
};

enum { POLICY_ADD = 0, POLICY_COUNT_IF_PRESENT = 1, POLICY_DBG_ONLY = 2 };

// Walks the positions [pos, pos+n) of the launch, crossing read boundaries when needed.
template <int MODE>
struct PositionWalker {
    KmerWalker<MODE> wk;
    int64_t read;
    int32_t in_read, npos;
    bool primed;
    __device__ __forceinline__ void start(const Ingest& g, int64_t pos, int k, const RollLut& lut) {
        locate(g, pos, read, in_read);
        npos = read_npos(g, read);
        wk.init(g, read_start(g, read) + in_read, k, lut);
        primed = true;
    }
    // moves to the next position (no-op on the very first call)
    __device__ __forceinline__ void advance(const Ingest& g, int k, const RollLut& lut) {
        if (primed) { primed = false; return; }
        if (++in_read < npos) { wk.roll(lut); return; }
        do { ++read; npos = read_npos(g, read); } while (npos <= 0);
        in_read = 0;
        wk.init(g, read_start(g, read), k, lut);
    }
};

// ---- graph.add / addCountIfPresent / addDbgOnly over reads (graph/BloomFilterDeBruijnGraph.java:405-436) ----------
// Every stage issues the memory operations of all kGroup k-mers before it looks at any result, so a group costs a handful of
// HBM/L2 round trips instead of one dependent chain per k-mer:
//   1 dbgbf probes            2 claim CAS (k-mers with a clear bit)  + red.or of the clear bits
//   3 cbf probes              4 lock+bump CAS of the designated counter -> raise CAS of the other minimum counters -> unlock
// Anything unusual (claim-slot collision, locked or concurrently changed counter) falls back to the generic loops of rb_device.cuh.
template <int MODE, int MAXH, int POLICY>
__global__ void __launch_bounds__(kThreads) k_graph_insert(const Ingest g, const GraphDev gd) {
    __shared__ RollLut lut;
    build_lut(&lut, gd.k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (pos >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - pos);
    PositionWalker<MODE> pw;
    pw.start(g, pos, gd.k, lut);

    for (int i0 = 0; i0 < n; i0 += kGroup) {
        uint64_t base[kGroup];
        uint32_t wd[kGroup][MAXH];
        uint32_t ok = 0;
        // stage 1: hash kGroup k-mers and put every dbgbf probe in flight
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
            if (i0 + j < n) {
                pw.advance(g, gd.k, lut);
                base[j] = pw.wk.base();
                if (pw.wk.bad == 0) {
                    ok |= 1u << j;
#pragma unroll
                    for (int h = 0; h < MAXH; ++h)
                        if (h < gd.dbg.num_hash) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.dbg.fm);
                            wd[j][h] = ld_cg(&gd.dbg.words[idx >> 5]);
                        }
                }
            }
        }
        // stage 2: dbgbf.lookupThenAdd (bloom/BloomFilter.java:147-155)
        uint32_t found = 0, need_claim = 0;
        uint32_t clear[kGroup];
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
            clear[j] = 0;
            if (ok & (1u << j)) {
#pragma unroll
                for (int h = 0; h < MAXH; ++h)
                    if (h < gd.dbg.num_hash) {
                        const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.dbg.fm);
                        if (!((wd[j][h] >> (idx & 31)) & 1u)) clear[j] |= 1u << h;
                    }
                if (clear[j] == 0) found |= 1u << j;
                else if (POLICY != POLICY_COUNT_IF_PRESENT) need_claim |= 1u << j;
            }
        }
        if (POLICY == POLICY_ADD && need_claim) {
            unsigned long long got[kGroup];
            uint64_t slot[kGroup];
#pragma unroll
            for (int j = 0; j < kGroup; ++j)   // all first-probe claims in flight together
                if ((need_claim & (1u << j)) && base[j] != 0) {
                    slot[j] = (base[j] * 0x9E3779B97F4A7C15ULL) >> gd.ct.shift;
                    got[j] = atomicCAS(&gd.ct.slots[slot[j]], 0ULL, (unsigned long long)base[j]);
                }
#pragma unroll
            for (int j = 0; j < kGroup; ++j)
                if (need_claim & (1u << j)) {
                    bool first;
                    if (base[j] == 0) first = claim_first(gd.ct, 0);
                    else if (got[j] == 0ULL) first = true;
                    else if (got[j] == base[j]) first = false;
                    else first = claim_first_from(gd.ct, base[j], (slot[j] + 1) & gd.ct.mask);
                    if (!first) { found |= 1u << j; need_claim &= ~(1u << j); }  // a concurrent duplicate owns the first sighting
                }
        }
        if (POLICY != POLICY_COUNT_IF_PRESENT) {
#pragma unroll
            for (int j = 0; j < kGroup; ++j)
                if (need_claim & (1u << j)) {
#pragma unroll
                    for (int h = 0; h < MAXH; ++h)
                        if (clear[j] & (1u << h)) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.dbg.fm);
                            atomicOr(&gd.dbg.words[idx >> 5], 1u << (idx & 31));
                        }
                }
        }
        if constexpr (POLICY != POLICY_DBG_ONLY) {
            if (!found) continue;
            // stage 3: cbf probes of the present k-mers in flight
            uint32_t wc[kGroup][MAXH];
#pragma unroll
            for (int j = 0; j < kGroup; ++j)
                if (found & (1u << j)) {
#pragma unroll
                    for (int h = 0; h < MAXH; ++h)
                        if (h < gd.cbf.num_hash) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.cbf.fm);
                            wc[j][h] = ld_cg(&gd.cbf.words[idx >> 2]);
                        }
                }
            // stage 4a: pick minimum / designated counter, issue every lock+bump CAS (rb_device.cuh "cbf_increment" protocol)
            uint32_t fast = 0, slow = 0, raise[kGroup], gotw[kGroup];
            int u[kGroup], dsg[kGroup];
#pragma unroll
            for (int j = 0; j < kGroup; ++j) {
                raise[j] = 0; u[j] = 0; dsg[j] = 0; gotw[j] = 0;
                if (found & (1u << j)) {
                    uint32_t lock = 0;
                    int mn = 127, v[MAXH];
                    uint64_t idx[MAXH];
#pragma unroll
                    for (int h = 0; h < MAXH; ++h) {
                        v[h] = 127; idx[h] = 0;
                        if (h < gd.cbf.num_hash) {
                            idx[h] = fm_index(expand_hash(base[j], h, gd.hm), gd.cbf.fm);
                            const uint32_t b = (wc[j][h] >> ((int)(idx[h] & 3) * 8)) & 0xFFu;
                            lock |= b & kLockBit;
                            v[h] = (int)(b & 0x7Fu);
                            mn = v[h] < mn ? v[h] : mn;
                        }
                    }
                    if (POLICY == POLICY_COUNT_IF_PRESENT && mn == 0) continue;   // "&& cbf.getCount(hashVals) > 0" (graph :425)
                    if (lock) { slow |= 1u << j; continue; }
                    const uint64_t rk = mix64(base[j] ^ gd.rng_seed) + (uint64_t)(pos + i0 + j) * 0x632BE59BD9B4E019ULL;
                    u[j] = minifloat_increment(mn, mix64(rk));
                    if (u[j] == mn) continue;
                    int D = 0;
#pragma unroll
                    for (int h = 0; h < MAXH; ++h) if (h < gd.cbf.num_hash && v[h] == mn) D = h;
                    uint64_t idxD = 0; uint32_t wD = 0;
#pragma unroll
                    for (int h = 0; h < MAXH; ++h) if (h == D) { idxD = idx[h]; wD = wc[j][h]; }
#pragma unroll
                    for (int h = 0; h < MAXH; ++h)
                        if (h < gd.cbf.num_hash && h != D && v[h] == mn && idx[h] != idxD) raise[j] |= 1u << h;
                    dsg[j] = D;
                    const int sh = (int)(idxD & 3) * 8;
                    const uint32_t nw = (wD & ~(0xFFu << sh)) | (((uint32_t)u[j] | kLockBit) << sh);
                    gotw[j] = atomicCAS(&gd.cbf.words[idxD >> 2], wD, nw);
                    if (true) fast |= 1u << j;
                }
            }
            // stage 4b: lock holders raise the other minimum counters (all raise CAS in flight), everybody else goes the slow way
            uint32_t gr[kGroup][MAXH];
#pragma unroll
            for (int j = 0; j < kGroup; ++j)
                if (fast & (1u << j)) {
                    uint32_t wD = 0;
#pragma unroll
                    for (int h = 0; h < MAXH; ++h) if (h == dsg[j]) wD = wc[j][h];
                    if (gotw[j] != wD) { fast &= ~(1u << j); slow |= 1u << j; continue; }
#pragma unroll
                    for (int h = 0; h < MAXH; ++h)
                        if (raise[j] & (1u << h)) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.cbf.fm);
                            const int sh = (int)(idx & 3) * 8;
                            const uint32_t old = wc[j][h];
                            gr[j][h] = atomicCAS(&gd.cbf.words[idx >> 2], old, (old & ~(0x7Fu << sh)) | ((uint32_t)u[j] << sh));
                        }
                }
            // stage 4c: finish raises that lost a race on another byte of their word, then release the locks
#pragma unroll
            for (int j = 0; j < kGroup; ++j)
                if (fast & (1u << j)) {
#pragma unroll
                    for (int h = 0; h < MAXH; ++h)
                        if ((raise[j] & (1u << h)) && gr[j][h] != wc[j][h]) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.cbf.fm);
                            byte_raise(&gd.cbf.words[idx >> 2], (int)(idx & 3) * 8, (uint32_t)u[j], gr[j][h]);
                        }
                    uint64_t idxD = 0;
#pragma unroll
                    for (int h = 0; h < MAXH; ++h) if (h == dsg[j]) idxD = fm_index(expand_hash(base[j], h, gd.hm), gd.cbf.fm);
                    atomicAnd(&gd.cbf.words[idxD >> 2], ~(kLockBit << ((int)(idxD & 3) * 8)));
                }
            // slow path only after this thread has released every lock it held
#pragma unroll
            for (int j = 0; j < kGroup; ++j)
                if (slow & (1u << j))
                    cbf_increment<MAXH>(gd.cbf, base[j], gd.hm, mix64(base[j] ^ gd.rng_seed) + (uint64_t)(pos + i0 + j) * 0x632BE59BD9B4E019ULL + 1, nullptr);
        }
    }
}

// ---- graph.getKmers / getCount over reads (graph :562-570, :1224-1226; HashFunction.java:55-85) --------------------
template <int MODE, int MAXH>
__global__ void __launch_bounds__(kThreads) k_graph_count(const Ingest g, const GraphDev gd, float* __restrict__ counts,
                                                         int64_t* __restrict__ fhash, int64_t* __restrict__ rhash) {
    __shared__ RollLut lut;
    build_lut(&lut, gd.k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (pos >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - pos);
    PositionWalker<MODE> pw;
    pw.start(g, pos, gd.k, lut);
    for (int i0 = 0; i0 < n; i0 += kGroup) {
        uint64_t base[kGroup];
        uint32_t wd[kGroup][MAXH], wc[kGroup][MAXH];
        uint32_t ok = 0;
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
            if (i0 + j < n) {
                pw.advance(g, gd.k, lut);
                base[j] = pw.wk.base();
                const int64_t o = g.out_base + pos + i0 + j;
                if (fhash) fhash[o] = (int64_t)pw.wk.f;
                if (rhash) rhash[o] = (int64_t)pw.wk.r;
                if (pw.wk.bad == 0) {
                    ok |= 1u << j;
#pragma unroll
                    for (int h = 0; h < MAXH; ++h) {
                        if (h < gd.dbg.num_hash) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.dbg.fm);
                            wd[j][h] = ld_cg(&gd.dbg.words[idx >> 5]);
                        }
                        if (h < gd.cbf.num_hash) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.cbf.fm);
                            wc[j][h] = ld_cg(&gd.cbf.words[idx >> 2]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
            if (i0 + j < n) {
                float c = 0.f;
                if (ok & (1u << j)) {
                    bool all = true;
                    int mn = 127;
#pragma unroll
                    for (int h = 0; h < MAXH; ++h) {
                        if (h < gd.dbg.num_hash) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.dbg.fm);
                            all = all && ((wd[j][h] >> (idx & 31)) & 1u);
                        }
                        if (h < gd.cbf.num_hash) {
                            const uint64_t idx = fm_index(expand_hash(base[j], h, gd.hm), gd.cbf.fm);
                            const int v = byte_of(wc[j][h], (int)(idx & 3) * 8);
                            mn = v < mn ? v : mn;
                        }
                    }
                    if (all) c = minifloat_to_float(mn) + 1.f;
                }
                if (counts) counts[g.out_base + pos + i0 + j] = c;
            }
        }
    }
}

// ---- the k-merizer alone: NTHashIterator family (bloom/hash/NTHashIterator.java:47-69 and twins) --------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_kmerize(const Ingest g, int k, int64_t* __restrict__ fhash, int64_t* __restrict__ rhash,
                                                     int64_t* __restrict__ base) {
    __shared__ RollLut lut;
    build_lut(&lut, k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (pos >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - pos);
    PositionWalker<MODE> pw;
    pw.start(g, pos, k, lut);
    for (int i = 0; i < n; ++i) {
        pw.advance(g, k, lut);
        const int64_t o = g.out_base + pos + i;
        if (fhash) fhash[o] = (int64_t)pw.wk.f;
        if (rhash) rhash[o] = (int64_t)pw.wk.r;
        if (base) base[o] = (int64_t)pw.wk.base();
    }
}

// ---- paired k-mers: Paired*NTHashIterator (bloom/hash/PairedNTHashIterator.java:55-85, Canonical...:39-60, RC...:35-56)
// positions are pair positions (len-k-d+1 per read).  PAIR_OP: 0 = write hValsP[0]; 1 = pkbf.add (graph :455-461);
// 2 = pkbf.add only when both k-mers are in dbgbf (RNABloom.java:389-399); 3 = pkbf.lookup (graph.lookupReadKmerPair / lookupFragmentKmerPair
// :526-532, the test inside breakWith{Read,Frag}PairedKmers, util/GraphUtils.java:4184-4310): one byte per pair position, written through
// pair_out seen as a byte array
__device__ __forceinline__ int count_masked(const uint32_t* mask, int64_t start, int n) {
    if (!mask) return 0;
    int c = 0;
    for (int64_t b = start; b < start + n;) {
        const int sh = (int)(b & 31);
        const int take = min(32 - sh, (int)(start + n - b));
        const uint32_t w = __ldg(&mask[b >> 5]) >> sh;
        c += __popc(take == 32 ? w : (w & ((1u << take) - 1u)));
        b += take;
    }
    return c;
}
template <int MODE, int MAXH, int PAIR_OP>
__global__ void __launch_bounds__(kThreads) k_pairs(const Ingest g, const GraphDev gd, const BitFilter pk, int d,
                                                   int64_t* __restrict__ pair_out) {
    __shared__ RollLut lut;
    build_lut(&lut, gd.k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (pos >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - pos);
    const int k = gd.k;
    int64_t read; int32_t in_read;
    locate(g, pos, read, in_read);
    int32_t npos = read_npos(g, read);
    KmerWalker<MODE> L, R;
    int span_bad = 0;
    bool primed = false;
    for (int i = 0; i < n; ++i) {
        if (!primed || in_read >= npos) {
            if (primed) { do { ++read; npos = read_npos(g, read); } while (npos <= 0); in_read = 0; }
            const int64_t s = read_start(g, read) + in_read;
            L.init(g, s, k, lut);
            R.init(g, s + d, k, lut);
            span_bad = count_masked(g.mask, s, k + d);
            primed = true;
        } else {
            // the base leaving the span is L's out base, the base entering is R's in base
            const int64_t bo = L.out.b, bi = R.in.b;
            if (g.mask) span_bad += (int)((__ldg(&g.mask[bi >> 5]) >> (bi & 31)) & 1u) - (int)((__ldg(&g.mask[bo >> 5]) >> (bo & 31)) & 1u);
            L.roll(lut);
            R.roll(lut);
        }
        uint64_t p;
        if (MODE == 0) p = combine_hash(L.f, R.f);
        else if (MODE == 1) p = combine_hash(R.r, L.r);
        else {
            const uint64_t p1 = combine_hash(L.f, R.f), p2 = combine_hash(R.r, L.r);
            p = ((int64_t)p2 < (int64_t)p1) ? p2 : p1;  // Math.min on long
        }
        if (PAIR_OP == 0) pair_out[g.out_base + pos + i] = (int64_t)p;
        else if (PAIR_OP == 3) reinterpret_cast<uint8_t*>(pair_out)[g.out_base + pos + i] = (span_bad == 0 && bf_lookup<MAXH>(pk, p, gd.hm)) ? 1 : 0;
        else if (span_bad == 0) {
            bool go = true;
            if (PAIR_OP == 2) go = bf_lookup<MAXH>(gd.dbg, L.base(), gd.hm) && bf_lookup<MAXH>(gd.dbg, R.base(), gd.hm);
            if (go) bf_add<MAXH>(pk, p, gd.hm);
        }
        ++in_read;
    }
}

// ---- f4: a lone Bloom filter over whole sequences -- the screening filter of the assembly stages (RNABloom.java:1680,2530-2536,4264;
// util/GraphUtils.java:627-650).  SEQ_OP 0: bf.add(kmer.getHash()) for every k-mer; 1: containsAllKmers; 2: lookupAndAddAllKmers.
// missing[r] is set when a k-mer of read r was not found, or is unusable ("kmer == null" :635); the caller has cleared it.  A read without
// any k-mer is the host's business (containsAllKmers :629-631 returns false, lookupAndAddAllKmers true).
enum { SEQ_ADD = 0, SEQ_CONTAINS_ALL = 1, SEQ_LOOKUP_AND_ADD_ALL = 2 };
template <int MODE, int MAXH, int SEQ_OP>
__global__ void __launch_bounds__(kThreads) k_seq_filter(const Ingest g, int k, const BitFilter bf, const HashMults hm, uint8_t* __restrict__ missing) {
    __shared__ RollLut lut;
    build_lut(&lut, k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (pos >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - pos);
    PositionWalker<MODE> pw;
    pw.start(g, pos, k, lut);
    for (int i = 0; i < n; ++i) {
        pw.advance(g, k, lut);
        bool found = false;
        if (pw.wk.bad == 0) {
            const uint64_t b = pw.wk.base();
            if (SEQ_OP == SEQ_ADD) { bf_add<MAXH>(bf, b, hm); found = true; }
            else {
                found = bf_lookup<MAXH>(bf, b, hm);
                // lookupThenAdd (bloom/BloomFilter.java:147-155) returns whether every bit was set before; two copies of a k-mer in one
                // batch may both report "absent" -- the reference's worker threads race in the same way on the shared screening filter
                if (SEQ_OP == SEQ_LOOKUP_AND_ADD_ALL && !found) bf_add<MAXH>(bf, b, hm);
            }
        }
        if (SEQ_OP != SEQ_ADD && !found) missing[g.read_base + pw.read] = 1;
    }
}

// ---- minimizers: MinimizerHashIterator.next() for every window of w consecutive k-mers (bloom/hash/MinimizerHashIterator.java:42-101 over
// util/LongRollingWindow.java:43-73: the signed minimum of hVals[0] over the window) -- the key generator of SeqSubsampler.minimizerBased
// (util/SeqSubsampler.java:50-117).  Positions are window positions (len - k - w + 2 per read); a thread walks kChunk consecutive windows
// with a ring of the last w hashes.  w <= kMaxMinimizerWindow.
constexpr int kMaxMinimizerWindow = 64;
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_minimizers(const Ingest g, int k, int w, int64_t* __restrict__ out) {
    __shared__ RollLut lut;
    build_lut(&lut, k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (pos >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - pos);
    int64_t read; int32_t in_read;
    locate(g, pos, read, in_read);
    int32_t npos = read_npos(g, read);
    KmerWalker<MODE> wk;
    int64_t ring[kMaxMinimizerWindow];
    int head = 0;   // ring[head] is the oldest hash of the window
    bool primed = false;
    for (int i = 0; i < n; ++i) {
        if (!primed || in_read >= npos) {
            if (primed) { do { ++read; npos = read_npos(g, read); } while (npos <= 0); in_read = 0; }
            wk.init(g, read_start(g, read) + in_read, k, lut);
            for (int j = 0; j < w; ++j) { if (j) wk.roll(lut); ring[j] = (int64_t)wk.base(); }
            head = 0;
            primed = true;
        } else {
            wk.roll(lut);
            ring[head] = (int64_t)wk.base();
            head = head + 1 == w ? 0 : head + 1;
        }
        int64_t mn = ring[0];
        for (int j = 1; j < w; ++j) mn = ring[j] < mn ? ring[j] : mn;
        out[g.out_base + pos + i] = mn;
        ++in_read;
    }
}

// ---- f3: k-mer multiplicity histogram by hash sampling -- what the reference shells out to `ntcard` for (RNABloom.java:5745-5768; the
// histogram it parses: util/NTCardHistogram.java:33-63).  A k-mer is sampled when the top `sample_bits` bits of a multiplicative mix of
// its hash are zero (ntCard samples on the hash in the same way); sampled k-mers are counted EXACTLY in an open-addressing table, so the
// histogram of the sample is exact and only the scaling by 2^sample_bits is an estimate.  One in 2^sample_bits k-mers reaches the table.
struct CardTable {
    unsigned long long* keys;     // n_slots + 1; 0 = empty; slot n_slots stands for key 0
    unsigned int* counts;
    uint64_t n_slots;             // power of two
    int shift;                    // 64 - log2(n_slots)
    int sample_bits;
    unsigned long long* totals;   // [0] usable k-mers seen (F1)  [1] sampled instances  [2] overflow flag
};
__device__ __forceinline__ uint64_t card_mix(uint64_t key) { return key * 0x9E3779B97F4A7C15ULL; }
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_card_add(const Ingest g, int k, const CardTable ct) {
    __shared__ RollLut lut;
    build_lut(&lut, k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (pos >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - pos);
    PositionWalker<MODE> pw;
    pw.start(g, pos, k, lut);
    unsigned int usable = 0, sampled = 0;
    for (int i = 0; i < n; ++i) {
        pw.advance(g, k, lut);
        if (pw.wk.bad) continue;
        ++usable;
        const uint64_t key = pw.wk.base(), m = card_mix(key);
        if (ct.sample_bits && (m >> (64 - ct.sample_bits)) != 0) continue;
        ++sampled;
        if (key == 0ULL) { atomicAdd(&ct.counts[ct.n_slots], 1u); continue; }
        uint64_t s = (m << ct.sample_bits) >> ct.shift;   // the bits below the sampling prefix
        for (uint64_t tries = 0;; ++tries) {
            const unsigned long long old = atomicCAS(&ct.keys[s], 0ULL, (unsigned long long)key);
            if (old == 0ULL || old == key) { atomicAdd(&ct.counts[s], 1u); break; }
            if (tries >= ct.n_slots) { atomicExch(&ct.totals[2], 1ULL); break; }   // table full
            s = (s + 1) & (ct.n_slots - 1);
        }
    }
    if (usable) atomicAdd(&ct.totals[0], (unsigned long long)usable);
    if (sampled) atomicAdd(&ct.totals[1], (unsigned long long)sampled);
}
// hist[m - 1] += 1 for every sampled distinct k-mer of multiplicity m <= max_mult; hist[max_mult] counts the ones above
__global__ void __launch_bounds__(kThreads) k_card_hist(const CardTable ct, unsigned long long* __restrict__ hist, int max_mult) {
    for (uint64_t s = (uint64_t)blockIdx.x * kThreads + threadIdx.x; s <= ct.n_slots; s += (uint64_t)gridDim.x * kThreads) {
        const unsigned int c = ct.counts[s];
        if (c) atomicAdd(&hist[c <= (unsigned int)max_mult ? c - 1 : max_mult], 1ULL);
    }
}

// ---- per-hash operators: the `long hashVal` overloads (bloom/BloomFilter.java:139-182, CountingBloomFilter.java:126-251) -
enum { OP_BF_ADD = 0, OP_BF_LOOKUP, OP_BF_LTA, OP_CBF_INC, OP_CBF_INC_GET, OP_CBF_COUNT, OP_GRAPH_ADD, OP_GRAPH_COUNT_IF_PRESENT,
       OP_GRAPH_COUNT };
template <int MAXH, int OP>
__global__ void __launch_bounds__(kThreads) k_hash_op(const int64_t* __restrict__ base, int64_t n, const GraphDev gd,
                                                     uint8_t* __restrict__ out8, float* __restrict__ outf) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const uint64_t b = (uint64_t)base[i];
    if (OP == OP_BF_ADD) bf_add<MAXH>(gd.dbg, b, gd.hm);
    else if (OP == OP_BF_LOOKUP) out8[i] = bf_lookup<MAXH>(gd.dbg, b, gd.hm) ? 1 : 0;
    else if (OP == OP_BF_LTA) out8[i] = bf_lookup_then_add<MAXH>(gd.dbg, gd.ct, b, gd.hm) ? 1 : 0;
    else if (OP == OP_CBF_INC) cbf_increment<MAXH>(gd.cbf, b, gd.hm, (mix64(b ^ gd.rng_seed) + (uint64_t)i * 0x632BE59BD9B4E019ULL), nullptr);
    else if (OP == OP_CBF_INC_GET) outf[i] = minifloat_to_float(cbf_increment<MAXH>(gd.cbf, b, gd.hm, (mix64(b ^ gd.rng_seed) + (uint64_t)i * 0x632BE59BD9B4E019ULL), nullptr));
    else if (OP == OP_CBF_COUNT) outf[i] = minifloat_to_float(cbf_min<MAXH>(gd.cbf, b, gd.hm));
    else if (OP == OP_GRAPH_ADD) {
        if (bf_lookup_then_add<MAXH>(gd.dbg, gd.ct, b, gd.hm)) cbf_increment<MAXH>(gd.cbf, b, gd.hm, (mix64(b ^ gd.rng_seed) + (uint64_t)i * 0x632BE59BD9B4E019ULL), nullptr);
    } else if (OP == OP_GRAPH_COUNT_IF_PRESENT) {
        if (bf_lookup<MAXH>(gd.dbg, b, gd.hm) && cbf_min<MAXH, true>(gd.cbf, b, gd.hm) > 0)   // graph :424-428; a locked slot is not a smaller one
            cbf_increment<MAXH>(gd.cbf, b, gd.hm, (mix64(b ^ gd.rng_seed) + (uint64_t)i * 0x632BE59BD9B4E019ULL), nullptr);
    } else if (OP == OP_GRAPH_COUNT) {
        outf[i] = bf_lookup<MAXH>(gd.dbg, b, gd.hm) ? minifloat_to_float(cbf_min<MAXH>(gd.cbf, b, gd.hm)) + 1.f : 0.f;
    }
}

// ---- the reference's packed fragment records (io/NucleotideBitsWriter.java:24-31, util/SeqBitsUtils.java:138-262) -> ingest layout -------
// A record = 4-byte big-endian length + ceil(len / 4) tetramer bytes; a tetramer byte = (b0 * 64 + b1 * 16 + b2 * 4 + b3) - 128, first base in
// the top two bits.  One thread per 32-base output word: 8 tetramer bytes -> one 64-bit word with base b at bits 2 * (b & 31).
__global__ void __launch_bounds__(kThreads) k_unpack_2bit(const uint8_t* __restrict__ records, const int64_t* __restrict__ data_off /* first tetramer byte of read r */,
                                                         const int32_t* __restrict__ read_len, const int64_t* __restrict__ word_off, int64_t n_reads,
                                                         int64_t n_words, uint64_t* __restrict__ packed) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= n_words) return;
    int64_t lo = 0, hi = n_reads;  // read whose word range contains t
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&word_off[mid]) <= t) lo = mid; else hi = mid;
    }
    const int len = __ldg(&read_len[lo]);
    const int first = (int)(t - __ldg(&word_off[lo])) * 32;          // first base of this word inside the read
    const uint8_t* src = records + __ldg(&data_off[lo]) + first / 4;
    const int n_bytes = min(8, (len - first + 3) / 4);
    uint64_t w = 0;
    for (int j = 0; j < n_bytes; ++j) {
        const uint32_t v = (uint32_t)src[j] ^ 0x80u;                  // + 128 (BYTE_OFFSET, SeqBitsUtils.java:35)
        const uint64_t four = (uint64_t)((v >> 6) & 3) | (uint64_t)((v >> 4) & 3) << 2 | (uint64_t)((v >> 2) & 3) << 4 | (uint64_t)(v & 3) << 6;
        w |= four << (8 * j);
    }
    packed[t] = w;
}

// ---- CascadingBloomFilter (bloom/CascadingBloomFilter.java:66-100): the keys a level reported as already present move on to the next one
__global__ void __launch_bounds__(kThreads) k_cascade_survivors(const int64_t* __restrict__ keys, const int32_t* __restrict__ idx, const uint8_t* __restrict__ found,
                                                               int64_t n, int64_t* __restrict__ keys_out, int32_t* __restrict__ idx_out, unsigned int* n_out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n || !found[i]) return;
    const unsigned int o = atomicAdd(n_out, 1u);
    keys_out[o] = keys[i];
    idx_out[o] = idx ? idx[i] : (int32_t)i;
}
__global__ void __launch_bounds__(kThreads) k_scatter_ones(const int32_t* __restrict__ idx, int64_t n, uint8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i < n) out[idx[i]] = 1;
}

// ---- f1 neighbours of a k-mer: Kmer.getSuccessors / getPredecessors (graph/Kmer.java:213-253), CanonicalKmer (:232-271) ----------------
// Hash side: bloom/hash/SuccessorsNTHashIterator.java:52-63, PredecessorsNTHashIterator.java:54-65 and the Canonical twins (:56-72).
// One thread per (k-mer, direction): the hashes of the 4 candidate neighbours (A,C,G,T) and graph.getCount of each, all
// 4 * (h_d + h_c) probes in flight before any is consumed.  out code = first base (successors) / last base (predecessors).
template <int MAXH>
__global__ void __launch_bounds__(kThreads) k_neighbors(const int64_t* __restrict__ fhash, const int64_t* __restrict__ rhash,
                                                       const uint8_t* __restrict__ first_code, const uint8_t* __restrict__ last_code, int64_t n,
                                                       const GraphDev gd, int canonical, float* __restrict__ counts, int64_t* __restrict__ nf,
                                                       int64_t* __restrict__ nr) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= 2 * n) return;
    const int64_t q = t >> 1;
    const bool succ = (t & 1) == 0;   // outputs: [q][0] successors, [q][1] predecessors, 4 entries each
    const int k = gd.k;
    const int out = (int)(succ ? first_code[q] : last_code[q]) & 3;
    const uint64_t f = (uint64_t)fhash[q], r = canonical ? (uint64_t)rhash[q] : 0ULL;
    const uint64_t tf = succ ? (rotl1(f) ^ rotl64(seed_of_code(out), k)) : (rotr1(f) ^ rotl64(seed_of_code(out), 63));
    const uint64_t tr = succ ? (rotr1(r) ^ rotl64(seed_of_code(3 - out), 63)) : (rotl1(r) ^ rotl64(seed_of_code(3 - out), k));
    uint64_t base[4], fn[4], rn[4];
    uint32_t wd[4][MAXH], wc[4][MAXH];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fn[c] = tf ^ (succ ? seed_of_code(c) : rotl64(seed_of_code(c), k - 1));
        rn[c] = canonical ? (tr ^ (succ ? rotl64(seed_of_code(3 - c), k - 1) : seed_of_code(3 - c))) : 0ULL;
        base[c] = (canonical && (int64_t)rn[c] < (int64_t)fn[c]) ? rn[c] : fn[c];
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
            if (h < gd.dbg.num_hash) wd[c][h] = ld_cg(&gd.dbg.words[fm_index(expand_hash(base[c], h, gd.hm), gd.dbg.fm) >> 5]);
            if (h < gd.cbf.num_hash) wc[c][h] = ld_cg(&gd.cbf.words[fm_index(expand_hash(base[c], h, gd.hm), gd.cbf.fm) >> 2]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        bool all = true;
        int mn = 127;
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
            if (h < gd.dbg.num_hash) { const uint64_t idx = fm_index(expand_hash(base[c], h, gd.hm), gd.dbg.fm); all = all && ((wd[c][h] >> (idx & 31)) & 1u); }
            if (h < gd.cbf.num_hash) { const uint64_t idx = fm_index(expand_hash(base[c], h, gd.hm), gd.cbf.fm); const int v = byte_of(wc[c][h], (int)(idx & 3) * 8); mn = v < mn ? v : mn; }
        }
        const int64_t o = t * 4 + c;
        counts[o] = all ? minifloat_to_float(mn) + 1.f : 0.f;   // graph.getCount :562-570
        if (nf) nf[o] = (int64_t)fn[c];
        if (nr) nr[o] = (int64_t)rn[c];
    }
}

// graph.getCount (:562-570) of one base hash: every probe in flight before any is consumed
template <int MAXH>
__device__ __forceinline__ float graph_count_of(const GraphDev& gd, uint64_t base) {
    uint32_t wd[MAXH], wc[MAXH];
    uint64_t id[MAXH], ic[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        if (h < gd.dbg.num_hash) { id[h] = fm_index(expand_hash(base, h, gd.hm), gd.dbg.fm); wd[h] = ld_cg(&gd.dbg.words[id[h] >> 5]); }
        if (h < gd.cbf.num_hash) { ic[h] = fm_index(expand_hash(base, h, gd.hm), gd.cbf.fm); wc[h] = ld_cg(&gd.cbf.words[ic[h] >> 2]); }
    }
    bool all = true;
    int mn = 127;
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        if (h < gd.dbg.num_hash) all = all && ((wd[h] >> (id[h] & 31)) & 1u);
        if (h < gd.cbf.num_hash) { const int v = byte_of(wc[h], (int)(ic[h] & 3) * 8); mn = v < mn ? v : mn; }
    }
    return all ? minifloat_to_float(mn) + 1.f : 0.f;
}

// ---- f1 variants: Kmer.getLeftVariants / getRightVariants (graph/Kmer.java:357-405), CanonicalKmer (:381-519); hash side
// bloom/hash/LeftVariantsNTHashIterator.java:40-46, RightVariantsNTHashIterator.java:38-44 and the Canonical twins (:42-52): the k-mers that
// differ from the query in its FIRST (side 0) or LAST (side 1) base.  One thread per (k-mer, side); entry c of the 4 outputs is the variant
// with base c there (c == the k-mer's own base: the k-mer itself); the caller drops that entry and applies "count >= minKmerCov".
template <int MAXH>
__global__ void __launch_bounds__(kThreads) k_variants(const int64_t* __restrict__ fhash, const int64_t* __restrict__ rhash, const uint8_t* __restrict__ first_code,
                                                      const uint8_t* __restrict__ last_code, int64_t n, const GraphDev gd, int canonical,
                                                      float* __restrict__ counts, int64_t* __restrict__ vf, int64_t* __restrict__ vr) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= 2 * n) return;
    const int64_t q = t >> 1;
    const bool left = (t & 1) == 0;   // outputs: [q][0] left variants, [q][1] right variants, 4 entries each
    const int k = gd.k;
    const int out = (int)(left ? first_code[q] : last_code[q]) & 3;
    const uint64_t f = (uint64_t)fhash[q], r = canonical ? (uint64_t)rhash[q] : 0ULL;
    // f = xor_i rotl(S[s_i], k-1-i),  r = xor_i rotl(S[3-s_i], i): base 0 sits at rotation k-1 (forward) / 0 (reverse), base k-1 the other way
    const int rot_f = left ? k - 1 : 0, rot_r = left ? 0 : k - 1;
    const uint64_t tf = f ^ rotl64(seed_of_code(out), rot_f), tr = r ^ rotl64(seed_of_code(3 - out), rot_r);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint64_t fn = tf ^ rotl64(seed_of_code(c), rot_f);
        const uint64_t rn = canonical ? (tr ^ rotl64(seed_of_code(3 - c), rot_r)) : 0ULL;
        const uint64_t base = (canonical && (int64_t)rn < (int64_t)fn) ? rn : fn;
        const int64_t o = t * 4 + c;
        counts[o] = graph_count_of<MAXH>(gd, base);
        if (vf) vf[o] = (int64_t)fn;
        if (vr) vr[o] = (int64_t)rn;
    }
}

// ---- f1 batched greedy extension: GraphUtils.greedyExtendRight / greedyExtendLeft with lookahead <= 1 (util/GraphUtils.java:501-527,
// 1961-1976): at every step the successor (predecessor) with the largest count >= min_cov, the first of equal ones in A, C, G, T order
// (Kmer.getMaxCovSuccessor, graph/Kmer.java:301-327), until there is none or `bound` k-mers were added.  One thread per start k-mer;
// the k-mer (k <= 64) travels as 2-bit codes in two 64-bit words, base i at bits 2 * (i & 31) of word i >> 5.
template <int MAXH>
__global__ void __launch_bounds__(kThreads) k_greedy_extend(const uint64_t* __restrict__ kmer_bits, const int64_t* __restrict__ fhash, const int64_t* __restrict__ rhash,
                                                           int64_t n, const GraphDev gd, int canonical, int right, int bound, float min_cov,
                                                           int32_t* __restrict__ ext_len, uint8_t* __restrict__ ext_codes, float* __restrict__ ext_counts) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= n) return;
    const int k = gd.k;
    uint64_t w0 = kmer_bits[2 * t], w1 = kmer_bits[2 * t + 1];
    uint64_t f = (uint64_t)fhash[t], r = canonical ? (uint64_t)rhash[t] : 0ULL;
    int len = 0;
    for (; len < bound; ++len) {
        // the base that leaves: the first one when extending to the right, the last one to the left
        const int po = right ? 0 : k - 1;
        const int out = (int)(((po < 32 ? w0 : w1) >> (2 * (po & 31))) & 3);
        const uint64_t tf = right ? (rotl1(f) ^ rotl64(seed_of_code(out), k)) : (rotr1(f) ^ rotl64(seed_of_code(out), 63));
        const uint64_t tr = right ? (rotr1(r) ^ rotl64(seed_of_code(3 - out), 63)) : (rotl1(r) ^ rotl64(seed_of_code(3 - out), k));
        float best = -1.f;
        int best_c = -1;
        uint64_t best_f = 0, best_r = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint64_t fn = tf ^ (right ? seed_of_code(c) : rotl64(seed_of_code(c), k - 1));
            const uint64_t rn = canonical ? (tr ^ (right ? rotl64(seed_of_code(3 - c), k - 1) : seed_of_code(3 - c))) : 0ULL;
            const uint64_t base = (canonical && (int64_t)rn < (int64_t)fn) ? rn : fn;
            const float cnt = graph_count_of<MAXH>(gd, base);
            if (cnt >= min_cov && cnt > best) { best = cnt; best_c = c; best_f = fn; best_r = rn; }
        }
        if (best_c < 0) break;
        ext_codes[t * bound + len] = (uint8_t)best_c;
        if (ext_counts) ext_counts[t * bound + len] = best;
        f = best_f; r = best_r;
        if (right) {            // shift the window one base to the right: drop base 0, append at k - 1
            w0 = (w0 >> 2) | (w1 << 62);
            w1 >>= 2;
            const int pi = k - 1;
            if (pi < 32) w0 = (w0 & ~(3ULL << (2 * pi))) | ((uint64_t)best_c << (2 * pi));
            else w1 = (w1 & ~(3ULL << (2 * (pi & 31)))) | ((uint64_t)best_c << (2 * (pi & 31)));
        } else {                // to the left: prepend at 0, drop base k - 1
            w1 = (w1 << 2) | (w0 >> 62);
            w0 = (w0 << 2) | (uint64_t)best_c;
            const int pd = k;   // the base that fell off now sits at position k: clear it
            if (pd < 32) w0 &= ~(3ULL << (2 * pd));
            else if (pd < 64) w1 &= ~(3ULL << (2 * (pd & 31)));
        }
    }
    ext_len[t] = len;
}

// ---- a8 getIndex exposed on its own (bloom/BloomFilter.java:108-111) -------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_index(const int64_t* __restrict__ hash, int64_t n, const FastMod fm, int64_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i < n) out[i] = (int64_t)fm_index((uint64_t)hash[i], fm);
}

// ---- popcount: set bits (Bloom) / non-zero bytes (counting)  (bloom/buffer/UnsafeByteBuffer.java:121-150) -------------
// Pure streaming read of the array: uint4 loads, warp shuffle reduction, one atomic per warp.
template <int BYTES_MODE>
__global__ void __launch_bounds__(kThreads) k_popcount(const uint4* __restrict__ v, int64_t n_vec, unsigned long long* out) {
    unsigned long long c = 0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * kThreads) {
        const uint4 x = __ldg(&v[i]);
        if (BYTES_MODE) c += (__popc(__vcmpne4(x.x, 0)) + __popc(__vcmpne4(x.y, 0)) + __popc(__vcmpne4(x.z, 0)) + __popc(__vcmpne4(x.w, 0))) >> 3;
        else c += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// ---- synthetic reads (bench / fixtures; a counter-based generator the CPU checker restates bit for bit) ----------------
__device__ __forceinline__ int synth_genome_base(uint64_t seed, uint64_t pos) { return (int)(mix64(seed ^ mix64(pos)) & 3); }
__global__ void __launch_bounds__(kThreads) k_synth_reads(uint64_t seed, uint64_t genome_len, uint64_t first_read, int64_t n_reads, int L,
                                                         uint32_t err_ppm, int64_t stride_bases, uint64_t* __restrict__ packed) {
    const int words_per_read = (int)(stride_bases >> 5);
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= n_reads * words_per_read) return;
    const int64_t rl = t / words_per_read;
    const int wi = (int)(t - rl * words_per_read);
    const uint64_t r = first_read + (uint64_t)rl;
    const uint64_t h = mix64(seed * 0x100000001B3ULL + 2 * r + 1);
    const uint64_t p0 = (h >> 1) % (genome_len - (uint64_t)L + 1);
    const int rc = (int)(h & 1);
    uint64_t w = 0;
    for (int j = 0; j < 32; ++j) {
        const int i = wi * 32 + j;
        if (i >= L) break;
        int b = rc ? 3 - synth_genome_base(seed, p0 + (uint64_t)(L - 1 - i)) : synth_genome_base(seed, p0 + (uint64_t)i);
        const uint64_t e = mix64((seed + 0x5851F42D4C957F2DULL) ^ mix64(r * 1024 + (uint64_t)i));
        if ((uint32_t)(e % 1000000u) < err_ppm) b = (b + 1 + (int)((e >> 40) % 3)) & 3;
        w |= (uint64_t)b << (2 * j);
    }
    packed[rl * words_per_read + wi] = w;
}

// Long reads (BASELINE configs[4]: ONT-like, ~2 kb, substitutions + insertions + deletions).  Integer-only, so that the CPU checker's
// twin (in oracle/rnabloom_oracle.c) produces the same bases: read r has length 500 + three draws of [0, 1000)
// (mean ~2 kb) and walks the virtual genome from a random start on a random strand; per emitted base one draw decides
// deletion (skip genome bases first) / insertion (random base, genome does not advance) / substitution / match.
__host__ __device__ __forceinline__ int synth_long_len(uint64_t seed, uint64_t r) {
    const uint64_t h = mix64(seed * 0x100000001B3ULL + 2 * r);
    return 500 + (int)(h % 1000u) + (int)((h >> 20) % 1000u) + (int)((h >> 40) % 1000u);
}
__global__ void __launch_bounds__(kThreads) k_synth_long_reads(uint64_t seed, uint64_t genome_len, uint64_t first_read, int64_t n_reads, uint32_t sub_ppm,
                                                              uint32_t ins_ppm, uint32_t del_ppm, const int64_t* __restrict__ read_off,
                                                              uint64_t* __restrict__ packed) {
    const int64_t rl = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (rl >= n_reads) return;
    const uint64_t r = first_read + (uint64_t)rl;
    const int L = synth_long_len(seed, r);
    const uint64_t h = mix64(seed * 0x100000001B3ULL + 2 * r + 1);
    const uint64_t span = (uint64_t)(2 * L + 64);                       // genome bases a read can consume at most (deletions)
    const uint64_t p0 = (h >> 1) % (genome_len - span + 1);
    const int rc = (int)(h & 1);
    uint64_t gpos = 0;                                                  // genome bases consumed so far
    uint64_t* out = packed + (read_off[rl] >> 5);
    uint64_t w = 0;
    for (int i = 0; i < L; ++i) {
        int b;
        for (int tries = 0;; ++tries) {
            const uint64_t e = mix64((seed + 0x5851F42D4C957F2DULL) ^ mix64(r * 8192 + (uint64_t)i * 4 + (uint64_t)min(tries, 3)));
            const uint32_t u = (uint32_t)(e % 1000000u);
            if (u < del_ppm && tries < 3 && gpos + 1 < span) { ++gpos; continue; }      // deletion: the genome base is skipped
            const int gb = rc ? 3 - synth_genome_base(seed, p0 + span - 1 - gpos) : synth_genome_base(seed, p0 + gpos);
            if (u < del_ppm + ins_ppm) { b = (int)((e >> 40) & 3); break; }             // insertion: the genome does not advance
            b = gb;
            if (u < del_ppm + ins_ppm + sub_ppm) b = (gb + 1 + (int)((e >> 40) % 3)) & 3;
            if (gpos + 1 < span) ++gpos;
            break;
        }
        w |= (uint64_t)b << (2 * (i & 31));
        if ((i & 31) == 31 || i == L - 1) { out[i >> 5] = w; w = 0; }
    }
}

// ---- FASTQ/FASTA front end: ASCII (+PHRED33) -> 2-bit codes + usable-base mask ------------------------------------------
// One thread per 32-base output word.  Replaces the host regex pre-pass (util/SeqUtils.java:1430-1438, RNABloom.java:567-577).
__global__ void __launch_bounds__(kThreads) k_pack_ascii(const char* __restrict__ bases, const char* __restrict__ quals,
                                                        const int64_t* __restrict__ ascii_off, const int64_t* __restrict__ word_off,
                                                        int64_t n_reads, int64_t n_words, int min_qual, uint64_t* __restrict__ packed,
                                                        uint32_t* __restrict__ mask, uint32_t* __restrict__ rcm, int uniform_len, int64_t uniform_a0) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= n_words) return;
    int64_t a0;
    int len, first;
    if (uniform_len > 0) {   // records of one length: no offset tables (ascii_off / word_off are not read)
        const int wpr = (uniform_len + 31) >> 5;
        const int64_t r = t / wpr;
        a0 = uniform_a0 + r * uniform_len;
        len = uniform_len;
        first = (int)(t - r * wpr) * 32;
    } else {
        int64_t lo = 0, hi = n_reads;  // read whose word range contains t
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(&word_off[mid]) <= t) lo = mid; else hi = mid;
        }
        a0 = __ldg(&ascii_off[lo]);
        len = (int)(__ldg(&ascii_off[lo + 1]) - a0);
        first = (int)(t - __ldg(&word_off[lo])) * 32;
    }
    uint64_t w = 0;
    uint32_t m = 0, rm = 0;
    const int qlo = '!' + min_qual;
    for (int j = 0; j < 32; ++j) {
        const int i = first + j;
        if (i >= len) { m |= 1u << j; continue; }
        const unsigned char c = (unsigned char)bases[a0 + i];
        int code = 0;
        bool ok = true;
        switch (c) {
            case 'A': case 'a': code = 0; break;
            case 'C': case 'c': code = 1; break;
            case 'G': case 'g': code = 2; break;
            case 'T': case 't': case 'U': case 'u': code = 3; break;
            default: {
                // not a nucleotide: 0 on the forward strand (msTab row c), but the reverse strand reads msTab row c & 0x07
                // (NTHash.java:100-101 rows 0..7 = N T N G A A N C): the code field carries the base whose FORWARD seed that row holds
                ok = false;
                const int row = c & 7;
                const int rc_code = row == 1 ? 3 : row == 3 ? 2 : (row == 4 || row == 5) ? 0 : row == 7 ? 1 : -1;
                if (rc_code >= 0) { code = rc_code; rm |= 1u << j; }
            }
        }
        if (quals) { const unsigned char q = (unsigned char)quals[a0 + i]; if (q < qlo || q > '~') ok = false; }
        w |= (uint64_t)code << (2 * j);
        if (!ok) m |= 1u << j;
    }
    packed[t] = w;
    mask[t] = m;
    if (rcm) rcm[t] = rm;
}

}  // namespace rb
