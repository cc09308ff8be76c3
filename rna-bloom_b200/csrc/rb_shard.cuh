// rb_shard.cuh -- phase kernels of the hash-sharded pipeline (filters split by index range over G GPUs).
//
// The logical filters stay exactly the reference's single arrays (graph/BloomFilterDeBruijnGraph.java:75-104): rank r owns indices
// [r*S, (r+1)*S) of each of them, so every *probe* (not every k-mer) is routed to the owner of its index.  Between the phases the
// fixed-capacity send regions are exchanged with an all-to-all over NVLink by the host orchestrator (rna-bloom_b200/sharded.py).
//
//   insert:  K1 route keys to their home rank (range of the mixed key)      -> all-to-all
//            K2 home: aggregate duplicates (key -> multiplicity)               (exact first-then-repeat semantics of graph.add :405-412)
//            K3 home: emit dbgbf test-and-set probes to the index owners     -> all-to-all
//            K4 owner: atomicOr, reply old bit                               -> all-to-all back
//            K5 home: present = AND(old bits); increments = m - 1 + present; emit cbf read probes -> all-to-all
//            K6 owner: reply counter bytes                                   -> all-to-all back
//            K7 home: replay `increments` min-increments on the h counters (CountingBloomFilter.java:170-194), emit raises -> all-to-all
//            K8 owner: counter = max(counter, value)
//   lookup:  K9 origin: emit dbgbf+cbf read probes -> all-to-all; K10 owner: reply bit/byte -> back; K11 origin: count (graph :562-570)
#pragma once
#include "rb_kernels.cuh"

namespace rb {

constexpr int kCntStride = 32;  // ints between two region counters: one counter per 128 B line (atomics to one line serialise in L2)
constexpr int kShardSub = 16;   // send regions per destination rank: spreads the cursor atomics (one hot counter per rank throttles at ~1.4 G/s)
struct ShardGeom {
    int n_ranks;
    int64_t cap;            // capacity (records) of each of the n_ranks * kShardSub regions of the send buffer
    uint64_t dbg_shard;     // bits per rank of the dbgbf (multiple of 1024)
    uint64_t cbf_shard;     // bytes per rank of the cbf (multiple of 4)
};

__device__ __forceinline__ int home_of(uint64_t key, int n_ranks) {
    const uint64_t m = (key * 0x9E3779B97F4A7C15ULL) >> 32;
    return (int)((m * (uint64_t)n_ranks) >> 32);
}

// Claims a slot of region `dest` (warp-aggregated) and returns its position in the send buffer, or -1 on overflow.
__device__ __forceinline__ int64_t region_push(int* __restrict__ cnt, int dest_rank, int64_t cap, int* __restrict__ overflow) {
    const int dest = dest_rank * kShardSub + (int)(blockIdx.x % kShardSub);
    const unsigned peers = __match_any_sync(__activemask(), dest);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(&cnt[dest * kCntStride], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const int p = base + __popc(peers & ((1u << lane) - 1u));
    if (p >= cap) { *overflow = 1; return -1; }
    return (int64_t)dest * cap + p;
}

// ---- K1: k-merise the reads, send every usable k-mer's base hash to its home rank ------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_route_keys(const Ingest g, int k, const ShardGeom sg, int64_t* __restrict__ send,
                                                        int* __restrict__ cnt, int* __restrict__ overflow) {
    __shared__ RollLut lut;
    build_lut(&lut, k);
    const int64_t pos = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (pos >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - pos);
    PositionWalker<MODE> pw;
    pw.start(g, pos, k, lut);
    for (int i = 0; i < n; ++i) {
        pw.advance(g, k, lut);
        if (pw.wk.bad == 0) {
            const uint64_t b = pw.wk.base();
            const int64_t p = region_push(cnt, home_of(b, sg.n_ranks), sg.cap, overflow);
            if (p >= 0) send[p] = (int64_t)b;
        }
    }
}

// ---- K2: home rank aggregates the received keys into (key, multiplicity) ---------------------------------------------
// table: T slots (power of two) of keys (0 = empty) + counts; slot T serves key 0.
struct AggTable {
    unsigned long long* keys;
    unsigned int* counts;
    uint64_t mask;
    int shift;
};
__global__ void __launch_bounds__(kThreads) k_agg_insert(const int64_t* __restrict__ recv, const int* __restrict__ recv_cnt, const ShardGeom sg,
                                                        const AggTable t) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int src = (int)(i / sg.cap);
    if (src >= sg.n_ranks * kShardSub || (i - (int64_t)src * sg.cap) >= recv_cnt[src * kCntStride]) return;
    const uint64_t key = (uint64_t)recv[i];
    if (key == 0) { atomicAdd(&t.counts[t.mask + 1], 1u); return; }
    uint64_t s = (key * 0x9E3779B97F4A7C15ULL) >> t.shift;
    for (;;) {
        const unsigned long long old = atomicCAS(&t.keys[s], 0ULL, (unsigned long long)key);
        if (old == 0ULL || old == key) { atomicAdd(&t.counts[s], 1u); return; }
        s = (s + 1) & t.mask;
    }
}

__device__ __forceinline__ bool agg_slot(const AggTable& t, int64_t s, uint64_t* key, unsigned int* m) {
    if (s > (int64_t)t.mask + 1) return false;
    *m = t.counts[s];
    if (*m == 0) return false;
    *key = (s == (int64_t)t.mask + 1) ? 0ULL : (uint64_t)t.keys[s];
    return true;
}

// ---- K3: emit the dbgbf test-and-set probes of every distinct key ------------------------------------------------------
// probe record = index local to the owner's shard; pos[slot*H+h] remembers where the reply will come back
template <int MAXH>
__global__ void __launch_bounds__(kThreads) k_emit_dbg(const AggTable t, const HashMults hm, const FastMod fm, int num_hash, const ShardGeom sg,
                                                      int64_t* __restrict__ send, int* __restrict__ cnt, int* __restrict__ pos,
                                                      int* __restrict__ overflow) {
    const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    uint64_t key; unsigned int m;
    const bool live = agg_slot(t, s, &key, &m);
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        if (h < num_hash) {
            // region_push is warp-collective over the active lanes with equal dest: keep the branch structure uniform per h
            if (live) {
                const uint64_t gidx = fm_index(expand_hash(key, h, hm), fm);
                const int dest = (int)(gidx / sg.dbg_shard);
                const int64_t p = region_push(cnt, dest, sg.cap, overflow);
                if (p >= 0) send[p] = (int64_t)(gidx - (uint64_t)dest * sg.dbg_shard);
                pos[s * num_hash + h] = (int)p;
            }
        }
    }
}

// ---- K4: owner applies test-and-set, replies the old bit ----------------------------------------------------------------
template <int SET>   // SET = 0: read only (addCountIfPresent never sets bits)
__global__ void __launch_bounds__(kThreads) k_apply_dbg(const int64_t* __restrict__ recv, const int* __restrict__ recv_cnt, const ShardGeom sg,
                                                       uint32_t* __restrict__ words, uint8_t* __restrict__ reply) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int src = (int)(i / sg.cap);
    if (src >= sg.n_ranks * kShardSub || (i - (int64_t)src * sg.cap) >= recv_cnt[src * kCntStride]) return;
    const uint64_t idx = (uint64_t)recv[i];
    const uint32_t bit = 1u << (idx & 31);
    uint32_t w = ld_cg(&words[idx >> 5]);
    if (SET && !(w & bit)) w = atomicOr(&words[idx >> 5], bit);
    reply[i] = (w & bit) ? 1 : 0;
}

// ---- K5: present = AND(old bits); increments = m - 1 + present; emit counter read probes ---------------------------------
template <int MAXH>
__global__ void __launch_bounds__(kThreads) k_combine_dbg_emit_cbf(const AggTable t, const HashMults hm, const FastMod cbf_fm, int hd, int hc,
                                                                  const ShardGeom sg, const int* __restrict__ pos_dbg,
                                                                  const uint8_t* __restrict__ reply_dbg, int policy,
                                                                  unsigned int* __restrict__ inc, int64_t* __restrict__ send,
                                                                  int* __restrict__ cnt, int* __restrict__ pos_cbf, int* __restrict__ overflow) {
    const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    uint64_t key; unsigned int m;
    bool live = agg_slot(t, s, &key, &m);
    unsigned int n_inc = 0;
    if (live) {
        bool present = true;
        for (int h = 0; h < hd; ++h) { const int p = pos_dbg[s * hd + h]; present = present && (p >= 0) && reply_dbg[p]; }
        // graph.add (:405-412): the first sighting of an absent k-mer only sets bits.  addCountIfPresent (:424-428): needs presence.
        n_inc = (policy == POLICY_COUNT_IF_PRESENT) ? (present ? m : 0u) : (m - 1u + (present ? 1u : 0u));
        inc[s] = n_inc;
    }
    live = live && n_inc > 0;
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        if (h < hc) {
            if (live) {
                const uint64_t gidx = fm_index(expand_hash(key, h, hm), cbf_fm);
                const int dest = (int)(gidx / sg.cbf_shard);
                const int64_t p = region_push(cnt, dest, sg.cap, overflow);
                if (p >= 0) send[p] = (int64_t)(gidx - (uint64_t)dest * sg.cbf_shard);
                pos_cbf[s * hc + h] = (int)p;
            }
        }
    }
}

// ---- K6: owner replies counter bytes -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_apply_cbf_read(const int64_t* __restrict__ recv, const int* __restrict__ recv_cnt, const ShardGeom sg,
                                                            const uint32_t* __restrict__ words, uint8_t* __restrict__ reply) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int src = (int)(i / sg.cap);
    if (src >= sg.n_ranks * kShardSub || (i - (int64_t)src * sg.cap) >= recv_cnt[src * kCntStride]) return;
    const uint64_t idx = (uint64_t)recv[i];
    reply[i] = (uint8_t)(ld_cg(&words[idx >> 2]) >> ((idx & 3) * 8));
}

// ---- K7: replay the increments on the h counter values, emit raises ---------------------------------------------------------------
// CountingBloomFilter.increment (:170-194) n times: each time every counter equal to the minimum becomes MiniFloat.increment(min).
// record = local index | value << 56
template <int MAXH>
__global__ void __launch_bounds__(kThreads) k_combine_cbf_emit_raise(const AggTable t, const HashMults hm, const FastMod cbf_fm, int hc,
                                                                    const ShardGeom sg, const unsigned int* __restrict__ inc,
                                                                    const int* __restrict__ pos_cbf, const uint8_t* __restrict__ reply_cbf,
                                                                    int policy, uint64_t rng_seed, int64_t* __restrict__ send,
                                                                    int* __restrict__ cnt, int* __restrict__ overflow) {
    const int64_t s = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    uint64_t key; unsigned int m;
    bool live = agg_slot(t, s, &key, &m);
    unsigned int n_inc = live ? inc[s] : 0u;
    live = live && n_inc > 0;
    int v0[MAXH], v[MAXH];
    if (live) {
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
            if (h < hc) { const int p = pos_cbf[s * hc + h]; v0[h] = p >= 0 ? (int)(reply_cbf[p] & 0x7Fu) : 127; } else v0[h] = 127;
            v[h] = v0[h];
        }
        // slots that map to the same counter must move together
        uint64_t gi[MAXH];
#pragma unroll
        for (int h = 0; h < MAXH; ++h) gi[h] = h < hc ? fm_index(expand_hash(key, h, hm), cbf_fm) : ~0ULL;
        if (policy == POLICY_COUNT_IF_PRESENT) {   // "&& cbf.getCount(hashVals) > 0" (graph :425) on the state before this batch
            int mn = 127;
#pragma unroll
            for (int h = 0; h < MAXH; ++h) if (h < hc) mn = min(mn, v[h]);
            if (mn == 0) n_inc = 0;
        }
        uint64_t r = mix64(key ^ rng_seed);
        for (unsigned int it = 0; it < n_inc; ++it) {
            int mn = 127;
#pragma unroll
            for (int h = 0; h < MAXH; ++h) if (h < hc) mn = min(mn, v[h]);
            if (mn >= 127) break;
            r = mix64(r + it);
            const int u = minifloat_increment(mn, r);
            if (u != mn) {
#pragma unroll
                for (int h = 0; h < MAXH; ++h) if (h < hc && v[h] == mn) v[h] = u;
            }
        }
#pragma unroll
        for (int h = 0; h < MAXH; ++h)   // one raise per distinct counter
            for (int h2 = 0; h2 < h; ++h2) if (h < hc && gi[h] == gi[h2]) v[h] = v0[h];
    }
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
        if (h < hc) {
            const bool emit = live && v[h] > v0[h];
            if (emit) {
                const uint64_t gidx = fm_index(expand_hash(key, h, hm), cbf_fm);
                const int dest = (int)(gidx / sg.cbf_shard);
                const int64_t p = region_push(cnt, dest, sg.cap, overflow);
                if (p >= 0) send[p] = (int64_t)((gidx - (uint64_t)dest * sg.cbf_shard) | ((uint64_t)v[h] << 56));
            }
        }
    }
}

// ---- K8: owner raises counters --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_apply_cbf_raise(const int64_t* __restrict__ recv, const int* __restrict__ recv_cnt, const ShardGeom sg,
                                                             uint32_t* __restrict__ words) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int src = (int)(i / sg.cap);
    if (src >= sg.n_ranks * kShardSub || (i - (int64_t)src * sg.cap) >= recv_cnt[src * kCntStride]) return;
    const uint64_t rec = (uint64_t)recv[i];
    const uint64_t idx = rec & 0x00FFFFFFFFFFFFFFULL;
    uint32_t* wp = &words[idx >> 2];
    byte_raise(wp, (int)(idx & 3) * 8, (uint32_t)(rec >> 56), ld_cg(wp));
}

// ---- K9: lookup probes (dbgbf bits + cbf bytes) of every usable k-mer instance -------------------------------------------------------
// record = local index | (1<<63 for cbf probes); pos[(instance*(hd+hc)) + j]
template <int MODE, int MAXH>
__global__ void __launch_bounds__(kThreads) k_route_lookup(const Ingest g, int k, const HashMults hm, const FastMod dbg_fm, const FastMod cbf_fm,
                                                          int hd, int hc, const ShardGeom sg, int64_t* __restrict__ send, int* __restrict__ cnt,
                                                          int* __restrict__ pos, int64_t* __restrict__ fhash, int64_t* __restrict__ rhash,
                                                          int* __restrict__ overflow) {
    __shared__ RollLut lut;
    build_lut(&lut, k);
    const int64_t p0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * kChunk;
    if (p0 >= g.n_pos) return;
    const int n = (int)min((int64_t)kChunk, g.n_pos - p0);
    PositionWalker<MODE> pw;
    pw.start(g, p0, k, lut);
    const int H = hd + hc;
    for (int i = 0; i < n; ++i) {
        pw.advance(g, k, lut);
        const int64_t inst = p0 + i;
        if (fhash) fhash[inst] = (int64_t)pw.wk.f;
        if (rhash) rhash[inst] = (int64_t)pw.wk.r;
        const bool ok = pw.wk.bad == 0;
        const uint64_t b = pw.wk.base();
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
            if (h < hd) {
                if (ok) {
                    const uint64_t gidx = fm_index(expand_hash(b, h, hm), dbg_fm);
                    const int dest = (int)(gidx / sg.dbg_shard);
                    const int64_t p = region_push(cnt, dest, sg.cap, overflow);
                    if (p >= 0) send[p] = (int64_t)(gidx - (uint64_t)dest * sg.dbg_shard);
                    pos[inst * H + h] = (int)p;
                } else pos[inst * H + h] = -2;
            }
        }
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
            if (h < hc) {
                if (ok) {
                    const uint64_t gidx = fm_index(expand_hash(b, h, hm), cbf_fm);
                    const int dest = (int)(gidx / sg.cbf_shard);
                    const int64_t p = region_push(cnt, dest, sg.cap, overflow);
                    if (p >= 0) send[p] = (int64_t)((gidx - (uint64_t)dest * sg.cbf_shard) | (1ULL << 63));
                    pos[inst * H + hd + h] = (int)p;
                } else pos[inst * H + hd + h] = -2;
            }
        }
    }
}

// ---- K10: owner answers lookup probes ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_apply_lookup(const int64_t* __restrict__ recv, const int* __restrict__ recv_cnt, const ShardGeom sg,
                                                          const uint32_t* __restrict__ dbg_words, const uint32_t* __restrict__ cbf_words,
                                                          uint8_t* __restrict__ reply) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int src = (int)(i / sg.cap);
    if (src >= sg.n_ranks * kShardSub || (i - (int64_t)src * sg.cap) >= recv_cnt[src * kCntStride]) return;
    const uint64_t rec = (uint64_t)recv[i];
    const uint64_t idx = rec & 0x7FFFFFFFFFFFFFFFULL;
    if (rec >> 63) reply[i] = (uint8_t)(ld_cg(&cbf_words[idx >> 2]) >> ((idx & 3) * 8));
    else reply[i] = (uint8_t)((ld_cg(&dbg_words[idx >> 5]) >> (idx & 31)) & 1u);
}

// ---- K11: origin combines the replies into counts (graph :562-570) -----------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_combine_lookup(int64_t n_inst, int hd, int hc, const int* __restrict__ pos, const uint8_t* __restrict__ reply,
                                                            float* __restrict__ counts) {
    const int64_t inst = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (inst >= n_inst) return;
    const int H = hd + hc;
    float c = 0.f;
    if (pos[inst * H] != -2) {
        bool all = true;
        for (int h = 0; h < hd; ++h) { const int p = pos[inst * H + h]; all = all && p >= 0 && reply[p]; }
        if (all) {
            int mn = 127;
            for (int h = 0; h < hc; ++h) { const int p = pos[inst * H + hd + h]; const int v = p >= 0 ? (int)(int8_t)reply[p] : 0; mn = v < mn ? v : mn; }
            c = minifloat_to_float(mn) + 1.f;
        }
    }
    counts[inst] = c;
}

}  // namespace rb
