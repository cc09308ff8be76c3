// rb_device.cuh -- device-side building blocks of the RNA-Bloom hot path on sm_100a.
//
// Everything here is integer / bitwise work bounded by HBM random-sector traffic (no tensor cores).
// Citations are relative to /root/reference/src/rnabloom/.
#pragma once
#include <stdint.h>
#ifdef RB_EMU
// host emulation of the kernels for logic tests without a GPU (tests/emu/cuda_emu.h; never part of the product library)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
// RB_LAUNCH(grid, block, dynamic smem bytes, stream, kernel)(args...);  the kernel goes last so template commas survive the preprocessor
#define RB_LAUNCH(grid, block, smem, stream, ...) __VA_ARGS__<<<(grid), (block), (smem), (stream)>>>
#define RB_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#endif

namespace rb {

// ---- a1 seeds (bloom/hash/NTHash.java:39-43) and the multi-hash constants (:33-36) -----------------------------
// 2-bit code order A0 C1 G2 T3; complement of code c is 3-c.
__host__ __device__ __forceinline__ constexpr uint64_t seed_of_code(int c) {
    return c == 0 ? 0x3c8bfbb395c60474ULL : c == 1 ? 0x3193c18562a02b4cULL : c == 2 ? 0x20323ed082572324ULL : 0x295549f54be24456ULL;
}
constexpr uint64_t kMultiSeed = 0x90b45d39fb6da1faULL;
constexpr int kMultiShift = 27;
constexpr int kMaxHash = 8;     // largest numHash the kernels are instantiated for
constexpr int kChunk = 16;      // k-mer positions per thread
constexpr int kGroup = 4;       // k-mers whose probes are in flight together per thread

__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t v, int s) { s &= 63; return (v << s) | (v >> ((64 - s) & 63)); }
__host__ __device__ __forceinline__ uint64_t rotr64(uint64_t v, int s) { s &= 63; return (v >> s) | (v << ((64 - s) & 63)); }
__device__ __forceinline__ uint64_t rotl1(uint64_t v) { return (v << 1) | (v >> 63); }
__device__ __forceinline__ uint64_t rotr1(uint64_t v) { return (v >> 1) | (v << 63); }

// ---- a8 index: (hash >>> 1) % size for an arbitrary 63-bit size (bloom/BloomFilter.java:108-111) ------------------
// pow2 sizes: AND.  otherwise Barrett: q = mulhi(n, floor(2^64/size)) is the true quotient or one less (n < 2^63).
struct FastMod {
    uint64_t size;
    uint64_t magic;  // floor(2^64 / size), only used when !pow2
    uint64_t mask;   // size-1 when pow2
    int pow2;
};
__device__ __forceinline__ uint64_t fm_index(uint64_t hash, const FastMod& fm) {
    const uint64_t n = hash >> 1;
    if (fm.pow2) return n & fm.mask;
    const uint64_t q = __umul64hi(n, fm.magic);
    uint64_t r = n - q * fm.size;
    if (r >= fm.size) r -= fm.size;
    return r;
}

// ---- a4 multi-hash expansion (bloom/hash/NTHash.java:518-527): mult[i] = i ^ (k * multiSeed) ----------------------
struct HashMults { uint64_t m[kMaxHash]; };
__device__ __forceinline__ uint64_t expand_hash(uint64_t base, int i, const HashMults& hm) {
    if (i == 0) return base;
    uint64_t t = base * hm.m[i];
    return t ^ (t >> kMultiShift);
}

// ---- a5 combineHashValues (bloom/hash/HashFunction.java:260-266); the int literal sign-extends -----------------
__host__ __device__ __forceinline__ uint64_t combine_hash(uint64_t a, uint64_t b) {
    return a ^ (b + 0xFFFFFFFF9E3779B9ULL + (a << 6) + (b >> 2));
}

// ---- a12 MiniFloat (util/MiniFloat.java:27-45) -------------------------------------------------------------------
__host__ __device__ __forceinline__ float minifloat_to_float(int b) {  // b is the signed byte
    if (b <= 7) return (float)b;
    return (float)((b & 7) | 8) * (float)(1u << ((b >> 3) - 1));      // exponent <= 14: exact in fp32
}
// rnd: 64 random bits standing in for Math.random(); Java draws (int)(u * Integer.MAX_VALUE) % (1 << e) == 0
__device__ __forceinline__ int minifloat_increment(int b, uint64_t rnd) {
    if (b <= 7) return (int)(int8_t)(b + 1);
    if (b < 127) {
        const int e = (b >> 3) - 1;
        const uint32_t r31 = (uint32_t)(((rnd >> 33) * 2147483647ULL) >> 31);
        if ((r31 & ((1u << e) - 1u)) == 0) return b + 1;
    }
    return b;
}
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser: counter-based RNG and table hashing
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// ---- HBM access: filters are written by other SMs during a kernel, so loads go to L2 (ld.global.cg) ---------------
__device__ __forceinline__ uint32_t ld_cg(const uint32_t* p) { return __ldcg(p); }
// The sliced engine keeps one filter slice resident in L2 while records and answers stream past it: accesses to the slice carry an
// L2 evict_last policy, the streams are read / written with .cs (evict first).  Without the hints the streams push slice lines out
// (ncu: 49 GB of DRAM reads per look-up round where records + one sweep of the filters are 23 GB).
#ifdef RB_EMU
struct L2Keep { };
__device__ __forceinline__ L2Keep l2_keep_policy() { return L2Keep(); }
__device__ __forceinline__ uint32_t ld_cg_keep(const uint32_t* p, L2Keep) { return __ldcg(p); }
__device__ __forceinline__ uint32_t atomic_or_keep(uint32_t* p, uint32_t v, L2Keep) { return atomicOr(p, v); }
__device__ __forceinline__ uint32_t atomic_cas_keep(uint32_t* p, uint32_t cmp, uint32_t v, L2Keep) { return atomicCAS(p, cmp, v); }
#else
struct L2Keep { uint64_t pol; };
__device__ __forceinline__ L2Keep l2_keep_policy() {
    L2Keep k;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(k.pol));
    return k;
}
__device__ __forceinline__ uint32_t ld_cg_keep(const uint32_t* p, L2Keep k) {
    uint32_t v;
    asm volatile("ld.global.cg.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(k.pol) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t atomic_or_keep(uint32_t* p, uint32_t v, L2Keep k) {
    uint32_t old;
    asm volatile("atom.global.or.L2::cache_hint.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(p), "r"(v), "l"(k.pol) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t atomic_cas_keep(uint32_t* p, uint32_t cmp, uint32_t v, L2Keep) { return atomicCAS(p, cmp, v); }   // PTX: atom.cas takes no cache hint
#endif

// ---- filter views ---------------------------------------------------------------------------------------------------
struct BitFilter {    // BloomFilter over UnsafeBitBuffer: bit i = byte i/8, mask 1<<(i%8) == bit (i&31) of LE word i>>5
    uint32_t* words;
    FastMod fm;
    int num_hash;
};
struct ByteFilter {   // CountingBloomFilter over UnsafeByteBuffer: slot i = byte (i&3) of LE word i>>2
    uint32_t* words;
    FastMod fm;
    int num_hash;
};

// byte RMW inside a 32-bit word.  Counter bytes are 0..127 (MiniFloat saturates at Byte.MAX_VALUE, util/MiniFloat.java:31-38),
// which leaves bit 7 free: during an insert kernel it is the per-slot lock of the increment protocol below and is always clear
// again when the kernel ends.
constexpr uint32_t kLockBit = 0x80u;
__device__ __forceinline__ int byte_of(uint32_t w, int sh) { return (int)(int8_t)(w >> sh); }
// if the byte (all 8 bits) equals `expect`: byte = nv, return true.  `old` is the caller's latest view of the word.
__device__ __forceinline__ bool byte_cas(uint32_t* wp, int sh, uint32_t expect, uint32_t nv, uint32_t old) {
    for (;;) {
        if (((old >> sh) & 0xFFu) != expect) return false;
        const uint32_t nw = (old & ~(0xFFu << sh)) | (nv << sh);
        const uint32_t got = atomicCAS(wp, old, nw);
        if (got == old) return true;
        old = got;
    }
}
// low 7 bits of the byte = max(low 7 bits, v); a lock bit held by somebody else is preserved.  Values only grow, so a
// stale `old` that already shows >= v proves there is nothing to do.
__device__ __forceinline__ void byte_raise(uint32_t* wp, int sh, uint32_t v, uint32_t old) {
    for (;;) {
        if (((old >> sh) & 0x7Fu) >= v) return;
        const uint32_t nw = (old & ~(0x7Fu << sh)) | (v << sh);
        const uint32_t got = atomicCAS(wp, old, nw);
        if (got == old) return;
        old = got;
    }
}

__device__ __forceinline__ void byte_raise_keep(uint32_t* wp, int sh, uint32_t v, uint32_t old, L2Keep k) {
    for (;;) {
        if (((old >> sh) & 0x7Fu) >= v) return;
        const uint32_t nw = (old & ~(0x7Fu << sh)) | (v << sh);
        const uint32_t got = atomic_cas_keep(wp, old, nw, k);
        if (got == old) return;
        old = got;
    }
}

// ---- claim table: elects exactly one "first" instance per distinct base hash between two clears ------------------
// BloomFilter.lookupThenAdd (bloom/BloomFilter.java:147-155) is test-and-set of h bits; two concurrent instances of
// one k-mer could each win a different bit and both report "absent".  Instances that see a clear bit therefore claim
// the k-mer here first: the winner sets the bits and reports absent, every later instance reports present -- exactly the
// outcome of the sequential reference for duplicates (DESIGN.md "Linearisation").
struct ClaimTable {
    unsigned long long* slots;  // capacity + 1 entries; slot[capacity] serves key 0
    uint64_t mask;              // capacity - 1
    int shift;                  // 64 - log2(capacity)
};
__device__ __forceinline__ bool claim_first(const ClaimTable& t, uint64_t key) {
    if (key == 0) return atomicExch(&t.slots[t.mask + 1], 1ULL) == 0ULL;
    uint64_t s = (key * 0x9E3779B97F4A7C15ULL) >> t.shift;
    for (;;) {
        const unsigned long long old = atomicCAS(&t.slots[s], 0ULL, (unsigned long long)key);
        if (old == 0ULL) return true;
        if (old == key) return false;
        s = (s + 1) & t.mask;
    }
}

// continue a claim whose first probe hit another key
__device__ __forceinline__ bool claim_first_from(const ClaimTable& t, uint64_t key, uint64_t s) {
    for (;;) {
        const unsigned long long old = atomicCAS(&t.slots[s], 0ULL, (unsigned long long)key);
        if (old == 0ULL) return true;
        if (old == key) return false;
        s = (s + 1) & t.mask;
    }
}

// ---- a11 CountingBloomFilter.increment (bloom/CountingBloomFilter.java:170-194), linearisable ---------------------
// Reference: min over the h slots, u = MiniFloat.increment(min), every slot that equals min becomes u.
// Concurrent version (DESIGN.md "Linearisation"): the LAST slot that holds the minimum is the k-mer's designated slot.
//   1. read the h slots; if one of them is locked, somebody (maybe a duplicate of this very k-mer) is mid-increment: retry
//   2. lock + bump the designated slot in one CAS  (min, unlocked) -> (u, locked)          <- linearisation point
//   3. raise every other slot that showed the minimum to u (max semantics, commutes with other k-mers' raises)
//   4. clear the lock bit (fire-and-forget atomicAnd)
// A thread holds at most one lock and never waits while holding it, so the protocol cannot deadlock; duplicates of one
// k-mer serialise on the designated slot (each adds exactly one increment), and k-mers that merely share a counter see
// each other's updates as if they had run one after the other.
constexpr int kLockSpinLimit = 1 << 16;  // only a corrupt (>127) uploaded byte could keep a slot "locked" for ever
template <int MAXH>
__device__ __forceinline__ int cbf_increment(const ByteFilter& cbf, uint64_t base, const HashMults& hm, uint64_t rng_key,
                                             const uint32_t* preloaded /* MAXH words or nullptr */) {
    uint64_t idx[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) idx[h] = (h < cbf.num_hash) ? fm_index(expand_hash(base, h, hm), cbf.fm) : 0;
    for (int attempt = 0;; ++attempt) {
        uint32_t w[MAXH];
        uint32_t any_lock = 0;
#pragma unroll
        for (int h = 0; h < MAXH; ++h)
            if (h < cbf.num_hash) w[h] = (attempt == 0 && preloaded) ? preloaded[h] : ld_cg(&cbf.words[idx[h] >> 2]);
        int v[MAXH];
        int mn = 127;
#pragma unroll
        for (int h = 0; h < MAXH; ++h) {
            if (h < cbf.num_hash) {
                const uint32_t b = (w[h] >> ((int)(idx[h] & 3) * 8)) & 0xFFu;
                any_lock |= b & kLockBit;
                v[h] = (int)(b & 0x7Fu);
                mn = v[h] < mn ? v[h] : mn;
            } else v[h] = 127;
        }
        if (any_lock && attempt < kLockSpinLimit) { __nanosleep(40); continue; }
        const int u = minifloat_increment(mn, mix64(rng_key + (uint64_t)attempt));
        if (u == mn) return u;
        int D = 0;
#pragma unroll
        for (int h = 0; h < MAXH; ++h) if (h < cbf.num_hash && v[h] == mn) D = h;
        uint64_t idxD = 0;
        uint32_t wD = 0;
#pragma unroll
        for (int h = 0; h < MAXH; ++h) if (h == D) { idxD = idx[h]; wD = w[h]; }
        const int shD = (int)(idxD & 3) * 8;
        if (!byte_cas(&cbf.words[idxD >> 2], shD, (uint32_t)mn, (uint32_t)u | kLockBit, wD)) continue;
#pragma unroll
        for (int h = 0; h < MAXH; ++h)
            if (h < cbf.num_hash && h != D && v[h] == mn && idx[h] != idxD)
                byte_raise(&cbf.words[idx[h] >> 2], (int)(idx[h] & 3) * 8, (uint32_t)u, w[h]);
        atomicAnd(&cbf.words[idxD >> 2], ~(kLockBit << shD));
        return u;
    }
}

// a11 getCount (bloom/CountingBloomFilter.java:235-251): MiniFloat of the minimum slot.  IGNORE_LOCK: called from a kernel that also
// increments (slots may carry the transient lock bit of another thread's increment): compare the 7 value bits only.
template <int MAXH, bool IGNORE_LOCK = false>
__device__ __forceinline__ int cbf_min(const ByteFilter& cbf, uint64_t base, const HashMults& hm) {
    uint32_t w[MAXH];
    uint64_t idx[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < cbf.num_hash) { idx[h] = fm_index(expand_hash(base, h, hm), cbf.fm); w[h] = ld_cg(&cbf.words[idx[h] >> 2]); }
    int mn = 127;
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < cbf.num_hash) {
            int v = byte_of(w[h], (int)(idx[h] & 3) * 8);
            if (IGNORE_LOCK) v &= 0x7F;
            mn = v < mn ? v : mn;
        }
    return mn;
}

// a9 BloomFilter.lookup (bloom/BloomFilter.java:170-178): all h probes issued together, then ANDed
template <int MAXH>
__device__ __forceinline__ bool bf_lookup(const BitFilter& bf, uint64_t base, const HashMults& hm) {
    uint32_t w[MAXH];
    uint64_t idx[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < bf.num_hash) { idx[h] = fm_index(expand_hash(base, h, hm), bf.fm); w[h] = ld_cg(&bf.words[idx[h] >> 5]); }
    bool all = true;
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < bf.num_hash) all = all && ((w[h] >> (idx[h] & 31)) & 1u);
    return all;
}
// a9 BloomFilter.add (bloom/BloomFilter.java:133-137): test before set keeps clean sectors clean
template <int MAXH>
__device__ __forceinline__ void bf_add(const BitFilter& bf, uint64_t base, const HashMults& hm) {
    uint32_t w[MAXH];
    uint64_t idx[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < bf.num_hash) { idx[h] = fm_index(expand_hash(base, h, hm), bf.fm); w[h] = ld_cg(&bf.words[idx[h] >> 5]); }
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < bf.num_hash && !((w[h] >> (idx[h] & 31)) & 1u)) atomicOr(&bf.words[idx[h] >> 5], 1u << (idx[h] & 31));
}
// a9 BloomFilter.lookupThenAdd (bloom/BloomFilter.java:147-155), linearised through the claim table
template <int MAXH>
__device__ __forceinline__ bool bf_lookup_then_add(const BitFilter& bf, const ClaimTable& ct, uint64_t base, const HashMults& hm) {
    uint32_t w[MAXH];
    uint64_t idx[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < bf.num_hash) { idx[h] = fm_index(expand_hash(base, h, hm), bf.fm); w[h] = ld_cg(&bf.words[idx[h] >> 5]); }
    bool all = true;
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < bf.num_hash) all = all && ((w[h] >> (idx[h] & 31)) & 1u);
    if (all) return true;
    if (!claim_first(ct, base)) return true;   // a duplicate already owns the first sighting
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
        if (h < bf.num_hash && !((w[h] >> (idx[h] & 31)) & 1u)) atomicOr(&bf.words[idx[h] >> 5], 1u << (idx[h] & 31));
    return false;
}

// ---- read ingest -----------------------------------------------------------------------------------------------------
struct Ingest {
    const uint64_t* packed;   // 2-bit codes, base b at bits 2*(b&31) of word b>>5
    const uint32_t* mask;     // nullable; bit (b&31) of word b>>5 set = unusable base
    const uint32_t* rcm;      // nullable; set only together with mask: the unusable base still contributes to the REVERSE-strand hash
                              // with the seed of its 2-bit code (NTHash.java:367-373 indexes msTab with c & 0x07, so e.g. 'Y' = 0x59
                              // hashes like the complement of 'A' on the reverse strand and as 0 on the forward strand)
    const int64_t* read_off;  // nullable (uniform layout)
    const int32_t* read_len;  // nullable (uniform layout)
    const int64_t* pos_off;   // exclusive prefix of per-read position counts (n_reads+1), nullable for uniform
    int64_t n_reads;
    int64_t n_pos;            // total positions (k-mers, or pairs) in this launch
    int64_t out_base;         // global output index of position 0 of this launch
    int64_t uniform_stride;   // bases between read starts in the uniform layout
    int32_t uniform_len;
    int32_t uniform_npos;     // positions per read in the uniform layout
    int64_t first_base;       // uniform layout: base offset of read 0 of this launch
    int64_t pos_bias;         // pos_off value of read 0 of this launch (pos_off holds absolute prefix values)
    int64_t read_base;        // index of read 0 of this launch in the caller's per-read outputs
};

// per-CTA lookup tables for the rolling update (a3: NTHash.java:584-586, 627-629, 491-495)
struct RollLut {
    uint64_t s[4];      // S[c]                      in-term of the forward strand
    uint64_t sk[4];     // rotl(S[c], k)              out-term of the forward strand
    uint64_t c[4];      // S[3-c]                     complement seed
    uint64_t cr1[4];    // rotr(S[3-c], 1)            out-term of the reverse strand
    uint64_t ck1[4];    // rotl(S[3-c], k-1)          in-term of the reverse strand
    uint64_t sr1[4];    // rotr(S[c], 1)              out-term of the reverse strand for an unusable base with a reverse seed (Ingest::rcm)
    uint64_t sk1[4];    // rotl(S[c], k-1)            in-term of the same
};
__device__ __forceinline__ void build_lut(RollLut* lut, int k) {
    if (threadIdx.x < 4) {
        const int c = threadIdx.x;
        const uint64_t s = seed_of_code(c), sc = seed_of_code(3 - c);
        lut->s[c] = s; lut->sk[c] = rotl64(s, k); lut->c[c] = sc; lut->cr1[c] = rotr64(sc, 1); lut->ck1[c] = rotl64(sc, k - 1);
        lut->sr1[c] = rotr64(s, 1); lut->sk1[c] = rotl64(s, k - 1);
    }
    __syncthreads();
}

// Sequential reader over the 2-bit stream (+ mask) that keeps the current words in registers.
// Words are fetched lazily, so nothing beyond the word holding the last requested base is touched.
struct BaseCursor {
    const uint64_t* packed;
    const uint32_t* mask;
    const uint32_t* rcm;
    int64_t b;        // absolute index of the next base
    int64_t widx;     // index of the cached words
    uint64_t w;
    uint32_t m, rm;
    __device__ __forceinline__ void seek(const uint64_t* p, const uint32_t* mk, int64_t base, const uint32_t* rc = nullptr) {
        packed = p; mask = mk; rcm = rc; b = base; widx = -1; w = 0; m = 0; rm = 0;
    }
    // returns code | (masked << 2) | (masked base with a reverse-strand seed << 3) and advances
    __device__ __forceinline__ int next() {
        const int64_t wi = b >> 5;
        if (wi != widx) {
            widx = wi;
            w = __ldg(&packed[wi]);
            m = mask ? __ldg(&mask[wi]) : 0u;
            rm = (mask && rcm) ? __ldg(&rcm[wi]) : 0u;
        }
        const int sh = (int)(b & 31);
        ++b;
        return (int)((w >> (2 * sh)) & 3) | (int)(((m >> sh) & 1u) << 2) | (int)(((rm >> sh) & 1u) << 3);
    }
};

// Rolling ntHash over one read: forward and/or reverse strand, plus the number of masked bases in the window.
template <int MODE>  // 0 fwd, 1 rc, 2 canonical (both)
struct KmerWalker {
    BaseCursor out, in;
    uint64_t f, r;
    int bad;
    // position the window on bases [start, start+k)
    __device__ __forceinline__ void init(const Ingest& g, int64_t start, int k, const RollLut& lut) {
        out.seek(g.packed, g.mask, start, g.rcm);
        in.seek(g.packed, g.mask, start, g.rcm);
        f = 0; r = 0; bad = 0;
        for (int j = 0; j < k; ++j) {
            const int c = in.next();
            const bool ok = !(c & 4);
            bad += (c >> 2) & 1;
            if (MODE != 1) f = rotl1(f) ^ (ok ? lut.s[c & 3] : 0ULL);
            if (MODE != 0) r ^= ok ? rotl64(lut.c[c & 3], j) : ((c & 8) ? rotl64(lut.s[c & 3], j) : 0ULL);
        }
    }
    __device__ __forceinline__ void roll(const RollLut& lut) {
        const int co = out.next(), ci = in.next();
        const bool oko = !(co & 4), oki = !(ci & 4);
        bad += ((ci >> 2) & 1) - ((co >> 2) & 1);
        if (MODE != 1) f = rotl1(f) ^ (oko ? lut.sk[co & 3] : 0ULL) ^ (oki ? lut.s[ci & 3] : 0ULL);
        if (MODE != 0) r = rotr1(r) ^ (oko ? lut.cr1[co & 3] : ((co & 8) ? lut.sr1[co & 3] : 0ULL)) ^ (oki ? lut.ck1[ci & 3] : ((ci & 8) ? lut.sk1[ci & 3] : 0ULL));
    }
    // hVals[0]: NTHashIterator / ReverseComplementNTHashIterator / CanonicalNTHashIterator (signed min, NTHash.java:494)
    __device__ __forceinline__ uint64_t base() const {
        if (MODE == 0) return f;
        if (MODE == 1) return r;
        return ((int64_t)r < (int64_t)f) ? r : f;
    }
};

// Locate position `pos` (launch-local) -> (read, offset inside read).  Uniform layout: one division.
__device__ __forceinline__ void locate(const Ingest& g, int64_t pos, int64_t& read, int32_t& in_read) {
    if (!g.pos_off) {
        read = pos / g.uniform_npos;
        in_read = (int32_t)(pos - read * g.uniform_npos);
        return;
    }
    int64_t lo = 0, hi = g.n_reads;  // last read with pos_off[read] <= pos
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(&g.pos_off[mid]) - g.pos_bias <= pos) lo = mid; else hi = mid;
    }
    read = lo;
    in_read = (int32_t)(pos - (__ldg(&g.pos_off[lo]) - g.pos_bias));
}
__device__ __forceinline__ int64_t read_start(const Ingest& g, int64_t read) {
    return g.read_off ? __ldg(&g.read_off[read]) : g.first_base + read * g.uniform_stride;
}
__device__ __forceinline__ int32_t read_npos(const Ingest& g, int64_t read) {
    return g.pos_off ? (int32_t)(__ldg(&g.pos_off[read + 1]) - __ldg(&g.pos_off[read])) : g.uniform_npos;
}

}  // namespace rb
