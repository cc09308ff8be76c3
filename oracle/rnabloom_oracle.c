/*
 * rnabloom_oracle.c -- CPU restatement of RNA-Bloom's k-mer / Bloom-filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path and the
 * "port" CPU baseline of bench.py.  Nothing in the product (rna-bloom_b200/, the C-ABI
 * library) may include, link or call it; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do.
 *
 * PARITY UNPINNED: the reference ships no golden vectors or asserting tests for this path
 * (test/Tests.java:27-67 prints bit strings only; NTHash.java:741-755 prints hashes without
 * expected values) and there is no JVM in this image, so the restatement below could not be
 * checked against reference output.  It is pinned only by (a) algebraic identities,
 * (b) an independent pure-Python restatement (oracle/pyref.py) and (c) the scratch vectors
 * of SURVEY.md section 4.  See DESIGN.md "Oracle".
 *
 * All citations are relative to /root/reference/src/rnabloom/.
 * Java semantics kept on purpose: signed 64-bit compares, wrapping long arithmetic,
 * int literal sign extension, rotate distance mod 64, signed bytes in the counting filter.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * a1  seeds / tables            bloom/hash/NTHash.java:30-43,100-168
 * ---------------------------------------------------------------------------------------- */
static const uint64_t SEED_A = 0x3c8bfbb395c60474ULL;
static const uint64_t SEED_C = 0x3193c18562a02b4cULL;
static const uint64_t SEED_G = 0x20323ed082572324ULL;
static const uint64_t SEED_T = 0x295549f54be24456ULL;
static const uint64_t MULTI_SEED = 0x90b45d39fb6da1faULL; /* NTHash.java:36 */
#define MULTI_SHIFT 27                                      /* NTHash.java:33 */
#define CP_OFF 0x07                                         /* NTHash.java:30 */

static uint64_t seed_tab[256];    /* NTHash.java:135-168 */
static uint64_t ms_tab[256][64];  /* NTHash.java:100-133 : msTab[c][i] = rotl(seed[c], i) */
static int tables_ready = 0;

static inline uint64_t rotl64(uint64_t v, int s) { s &= 63; return s ? (v << s) | (v >> (64 - s)) : v; }
static inline uint64_t rotr64(uint64_t v, int s) { s &= 63; return s ? (v >> s) | (v << (64 - s)) : v; }

static void init_tables(void) {
    if (tables_ready) return;
    memset(seed_tab, 0, sizeof seed_tab);
    /* rows 0..7 of seedTab: N T N G A A N C  (so that c & 7 of a base indexes its complement) */
    seed_tab[1] = SEED_T; seed_tab[3] = SEED_G; seed_tab[4] = SEED_A; seed_tab[5] = SEED_A; seed_tab[7] = SEED_C;
    seed_tab['A'] = SEED_A; seed_tab['C'] = SEED_C; seed_tab['G'] = SEED_G; seed_tab['T'] = SEED_T; seed_tab['U'] = SEED_T;
    seed_tab['a'] = SEED_A; seed_tab['c'] = SEED_C; seed_tab['g'] = SEED_G; seed_tab['t'] = SEED_T; seed_tab['u'] = SEED_T;
    for (int c = 0; c < 256; ++c)
        for (int i = 0; i < 64; ++i) ms_tab[c][i] = rotl64(seed_tab[c], i);
    tables_ready = 1;
}

ORC_API void orc_init(void) { init_tables(); }
ORC_API uint64_t orc_seed(int c) { init_tables(); return seed_tab[c & 255]; }
ORC_API uint64_t orc_mstab(int c, int i) { init_tables(); return ms_tab[c & 255][i & 63]; }

/* ------------------------------------------------------------------------------------------
 * a2  first k-mer hashes        NTHash.java:332-337 (NTP64), :367-373 (NTP64RC), :467-475 (NTPC64)
 * ---------------------------------------------------------------------------------------- */
ORC_API int64_t orc_ntp64(const uint8_t* seq, int k, int start) {
    init_tables();
    uint64_t h = 0;
    for (int i = 0; i < k; ++i) h ^= ms_tab[seq[start + i]][(k - 1 - i) % 64];
    return (int64_t)h;
}
ORC_API int64_t orc_ntp64rc(const uint8_t* seq, int k, int start) {
    init_tables();
    uint64_t h = 0;
    for (int i = 0; i < k; ++i) h ^= ms_tab[seq[start + i] & CP_OFF][i % 64];
    return (int64_t)h;
}
/* canonical: SIGNED compare (NTHash.java:474) */
ORC_API int64_t orc_ntpc64(const uint8_t* seq, int k, int start, int64_t* frh) {
    frh[0] = orc_ntp64(seq, k, start);
    frh[1] = orc_ntp64rc(seq, k, start);
    return (frh[1] < frh[0]) ? frh[1] : frh[0];
}

/* ------------------------------------------------------------------------------------------
 * a4  multi-hash expansion      NTHash.java:518-527
 *     tVal = bVal * (i ^ k * multiSeed)  -- '*' binds tighter than '^'; wrapping long math
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_ntm64(int64_t b, int64_t* hv, int k, int m) {
    hv[0] = b;
    for (int i = 1; i < m; ++i) {
        uint64_t t = (uint64_t)b * ((uint64_t)(int64_t)i ^ ((uint64_t)(int64_t)k * MULTI_SEED));
        t ^= t >> MULTI_SHIFT;
        hv[i] = (int64_t)t;
    }
}

/* ------------------------------------------------------------------------------------------
 * a3  rolling updates           NTHash.java:584-586 (fwd), :627-629 (RC), :491-495 (canonical)
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t roll_fwd(uint64_t f, uint8_t out, uint8_t in, int kMod64) {
    return rotl64(f, 1) ^ ms_tab[out][kMod64] ^ ms_tab[in][0];
}
static inline uint64_t roll_rc(uint64_t r, uint8_t out, uint8_t in, int k) {
    return rotr64(r, 1) ^ rotr64(seed_tab[out & CP_OFF], 1) ^ rotl64(seed_tab[in & CP_OFF], k - 1);
}
ORC_API int64_t orc_roll_fwd(int64_t f, int out, int in, int k) { init_tables(); return (int64_t)roll_fwd((uint64_t)f, (uint8_t)out, (uint8_t)in, k % 64); }
ORC_API int64_t orc_roll_rc(int64_t r, int out, int in, int k) { init_tables(); return (int64_t)roll_rc((uint64_t)r, (uint8_t)out, (uint8_t)in, k); }

/* ------------------------------------------------------------------------------------------
 * a5  combineHashValues         bloom/hash/HashFunction.java:260-266
 *     0x9e3779b9 is a Java *int* literal: it is negative and sign-extends to 0xFFFFFFFF9E3779B9
 * ---------------------------------------------------------------------------------------- */
ORC_API int64_t orc_combine(int64_t a, int64_t b) {
    uint64_t ua = (uint64_t)a, ub = (uint64_t)b;
    uint64_t lit = (uint64_t)(int64_t)(int32_t)0x9e3779b9u;
    return (int64_t)(ua ^ (ub + lit + (ua << 6) + (ub >> 2)));
}

/* ------------------------------------------------------------------------------------------
 * a6  k-mer iterators           NTHashIterator.java:47-69, CanonicalNTHashIterator.java:36-48,
 *                               ReverseComplementNTHashIterator.java:31-42
 * mode: 0 = stranded forward, 1 = stranded reverse-complement, 2 = canonical
 * Writes for every k-mer position pos in [start, end-k]: fh (forward hash, modes 0,2),
 * rh (reverse hash, modes 1,2), base (the hVals[0] the filters see).  Any may be NULL.
 * Returns the number of k-mers (0 if end-start < k).
 * ---------------------------------------------------------------------------------------- */
#define ORC_MODE_FWD 0
#define ORC_MODE_RC 1
#define ORC_MODE_CANON 2

ORC_API int64_t orc_kmer_hashes(const uint8_t* seq, int start, int end, int k, int mode,
                                int64_t* fh, int64_t* rh, int64_t* base) {
    init_tables();
    int max = end - k;
    if (max < start) return 0;
    const int kMod64 = k % 64;
    uint64_t f = 0, r = 0;
    int64_t n = 0;
    for (int pos = start; pos <= max; ++pos, ++n) {
        if (pos == start) {
            if (mode != ORC_MODE_RC) f = (uint64_t)orc_ntp64(seq, k, pos);
            if (mode != ORC_MODE_FWD) r = (uint64_t)orc_ntp64rc(seq, k, pos);
        } else {
            uint8_t out = seq[pos - 1], in = seq[pos - 1 + k];
            if (mode != ORC_MODE_RC) f = roll_fwd(f, out, in, kMod64);
            if (mode != ORC_MODE_FWD) r = roll_rc(r, out, in, k);
        }
        int64_t b;
        if (mode == ORC_MODE_FWD) b = (int64_t)f;
        else if (mode == ORC_MODE_RC) b = (int64_t)r;
        else b = ((int64_t)r < (int64_t)f) ? (int64_t)r : (int64_t)f;
        if (fh) fh[n] = (int64_t)f;
        if (rh) rh[n] = (int64_t)r;
        if (base) base[n] = b;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * a7  paired k-mer iterators    PairedNTHashIterator.java:55-85, CanonicalPaired...:39-60,
 *                               ReverseComplementPaired...:35-56
 * For pos in [start, end-k-d]: L at pos, R at pos+d; pair base hash P:
 *   stranded fwd: combine(L0,R0); stranded RC: combine(R0,L0) with RC hashes;
 *   canonical: signed min(combine(fL,fR), combine(rR,rL)).
 * ---------------------------------------------------------------------------------------- */
ORC_API int64_t orc_pair_hashes(const uint8_t* seq, int start, int end, int k, int d, int mode,
                                int64_t* baseL, int64_t* baseR, int64_t* baseP) {
    init_tables();
    int max = end - k - d;
    if (max < start) return 0;
    int nk = end - start - k + 1;
    int64_t* f = (int64_t*)malloc(sizeof(int64_t) * nk);
    int64_t* r = (int64_t*)malloc(sizeof(int64_t) * nk);
    int64_t* b = (int64_t*)malloc(sizeof(int64_t) * nk);
    /* the L and R iterators of the reference roll independently from pos and pos+d; a rolled
       hash equals the direct hash (tests/test_oracle.py::test_roll_equals_direct), so one
       pass over the segment yields both */
    orc_kmer_hashes(seq, start, end, k, mode, f, r, b);
    int64_t n = 0;
    for (int pos = start; pos <= max; ++pos, ++n) {
        int i = pos - start, j = i + d;
        int64_t p;
        if (mode == ORC_MODE_FWD) p = orc_combine(f[i], f[j]);
        else if (mode == ORC_MODE_RC) p = orc_combine(r[j], r[i]);
        else {
            int64_t p1 = orc_combine(f[i], f[j]);
            int64_t p2 = orc_combine(r[j], r[i]);
            p = p1 < p2 ? p1 : p2; /* Math.min on long: signed */
        }
        if (baseL) baseL[n] = b[i];
        if (baseR) baseR[n] = b[j];
        if (baseP) baseP[n] = p;
    }
    free(f); free(r); free(b);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * a8  index                     bloom/BloomFilter.java:108-111, CountingBloomFilter.java:101-104
 * ---------------------------------------------------------------------------------------- */
ORC_API int64_t orc_index(int64_t hash, int64_t size) { return (int64_t)(((uint64_t)hash >> 1) % (uint64_t)size); }

/* ------------------------------------------------------------------------------------------
 * a10 buffers                   bloom/buffer/UnsafeBitBuffer.java:42-77, UnsafeByteBuffer.java:40-150
 *     bit i <-> byte i/8, mask 1 << (i%8); ceil(size/8) bytes, zero-initialised; non-atomic RMW
 * a9  BloomFilter               bloom/BloomFilter.java:133-178
 * ---------------------------------------------------------------------------------------- */
#define ORC_MAX_HASH 16

typedef struct {
    int64_t size;     /* bits */
    int64_t nbytes;
    int num_hash;
    int k;            /* for NTM64 expansion of single base hashes (HashFunction.java:185-189) */
    uint8_t* bytes;
} orc_bf;

ORC_API orc_bf* orc_bf_create(int64_t size, int num_hash, int k) {
    orc_bf* f = (orc_bf*)calloc(1, sizeof *f);
    f->size = size; f->num_hash = num_hash; f->k = k;
    f->nbytes = size / 8 + ((size % 8) > 0 ? 1 : 0);
    f->bytes = (uint8_t*)calloc((size_t)(f->nbytes ? f->nbytes : 1), 1);
    if (!f->bytes) { free(f); return NULL; }
    return f;
}
ORC_API void orc_bf_destroy(orc_bf* f) { if (f) { free(f->bytes); free(f); } }
ORC_API void orc_bf_empty(orc_bf* f) { memset(f->bytes, 0, (size_t)f->nbytes); }
ORC_API uint8_t* orc_bf_bytes(orc_bf* f) { return f->bytes; }
ORC_API int64_t orc_bf_nbytes(orc_bf* f) { return f->nbytes; }
ORC_API int64_t orc_bf_size(orc_bf* f) { return f->size; }

static inline void bit_set(orc_bf* f, int64_t i) { f->bytes[i / 8] = (uint8_t)(f->bytes[i / 8] | (1u << (i % 8))); }
static inline int bit_get(const orc_bf* f, int64_t i) { return (f->bytes[i / 8] & (1u << (i % 8))) != 0; }
/* UnsafeByteBuffer.compareAndOr :59-69 -- returns true when the bit was already set */
static inline int bit_get_and_set(orc_bf* f, int64_t i) {
    uint8_t b = f->bytes[i / 8], nb = (uint8_t)(b | (1u << (i % 8)));
    if (b != nb) { f->bytes[i / 8] = nb; return 0; }
    return 1;
}

ORC_API void orc_bf_add(orc_bf* f, const int64_t* hv) {               /* BloomFilter.java:133-137 */
    for (int h = 0; h < f->num_hash; ++h) bit_set(f, orc_index(hv[h], f->size));
}
ORC_API int orc_bf_lookup(const orc_bf* f, const int64_t* hv) {       /* :170-178 */
    for (int h = 0; h < f->num_hash; ++h) if (!bit_get(f, orc_index(hv[h], f->size))) return 0;
    return 1;
}
ORC_API int orc_bf_lookup_then_add(orc_bf* f, const int64_t* hv) {    /* :147-155 (no short-circuit) */
    int found = 1;
    for (int h = 0; h < f->num_hash; ++h) found = bit_get_and_set(f, orc_index(hv[h], f->size)) && found;
    return found;
}
/* single base-hash overloads: BloomFilter.java:139-145,180-182 via HashFunction.java:185-189 */
ORC_API void orc_bf_add1(orc_bf* f, int64_t b) { int64_t hv[ORC_MAX_HASH]; orc_ntm64(b, hv, f->k, f->num_hash); orc_bf_add(f, hv); }
ORC_API int orc_bf_lookup1(const orc_bf* f, int64_t b) { int64_t hv[ORC_MAX_HASH]; orc_ntm64(b, hv, f->k, f->num_hash); return orc_bf_lookup(f, hv); }
ORC_API int orc_bf_lookup_then_add1(orc_bf* f, int64_t b) { int64_t hv[ORC_MAX_HASH]; orc_ntm64(b, hv, f->k, f->num_hash); return orc_bf_lookup_then_add(f, hv); }

/* a17 popcount / FPR            UnsafeByteBuffer.java:131-150, BloomFilter.java:185-194 */
ORC_API int64_t orc_bf_popcount(const orc_bf* f) {
    int64_t c = 0;
    for (int64_t i = 0; i < f->nbytes; ++i) c += __builtin_popcount(f->bytes[i]);
    return c;
}
ORC_API float orc_bf_fpr(const orc_bf* f) { return (float)pow((double)orc_bf_popcount(f) / (double)f->size, f->num_hash); }

/* a18 getExpectedSize           BloomFilter.java:196-199 */
ORC_API int64_t orc_expected_size(int64_t n, float fpr, int num_hash) {
    double r = (double)(-num_hash) / log(1 - exp(log(fpr) / (double)num_hash));
    return (int64_t)ceil((double)n * r);
}

/* ------------------------------------------------------------------------------------------
 * a12 MiniFloat                 util/MiniFloat.java:27-45
 * ---------------------------------------------------------------------------------------- */
static __thread uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
ORC_API void orc_seed_rng(uint64_t s) { rng_state = s ? s : 0x9E3779B97F4A7C15ULL; }
static inline double java_random(void) { /* stand-in for Math.random(): uniform double in [0,1) */
    uint64_t x = rng_state; x ^= x >> 12; x ^= x << 25; x ^= x >> 27; rng_state = x;
    return (double)((x * 0x2545F4914F6CDD1DULL) >> 11) * (1.0 / 9007199254740992.0);
}
ORC_API int8_t orc_minifloat_increment(int8_t b) {
    if (b <= 7 || (b < 127 && ((int32_t)(java_random() * 2147483647.0)) % (1 << ((b >> 3) - 1)) == 0))
        return (int8_t)(b + 1);
    return b;
}
ORC_API float orc_minifloat_to_float(int8_t b) {
    if (b <= 7) return (float)b;
    return (float)scalbn((double)((b & 7) | 8), (b >> 3) - 1);
}

/* ------------------------------------------------------------------------------------------
 * a11 CountingBloomFilter       bloom/CountingBloomFilter.java:170-194,196-222,235-251
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t size;   /* bytes */
    int num_hash;
    int k;
    int8_t* counts;
} orc_cbf;

ORC_API orc_cbf* orc_cbf_create(int64_t size, int num_hash, int k) {
    orc_cbf* f = (orc_cbf*)calloc(1, sizeof *f);
    f->size = size; f->num_hash = num_hash; f->k = k;
    f->counts = (int8_t*)calloc((size_t)(size ? size : 1), 1);
    if (!f->counts) { free(f); return NULL; }
    return f;
}
ORC_API void orc_cbf_destroy(orc_cbf* f) { if (f) { free(f->counts); free(f); } }
ORC_API void orc_cbf_empty(orc_cbf* f) { memset(f->counts, 0, (size_t)f->size); }
ORC_API int8_t* orc_cbf_bytes(orc_cbf* f) { return f->counts; }
ORC_API int64_t orc_cbf_size(orc_cbf* f) { return f->size; }

static int8_t cbf_min(const orc_cbf* f, const int64_t* hv) {           /* :172-183 */
    int8_t min = f->counts[orc_index(hv[0], f->size)];
    if (min != 0) {
        for (int h = 1; h < f->num_hash; ++h) {
            int8_t c = f->counts[orc_index(hv[h], f->size)];
            if (c < min) { min = c; if (min == 0) break; }
        }
    }
    return min;
}
/* returns the updated byte (incrementAndGet :196-222 returns toFloat(updated)) */
ORC_API int8_t orc_cbf_increment(orc_cbf* f, const int64_t* hv) {      /* :170-194 */
    int8_t min = cbf_min(f, hv);
    int8_t updated = orc_minifloat_increment(min);
    if (updated != min) {
        for (int h = 0; h < f->num_hash; ++h) {
            /* UnsafeByteBuffer.compareAndSwap :94-103 (non-atomic) */
            int64_t i = orc_index(hv[h], f->size);
            if (f->counts[i] == min) f->counts[i] = updated;
        }
    }
    return updated;
}
ORC_API float orc_cbf_increment_and_get(orc_cbf* f, const int64_t* hv) { return orc_minifloat_to_float(orc_cbf_increment(f, hv)); }
ORC_API float orc_cbf_get_count(const orc_cbf* f, const int64_t* hv) { /* :235-251 */
    int8_t min = f->counts[orc_index(hv[0], f->size)];
    for (int h = 1; h < f->num_hash; ++h) {
        int8_t c = f->counts[orc_index(hv[h], f->size)];
        if (c < min) min = c;
        if (min == 0) return 0;
    }
    return orc_minifloat_to_float(min);
}
ORC_API int8_t orc_cbf_increment1(orc_cbf* f, int64_t b) { int64_t hv[ORC_MAX_HASH]; orc_ntm64(b, hv, f->k, f->num_hash); return orc_cbf_increment(f, hv); }
ORC_API float orc_cbf_get_count1(const orc_cbf* f, int64_t b) { int64_t hv[ORC_MAX_HASH]; orc_ntm64(b, hv, f->k, f->num_hash); return orc_cbf_get_count(f, hv); }
ORC_API int64_t orc_cbf_popcount(const orc_cbf* f) {                   /* UnsafeByteBuffer.java:121-129 */
    int64_t c = 0;
    for (int64_t i = 0; i < f->size; ++i) if (f->counts[i] != 0) ++c;
    return c;
}
ORC_API float orc_cbf_fpr(const orc_cbf* f) { return (float)pow((double)orc_cbf_popcount(f) / (double)f->size, f->num_hash); }

/* ------------------------------------------------------------------------------------------
 * a19 CascadingBloomFilter      bloom/CascadingBloomFilter.java:34-100
 * ---------------------------------------------------------------------------------------- */
typedef struct { int num_levels; orc_bf** bfs; } orc_cascade;
ORC_API orc_cascade* orc_cascade_create(int64_t size, int num_hash, int k, int num_levels) {
    orc_cascade* c = (orc_cascade*)calloc(1, sizeof *c);
    c->num_levels = num_levels;
    c->bfs = (orc_bf**)calloc((size_t)num_levels, sizeof(orc_bf*));
    for (int i = 0; i < num_levels; ++i) c->bfs[i] = orc_bf_create(size / num_levels, num_hash, k);
    return c;
}
ORC_API void orc_cascade_destroy(orc_cascade* c) { for (int i = 0; i < c->num_levels; ++i) orc_bf_destroy(c->bfs[i]); free(c->bfs); free(c); }
ORC_API orc_bf* orc_cascade_level(orc_cascade* c, int i) { return c->bfs[i]; }
ORC_API void orc_cascade_add1(orc_cascade* c, int64_t b) {             /* :66-72 */
    for (int i = 0; i < c->num_levels; ++i) if (!orc_bf_lookup_then_add1(c->bfs[i], b)) break;
}
ORC_API int orc_cascade_lookup1(orc_cascade* c, int64_t b) { return orc_bf_lookup1(c->bfs[c->num_levels - 1], b); } /* :79-85 */
ORC_API int orc_cascade_lookup_then_add1(orc_cascade* c, int64_t b) {  /* :93-100 */
    for (int i = 0; i < c->num_levels; ++i) if (!orc_bf_lookup_then_add1(c->bfs[i], b)) return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * a13-a15 BloomFilterDeBruijnGraph   graph/BloomFilterDeBruijnGraph.java:75-104,405-570
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    orc_bf* dbgbf; orc_cbf* cbf; orc_bf* rpkbf; orc_bf* fpkbf;
    int k, stranded, hd, hc, hp, hmax;
    int d_read, d_frag;
} orc_graph;

ORC_API orc_graph* orc_graph_create(int64_t dbg_bits, int64_t cbf_bytes, int64_t pkbf_bits,
                                    int hd, int hc, int hp, int k, int stranded, int use_read_pairs) {
    orc_graph* g = (orc_graph*)calloc(1, sizeof *g);
    g->k = k; g->stranded = stranded; g->hd = hd; g->hc = hc; g->hp = hp;
    g->hmax = hd > hc ? hd : hc;
    g->d_read = -1; g->d_frag = -1;
    g->dbgbf = orc_bf_create(dbg_bits, hd, k);
    g->cbf = orc_cbf_create(cbf_bytes, hc, k);
    if (use_read_pairs) g->rpkbf = orc_bf_create(pkbf_bits, hp, k);
    return g;
}
ORC_API void orc_graph_init_fpkbf(orc_graph* g, int64_t bits, int hp) {   /* :352-359 */
    if (!g->fpkbf) g->fpkbf = orc_bf_create(bits, hp, g->k); else orc_bf_empty(g->fpkbf);
}
ORC_API void orc_graph_destroy(orc_graph* g) {
    orc_bf_destroy(g->dbgbf); orc_cbf_destroy(g->cbf); orc_bf_destroy(g->rpkbf); orc_bf_destroy(g->fpkbf); free(g);
}
ORC_API orc_bf* orc_graph_dbgbf(orc_graph* g) { return g->dbgbf; }
ORC_API orc_cbf* orc_graph_cbf(orc_graph* g) { return g->cbf; }
ORC_API orc_bf* orc_graph_rpkbf(orc_graph* g) { return g->rpkbf; }
ORC_API orc_bf* orc_graph_fpkbf(orc_graph* g) { return g->fpkbf; }
ORC_API void orc_graph_set_distances(orc_graph* g, int d_read, int d_frag) { g->d_read = d_read; g->d_frag = d_frag; }

ORC_API void orc_graph_add(orc_graph* g, const int64_t* hv) {          /* :405-412 */
    if (orc_bf_lookup_then_add(g->dbgbf, hv)) orc_cbf_increment(g->cbf, hv);
}
ORC_API void orc_graph_add_count_if_present(orc_graph* g, const int64_t* hv) { /* :424-428 */
    if (orc_bf_lookup(g->dbgbf, hv) && orc_cbf_get_count(g->cbf, hv) > 0) orc_cbf_increment(g->cbf, hv);
}
ORC_API void orc_graph_add_dbg_only(orc_graph* g, const int64_t* hv) { orc_bf_add(g->dbgbf, hv); } /* :434-436 */
ORC_API int orc_graph_contains(const orc_graph* g, const int64_t* hv) { return orc_bf_lookup(g->dbgbf, hv); } /* :538-540 */
ORC_API float orc_graph_get_count(const orc_graph* g, const int64_t* hv) {   /* :562-570 */
    if (orc_bf_lookup(g->dbgbf, hv)) return orc_cbf_get_count(g->cbf, hv) + 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * f1  neighbours of a k-mer      graph/Kmer.java:213-253 (getPredecessors / getSuccessors), graph/CanonicalKmer.java:232-271;
 *     hash side: bloom/hash/SuccessorsNTHashIterator.java:52-63, PredecessorsNTHashIterator.java:54-65,
 *     CanonicalSuccessorsNTHashIterator.java:56-72, CanonicalPredecessorsNTHashIterator.java:56-72
 * For the four candidate bases A,C,G,T (SeqUtils.NUCLEOTIDES_BYTES): the neighbour's forward / reverse hash and graph.getCount of it.
 * char_out = bytes[0] for successors, bytes[k-1] for predecessors.  The caller applies "count >= minKmerCov".
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_graph_neighbors(const orc_graph* g, int64_t fh, int64_t rh, int char_out, int successors, float* counts4, int64_t* f4,
                                 int64_t* r4) {
    static const uint8_t nt[4] = {'A', 'C', 'G', 'T'};
    init_tables();
    const int k = g->k, km1 = (k - 1) % 64;
    const uint8_t co = (uint8_t)char_out;
    uint64_t tf, tr = 0;
    if (successors) {
        tf = rotl64((uint64_t)fh, 1) ^ ms_tab[co][k % 64];
        if (!g->stranded) tr = rotr64((uint64_t)rh, 1) ^ ms_tab[co & CP_OFF][63];
    } else {
        tf = rotr64((uint64_t)fh, 1) ^ ms_tab[co][63];
        if (!g->stranded) tr = rotl64((uint64_t)rh, 1) ^ ms_tab[co & CP_OFF][k % 64];
    }
    for (int i = 0; i < 4; ++i) {
        const uint8_t ci = nt[i];
        const uint64_t f = tf ^ (successors ? ms_tab[ci][0] : ms_tab[ci][km1]);
        uint64_t r = 0, b = f;
        if (!g->stranded) {
            r = tr ^ (successors ? ms_tab[ci & CP_OFF][km1] : ms_tab[ci & CP_OFF][0]);
            b = ((int64_t)r < (int64_t)f) ? r : f;           /* Math.min on long */
        }
        int64_t hv[ORC_MAX_HASH];
        orc_ntm64((int64_t)b, hv, k, g->hmax);
        counts4[i] = orc_graph_get_count(g, hv);
        if (f4) f4[i] = (int64_t)f;
        if (r4) r4[i] = (int64_t)r;
    }
}

/* Kmer.getLeftVariants / getRightVariants (graph/Kmer.java:357-405): counts of the k-mers that differ in the first (side 0) / last (side 1)
 * base, computed here from the SEQUENCE (NTP64 / NTP64RC of the edited k-mer), not incrementally: an independent check of the variant
 * iterators' algebra.  counts4[c] = graph.getCount of the k-mer with base "ACGT"[c] there (own base: the k-mer itself). */
ORC_API void orc_graph_variants(const orc_graph* g, const uint8_t* kmer, int side, float* counts4, int64_t* f4, int64_t* r4) {
    static const uint8_t nt[4] = {'A', 'C', 'G', 'T'};
    init_tables();
    const int k = g->k;
    uint8_t* tmp = (uint8_t*)malloc((size_t)k);
    memcpy(tmp, kmer, (size_t)k);
    for (int c = 0; c < 4; ++c) {
        tmp[side ? k - 1 : 0] = nt[c];
        const int64_t f = orc_ntp64(tmp, k, 0), r = orc_ntp64rc(tmp, k, 0);
        const int64_t b = g->stranded ? f : (r < f ? r : f);
        int64_t hv[ORC_MAX_HASH];
        orc_ntm64(b, hv, k, g->hmax);
        counts4[c] = orc_graph_get_count(g, hv);
        if (f4) f4[c] = f;
        if (r4) r4[c] = g->stranded ? 0 : r;
    }
    free(tmp);
}
/* GraphUtils.greedyExtendRight / greedyExtendLeft (util/GraphUtils.java:1961-1976, :1925-1958) with lookahead <= 1, where
 * greedyExtendRightOnce (:501-527) reduces to: no candidate -> stop; otherwise the candidate (getSuccessors: count > 0, A C G T order,
 * graph/Kmer.java:213-231) with the largest count, the first one winning ties.  kmer: k ASCII bases; out: up to `bound` added bases in
 * walking order.  Returns the number of k-mers added. */
ORC_API int orc_graph_greedy_extend(const orc_graph* g, const uint8_t* kmer, int right, int bound, float min_cov, uint8_t* out) {
    static const uint8_t nt[4] = {'A', 'C', 'G', 'T'};
    init_tables();
    const int k = g->k;
    uint8_t* cur = (uint8_t*)malloc((size_t)k);
    memcpy(cur, kmer, (size_t)k);
    int n = 0;
    for (; n < bound; ++n) {
        const int64_t fh = orc_ntp64(cur, k, 0), rh = orc_ntp64rc(cur, k, 0);
        float c4[4];
        orc_graph_neighbors(g, fh, rh, right ? cur[0] : cur[k - 1], right, c4, NULL, NULL);
        float best = -1.f; int bi = -1;
        for (int c = 0; c < 4; ++c) if (c4[c] >= min_cov && c4[c] > best) { best = c4[c]; bi = c; }
        if (bi < 0) break;
        out[n] = nt[bi];
        if (right) { memmove(cur, cur + 1, (size_t)(k - 1)); cur[k - 1] = nt[bi]; }
        else { memmove(cur + 1, cur, (size_t)(k - 1)); cur[0] = nt[bi]; }
    }
    free(cur);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * a20 segmentation              RNABloom.java:567-595 (FASTQ), :677-716 (FASTA); util/SeqUtils.java:1430-1438
 *   qualPattern = [chars >= '!'+minQual .. '~']{k,}   seqPattern = [ACGTUacgtu]{k,}
 *   -> maximal runs of length >= k; the sequence pattern is searched inside every quality run.
 * Writes (start,end) pairs; returns the number of segments.  qual may be NULL (FASTA path).
 * ---------------------------------------------------------------------------------------- */
static inline int is_nt(uint8_t c) {
    switch (c) { case 'A': case 'C': case 'G': case 'T': case 'U': case 'a': case 'c': case 'g': case 't': case 'u': return 1; default: return 0; }
}
static int seq_runs(const uint8_t* seq, int from, int to, int k, int32_t* starts, int32_t* ends, int n) {
    int i = from;
    while (i < to) {
        if (!is_nt(seq[i])) { ++i; continue; }
        int j = i; while (j < to && is_nt(seq[j])) ++j;
        if (j - i >= k) { if (starts) { starts[n] = i; ends[n] = j; } ++n; }
        i = j;
    }
    return n;
}
ORC_API int orc_segment(const uint8_t* seq, const uint8_t* qual, int len, int k, int min_qual,
                        int32_t* starts, int32_t* ends) {
    if (len < k) return 0;                       /* RNABloom.java:562-565 */
    if (!qual) return seq_runs(seq, 0, len, k, starts, ends, 0);
    int n = 0, i = 0;
    const int lo = '!' + min_qual, hi = '~';
    while (i < len) {
        if (qual[i] < lo || qual[i] > hi) { ++i; continue; }
        int j = i; while (j < len && qual[j] >= lo && qual[j] <= hi) ++j;
        if (j - i >= k) n = seq_runs(seq, i, j, k, starts, ends, n);
        i = j;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * a20 caller contract: one read through the insert workers
 *   FastqToGraphWorker.run RNABloom.java:551-634, FastaToGraphWorker :677-716,
 *   FragmentsToGraphWorker :1489-1530
 * flags mirror the C-ABI (include/rnabloom_gpu.h)
 * ---------------------------------------------------------------------------------------- */
#define ORC_F_REVCOMP 1
#define ORC_F_ADD_COUNT_IF_PRESENT 2
#define ORC_F_DBG_ONLY 4
#define ORC_F_STORE_READ_PAIRS 8
#define ORC_F_STORE_FRAG_PAIRS 16

static int graph_mode(const orc_graph* g, int flags) {
    if (!g->stranded) return ORC_MODE_CANON;     /* CanonicalHashFunction.java:179-206 ignores revcomp */
    return (flags & ORC_F_REVCOMP) ? ORC_MODE_RC : ORC_MODE_FWD;
}

ORC_API int64_t orc_graph_add_segment(orc_graph* g, const uint8_t* seq, int start, int end, int flags) {
    const int k = g->k, mode = graph_mode(g, flags);
    int nk = end - start - k + 1;
    if (nk <= 0) return 0;
    int64_t* base = (int64_t*)malloc(sizeof(int64_t) * nk);
    int64_t hv[ORC_MAX_HASH];
    orc_kmer_hashes(seq, start, end, k, mode, NULL, NULL, base);
    for (int i = 0; i < nk; ++i) {
        orc_ntm64(base[i], hv, k, g->hmax);
        if (flags & ORC_F_DBG_ONLY) orc_graph_add_dbg_only(g, hv);
        else if (flags & ORC_F_ADD_COUNT_IF_PRESENT) orc_graph_add_count_if_present(g, hv);
        else orc_graph_add(g, hv);
    }
    if ((flags & ORC_F_STORE_READ_PAIRS) && g->rpkbf && g->d_read > 0) {
        int np = (int)orc_pair_hashes(seq, start, end, k, g->d_read, mode, NULL, NULL, base);
        for (int i = 0; i < np; ++i) orc_bf_add1(g->rpkbf, base[i]);   /* graph :455-457 */
    }
    if ((flags & ORC_F_STORE_FRAG_PAIRS) && g->fpkbf && g->d_frag > 0) {
        int np = (int)orc_pair_hashes(seq, start, end, k, g->d_frag, mode, NULL, NULL, base);
        for (int i = 0; i < np; ++i) orc_bf_add1(g->fpkbf, base[i]);   /* graph :459-461 */
    }
    free(base);
    return nk;
}

/* One FASTQ/FASTA record: segmentation + per-segment insert.  Returns k-mers inserted. */
ORC_API int64_t orc_graph_add_read(orc_graph* g, const uint8_t* seq, const uint8_t* qual, int len,
                                   int min_qual, int flags) {
    int32_t st[4096], en[4096];
    int32_t *ps = st, *pe = en;
    int cap = len / (g->k > 0 ? g->k : 1) + 1;
    if (cap > 4096) { ps = (int32_t*)malloc(sizeof(int32_t) * cap); pe = (int32_t*)malloc(sizeof(int32_t) * cap); }
    int n = orc_segment(seq, qual, len, g->k, min_qual, ps, pe);
    int64_t total = 0;
    for (int i = 0; i < n; ++i) total += orc_graph_add_segment(g, seq, ps[i], pe[i], flags);
    if (ps != st) { free(ps); free(pe); }
    return total;
}

/* a16 getKmers                  HashFunction.java:55-85, CanonicalHashFunction.java:46-78
 * Whole-sequence lookup: every position is hashed through (invalid chars contribute seed 0);
 * a k-mer that covers an invalid nucleotide gets count 0. */
ORC_API int64_t orc_graph_count_seq(const orc_graph* g, const uint8_t* seq, int start, int end,
                                    float* counts, int64_t* fh, int64_t* rh) {
    const int k = g->k;
    int nk = end - start - k + 1;
    if (nk <= 0) return 0;
    int mode = g->stranded ? ORC_MODE_FWD : ORC_MODE_CANON;
    int64_t* base = (int64_t*)malloc(sizeof(int64_t) * nk);
    int64_t hv[ORC_MAX_HASH];
    orc_kmer_hashes(seq, start, end, k, mode, fh, (mode == ORC_MODE_CANON) ? rh : NULL, base);
    int bad = 0; /* number of invalid chars inside the current window */
    for (int i = start; i < start + k - 1; ++i) bad += !is_nt(seq[i]);
    for (int i = 0; i < nk; ++i) {
        bad += !is_nt(seq[start + i + k - 1]);
        if (bad) counts[i] = 0;
        else { orc_ntm64(base[i], hv, k, g->hmax); counts[i] = orc_graph_get_count(g, hv); }
        bad -= !is_nt(seq[start + i]);
    }
    free(base);
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic reads (bench + fixtures): counter-based so the CUDA generator and this one agree
 * bit for bit.  Not part of the reference; see DESIGN.md "Synthetic workload".
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline int genome_base(uint64_t seed, uint64_t pos) { return (int)(splitmix64(seed ^ splitmix64(pos)) & 3); }

/* read r of length L from a virtual genome of G bases; err_ppm substitutions per million bases.
 * out = ASCII bases.  strand chosen per read; reads are emitted 5'->3' on their strand. */
ORC_API void orc_synth_read(uint64_t seed, uint64_t genome_len, uint64_t r, int L, uint32_t err_ppm, uint8_t* out) {
    static const char NT[4] = {'A', 'C', 'G', 'T'};
    uint64_t h = splitmix64(seed * 0x100000001B3ULL + 2 * r + 1);
    uint64_t pos = (h >> 1) % (genome_len - (uint64_t)L + 1);
    int rc = (int)(h & 1);
    for (int i = 0; i < L; ++i) {
        int b = rc ? 3 - genome_base(seed, pos + (uint64_t)(L - 1 - i)) : genome_base(seed, pos + (uint64_t)i);
        uint64_t e = splitmix64((seed + 0x5851F42D4C957F2DULL) ^ splitmix64(r * 1024 + (uint64_t)i));
        if ((uint32_t)(e % 1000000u) < err_ppm) b = (b + 1 + (int)((e >> 40) % 3)) & 3;
        out[i] = (uint8_t)NT[b];
    }
}
ORC_API void orc_synth_reads(uint64_t seed, uint64_t genome_len, uint64_t first, uint64_t n, int L, uint32_t err_ppm, uint8_t* out) {
    for (uint64_t r = 0; r < n; ++r) orc_synth_read(seed, genome_len, first + r, L, err_ppm, out + r * (uint64_t)L);
}

/* util/SeqBitsUtils.java:138-247 + io/NucleotideBitsWriter.java:24-31: one record = intToFourBytes(len) (big endian, :209) + seqToBits(seq):
 * getByte(i,j,k,l) = i*64 + j*16 + k*4 + l - 128 (:158-160), missing bases of the last tetramer = 0 (:236-243).  ACGTU only (anything else is
 * a random base in the reference).  Returns the record's size; out may be NULL. */
ORC_API int64_t orc_2bit_record(const uint8_t* seq, int len, uint8_t* out) {
    int64_t at = 0;
    if (out) { out[0] = (uint8_t)(len >> 24); out[1] = (uint8_t)(len >> 16); out[2] = (uint8_t)(len >> 8); out[3] = (uint8_t)len; }
    at = 4;
    for (int i = 0; i < len; i += 4) {
        int idx[4] = {0, 0, 0, 0};
        for (int j = 0; j < 4 && i + j < len; ++j) {
            switch (seq[i + j]) {
                case 'A': case 'a': idx[j] = 0; break;
                case 'C': case 'c': idx[j] = 1; break;
                case 'G': case 'g': idx[j] = 2; break;
                default: idx[j] = 3;   /* T t U u */
            }
        }
        if (out) out[at] = (uint8_t)(int8_t)(idx[0] * 64 + idx[1] * 16 + idx[2] * 4 + idx[3] - 128);
        ++at;
    }
    return at;
}
/* bitsToSeq (:249-262): tetramer lookup by byte + 128 */
ORC_API void orc_2bit_decode(const uint8_t* bytes, int len, uint8_t* seq) {
    static const char NT[4] = {'A', 'C', 'G', 'T'};
    for (int i = 0; i < len; ++i) {
        int v = (int)(int8_t)bytes[i / 4] + 128;
        seq[i] = (uint8_t)NT[(v >> (2 * (3 - (i & 3)))) & 3];
    }
}

/* long reads (BASELINE configs[4]): the twin of k_synth_long_reads (rna-bloom_b200/csrc/rb_kernels.cuh), integer arithmetic only */
ORC_API int orc_synth_long_len(uint64_t seed, uint64_t r) {
    uint64_t h = splitmix64(seed * 0x100000001B3ULL + 2 * r);
    return 500 + (int)(h % 1000u) + (int)((h >> 20) % 1000u) + (int)((h >> 40) % 1000u);
}
ORC_API int orc_synth_long_read(uint64_t seed, uint64_t genome_len, uint64_t r, uint32_t sub_ppm, uint32_t ins_ppm, uint32_t del_ppm, uint8_t* out) {
    static const char NT[4] = {'A', 'C', 'G', 'T'};
    int L = orc_synth_long_len(seed, r);
    uint64_t h = splitmix64(seed * 0x100000001B3ULL + 2 * r + 1);
    uint64_t span = (uint64_t)(2 * L + 64);
    uint64_t p0 = (h >> 1) % (genome_len - span + 1);
    int rc = (int)(h & 1);
    uint64_t gpos = 0;
    for (int i = 0; i < L; ++i) {
        int b;
        for (int tries = 0;; ++tries) {
            uint64_t e = splitmix64((seed + 0x5851F42D4C957F2DULL) ^ splitmix64(r * 8192 + (uint64_t)i * 4 + (uint64_t)(tries < 3 ? tries : 3)));
            uint32_t u = (uint32_t)(e % 1000000u);
            if (u < del_ppm && tries < 3 && gpos + 1 < span) { ++gpos; continue; }
            int gb = rc ? 3 - genome_base(seed, p0 + span - 1 - gpos) : genome_base(seed, p0 + gpos);
            if (u < del_ppm + ins_ppm) { b = (int)((e >> 40) & 3); break; }
            b = gb;
            if (u < del_ppm + ins_ppm + sub_ppm) b = (gb + 1 + (int)((e >> 40) % 3)) & 3;
            if (gpos + 1 < span) ++gpos;
            break;
        }
        out[i] = (uint8_t)NT[b];
    }
    return L;
}

/* ------------------------------------------------------------------------------------------
 * CPU baseline: reference-faithful threaded insert / lookup (RNABloom.java:1189-1205 thread model:
 * N workers share one graph and pull reads from one synchronized reader, io/FastqReader.java:140-151;
 * filter RMWs stay non-atomic like UnsafeByteBuffer).
 * reads: n fixed-length ASCII reads of length L.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    orc_graph* g; const uint8_t* reads; int64_t n; int L; int flags; int lookup;
    const int64_t* off;   /* ragged reads: read r = reads[off[r], off[r+1]); NULL: fixed length L */
    int64_t cursor; pthread_mutex_t mu; int64_t kmers; double checksum;
} mt_job;

static void* mt_worker(void* arg) {
    mt_job* j = (mt_job*)arg;
    int64_t kmers = 0; double cs = 0;
    float* counts = (float*)malloc(sizeof(float) * (size_t)(j->L + 1));
    rng_state = 0x1234567ULL + (uint64_t)(uintptr_t)&kmers;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int64_t r = j->cursor++;
        pthread_mutex_unlock(&j->mu);
        if (r >= j->n) break;
        const uint8_t* s = j->off ? j->reads + j->off[r] : j->reads + r * (int64_t)j->L;
        const int len = j->off ? (int)(j->off[r + 1] - j->off[r]) : j->L;
        if (j->lookup) {
            int64_t nk = orc_graph_count_seq(j->g, s, 0, len, counts, NULL, NULL);
            for (int64_t i = 0; i < nk; ++i) cs += counts[i];
            kmers += nk;
        } else {
            kmers += orc_graph_add_read(j->g, s, NULL, len, 0, j->flags);
        }
    }
    free(counts);
    pthread_mutex_lock(&j->mu);
    j->kmers += kmers; j->checksum += cs;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}

static int64_t run_mt(orc_graph* g, const uint8_t* reads, const int64_t* off, int64_t n, int L, int flags, int lookup, int nthreads, double* checksum);
ORC_API int64_t orc_graph_run_mt(orc_graph* g, const uint8_t* reads, int64_t n, int L, int flags,
                                 int lookup, int nthreads, double* checksum) {
    return run_mt(g, reads, NULL, n, L, flags, lookup, nthreads, checksum);
}
/* ragged reads: off has n + 1 entries; max_len bounds the longest read */
ORC_API int64_t orc_graph_run_mt_ragged(orc_graph* g, const uint8_t* bases, const int64_t* off, int64_t n, int max_len, int flags,
                                        int lookup, int nthreads, double* checksum) {
    return run_mt(g, bases, off, n, max_len, flags, lookup, nthreads, checksum);
}
ORC_API int64_t orc_synth_long_reads(uint64_t seed, uint64_t genome_len, uint64_t first, uint64_t n, uint32_t sub_ppm, uint32_t ins_ppm,
                                     uint32_t del_ppm, uint8_t* out, int64_t* off) {
    int64_t at = 0;
    for (uint64_t r = 0; r < n; ++r) { off[r] = at; at += orc_synth_long_read(seed, genome_len, first + r, sub_ppm, ins_ppm, del_ppm, out + at); }
    off[n] = at;
    return at;
}
static int64_t run_mt(orc_graph* g, const uint8_t* reads, const int64_t* off, int64_t n, int L, int flags, int lookup, int nthreads, double* checksum) {
    init_tables();
    mt_job j; memset(&j, 0, sizeof j);
    j.g = g; j.reads = reads; j.n = n; j.L = L; j.flags = flags; j.lookup = lookup; j.off = off;
    pthread_mutex_init(&j.mu, NULL);
    if (nthreads < 1) nthreads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, mt_worker, &j);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    pthread_mutex_destroy(&j.mu);
    if (checksum) *checksum = j.checksum;
    return j.kmers;
}
