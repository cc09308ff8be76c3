// rnabloom_gpu.cu -- host side of librnabloom_gpu.so: the C-ABI of include/rnabloom_gpu.h over the kernels in
// rb_kernels.cuh.  Owns device memory, streams, sub-batching and the reference's file formats.  No torch, no oracle:
// if CUDA is unavailable every entry point fails with RB_ECUDA.  Citations: /root/reference/src/rnabloom/.
#include "../../include/rnabloom_gpu.h"
#include "rb_kernels.cuh"
#include "rb_sliced.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace rb;

// ------------------------------------------------------------------------------------------------------------------
struct rb_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::recursive_mutex mu;
    std::string err;
    uint64_t rng_seed = 0x243F6A8885A308D3ULL;
    int64_t subbatch_kmers = 1LL << 25;
    bool subbatch_user_set = false;   // the sliced engine picks its own (much larger) round size unless the caller chose one
    int64_t launches = 0;
    int sm_count = 148;
    // claim table (DESIGN.md "Linearisation")
    unsigned long long* claim = nullptr;
    int64_t claim_cap = 0, claim_used = 0;
    const void* claim_owner = nullptr;  // the bit array the current claims refer to
    // staging (host-pointer entry points)
    void* stage[32] = {};
    int64_t stage_bytes[32] = {};
    cudaStream_t copy_stream = nullptr;             // D2H of lookup results overlaps the next launch
    cudaEvent_t ev_kernel[2] = {nullptr, nullptr};  // kernel of parity p finished (results ready in staging p)
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};    // D2H out of staging p finished (staging p reusable)
    int64_t count_launch_index = 0;
    cudaEvent_t ticket_ev[8] = {};                  // rb_graph_count_reads_async: ticket t completes with event (t - 1) % 8 on the copy stream
    int64_t ticket_seq = 0;
    unsigned long long* scratch = nullptr;  // 8-byte device scalar
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // per-kernel CUDA-event timing (rb_ctx_profile_enable / rb_ctx_profile_read): bench.py's roofline numbers come from here
    struct ProfSpan { const char* name; cudaEvent_t a, b; };
    bool prof_on = false;
    const char* prof_pending = nullptr;
    cudaEvent_t prof_pending_ev = nullptr;
    std::vector<ProfSpan> prof_spans;
    std::vector<cudaEvent_t> prof_pool;
};
struct CellStore;
struct rb_filter {
    rb_ctx* ctx;
    int kind;
    int64_t size;      // bits or bytes
    int64_t nbytes;    // logical byte length (file size)
    int64_t alloc;     // allocated bytes (multiple of 256, >= nbytes + 16)
    int num_hash, k;
    uint32_t* dev;
    bool in_graph;
    CellStore* cs;     // dbgbf / cbf of a graph whose sliced engine works on co-located cells (SlGeom::cells); nullptr otherwise
};
// The co-located copy of a graph's dbgbf + cbf (rb_sliced.cuh SlGeom::cells).  Exactly one of the two representations is current:
// in_cells = the cell array (the sliced engine's rounds), else the logical arrays (everything else).  Whoever needs the other one
// converts first (one streaming pass over both, ~5 ms for 16 GiB): cells_ensure_logical at the top of every entry point that reads or
// writes the logical arrays, cells_ensure_cells before a sliced round.
struct CellStore {
    uint32_t* cells;   // C 16-bit cells
    int64_t C;         // counters of the share (= cbf bytes)
    int q;             // chunks of C bits in the dbgbf share
    bool in_cells;
    rb_filter *dbg, *cbf;
};
struct SlicedEngine;
struct rb_graph {
    rb_ctx* ctx;
    rb_filter *dbg, *cbf, *rpk, *fpk;
    int k, stranded, hd, hc, hp, hmax;
    int d_read, d_frag;
    int engine;            // RB_ENGINE_AUTO / RB_ENGINE_DIRECT / RB_ENGINE_SLICED
    SlicedEngine* se;      // lazily built
    bool cells_tried;      // the cell store (g->dbg->cs) is created with the first sliced engine, once
};

static thread_local std::string g_create_err;

static int32_t fail(rb_ctx* c, int32_t code, const std::string& msg) {
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? RB_ENOMEM : RB_ECUDA,                      \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                              \
    } while (0)
#define LOCK(c) std::lock_guard<std::recursive_mutex> lock_((c)->mu); cudaSetDevice((c)->device)
static void prof_begin(rb_ctx* c, const char* name);
static void prof_end(rb_ctx* c);
#define PROF(name) do { if (ctx->prof_on) prof_begin(ctx, name); } while (0)
#define LAUNCH_CHECK()                                                                                    \
    do {                                                                                                  \
        ++ctx->launches;                                                                                  \
        if (ctx->prof_pending) prof_end(ctx);                                                             \
        cudaError_t e_ = cudaGetLastError();                                                              \
        if (e_ != cudaSuccess) return fail(ctx, RB_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
    } while (0)

static cudaEvent_t prof_event(rb_ctx* c) {
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
static void prof_begin(rb_ctx* c, const char* name) {
    c->prof_pending = name;
    c->prof_pending_ev = prof_event(c);
    cudaEventRecord(c->prof_pending_ev, c->stream);
}
static void prof_end(rb_ctx* c) {
    cudaEvent_t b = prof_event(c);
    cudaEventRecord(b, c->stream);
    c->prof_spans.push_back({c->prof_pending, c->prof_pending_ev, b});
    c->prof_pending = nullptr;
}
extern "C" int32_t rb_ctx_profile_enable(rb_ctx* ctx, int32_t on) {
    if (!ctx) return RB_EINVAL;
    LOCK(ctx);
    ctx->prof_on = on != 0;
    return RB_OK;
}
// Sums the device time of every profiled launch since the last read, per kernel name.  names receives the names joined by '\n'.
extern "C" int32_t rb_ctx_profile_read(rb_ctx* ctx, char* names, int64_t names_len, float* ms, int32_t* calls, int32_t max_entries, int32_t* n_out) {
    if (!ctx || !names || !ms || !calls || !n_out || names_len < 1) return RB_EINVAL;
    LOCK(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<std::string> order;
    std::vector<double> tot;
    std::vector<int> cnt;
    for (auto& sp : ctx->prof_spans) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, sp.a, sp.b));
        size_t i = 0;
        while (i < order.size() && order[i] != sp.name) ++i;
        if (i == order.size()) { order.push_back(sp.name); tot.push_back(0.0); cnt.push_back(0); }
        tot[i] += t; cnt[i] += 1;
        ctx->prof_pool.push_back(sp.a); ctx->prof_pool.push_back(sp.b);
    }
    ctx->prof_spans.clear();
    std::string joined;
    int32_t n = 0;
    for (size_t i = 0; i < order.size() && n < max_entries; ++i) {
        if ((int64_t)(joined.size() + order[i].size() + 2) > names_len) break;
        if (n) joined += "\n";
        joined += order[i];
        ms[n] = (float)tot[i]; calls[n] = cnt[i];
        ++n;
    }
    memcpy(names, joined.c_str(), joined.size() + 1);
    *n_out = n;
    return RB_OK;
}

static FastMod make_fm(int64_t size) {
    FastMod fm;
    fm.size = (uint64_t)size;
    fm.pow2 = (size & (size - 1)) == 0;
    fm.mask = (uint64_t)size - 1;
    fm.magic = fm.pow2 ? 0 : (uint64_t)((((unsigned __int128)1) << 64) / (unsigned __int128)size);
    return fm;
}
static HashMults make_hm(int k) {
    HashMults hm;
    for (int i = 0; i < kMaxHash; ++i) hm.m[i] = (uint64_t)(int64_t)i ^ ((uint64_t)(int64_t)k * kMultiSeed);
    return hm;
}
static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid-stride kernels: enough CTAs to fill the part (the host emulation of tests/emu runs one fiber per CUDA thread: a handful there)
static int cells_grid(rb_ctx* ctx, const CellStore* cs) {
#ifdef RB_EMU
    (void)ctx;
    return (int)std::min<int64_t>(div_up(cs->C >> 5, kSlThreads), 2);
#else
    return (int)std::min<int64_t>(div_up(cs->C >> 5, kSlThreads), (int64_t)ctx->sm_count * 32);
#endif
}
static int32_t cells_ensure_logical(rb_ctx* ctx, CellStore* cs) {
    if (!cs || !cs->in_cells) return RB_OK;
    PROF("k_cells_unpack");
    RB_LAUNCH(cells_grid(ctx, cs), kSlThreads, 0, ctx->stream, k_cells_unpack)(cs->cells, cs->dbg->dev, cs->cbf->dev, cs->C, cs->q);
    LAUNCH_CHECK();
    cs->in_cells = false;
    return RB_OK;
}
static int32_t cells_ensure_cells(rb_ctx* ctx, CellStore* cs) {
    if (!cs || cs->in_cells) return RB_OK;
    PROF("k_cells_pack");
    RB_LAUNCH(cells_grid(ctx, cs), kSlThreads, 0, ctx->stream, k_cells_pack)(cs->dbg->dev, cs->cbf->dev, cs->cells, cs->C, cs->q);
    LAUNCH_CHECK();
    cs->in_cells = true;
    return RB_OK;
}
static int32_t filter_ensure_logical(rb_filter* f) { return f ? cells_ensure_logical(f->ctx, f->cs) : RB_OK; }
// a cell store for the pair (dbg, cbf) if the geometry allows it (cbf_bytes = C a power of two >= 32, dbg_bits = q * C with q <= 8) and
// the memory is there; RB_SLICED_CELLS=0 turns it off
static CellStore* cells_create(rb_ctx* ctx, rb_filter* dbg, rb_filter* cbf) {
    const char* v = getenv("RB_SLICED_CELLS");
    if (v && !strcmp(v, "0")) return nullptr;
    const int64_t C = cbf->size;
    if (C < 32 || (C & (C - 1)) != 0 || dbg->size % C != 0 || dbg->size / C > 8) return nullptr;
#ifdef RB_EMU
    if (C > (1LL << 26)) return nullptr;   // tests/emu runs one fiber per CUDA thread: converting GiB-sized filters would dominate the CPU suite
#endif
    CellStore* cs = new CellStore();
    cs->C = C; cs->q = (int)(dbg->size / C); cs->in_cells = false; cs->dbg = dbg; cs->cbf = cbf; cs->cells = nullptr;
    if (cudaMalloc(&cs->cells, (size_t)C * 2 + 256) != cudaSuccess) { cudaGetLastError(); delete cs; return nullptr; }
    (void)ctx;
    dbg->cs = cs; cbf->cs = cs;
    return cs;
}
static void cells_destroy(CellStore* cs) {
    if (!cs) return;
    if (cs->dbg) cs->dbg->cs = nullptr;
    if (cs->cbf) cs->cbf->cs = nullptr;
    cudaFree(cs->cells);
    delete cs;
}

// ---- context ---------------------------------------------------------------------------------------------------------
extern "C" int32_t rb_version(void) { return 120; }   // 120: raises ride the probe records' answer bytes, rb_graph_count_reads_async / rb_ctx_wait

extern "C" int32_t rb_ctx_create(int32_t device, rb_ctx** out) {
    if (!out) return fail(nullptr, RB_EINVAL, "rb_ctx_create: out is NULL");
    *out = nullptr;
    rb_ctx* ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, RB_ECUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count is 0"));
    if (device < 0 || device >= n) return fail(nullptr, RB_EINVAL, "rb_ctx_create: bad device index");
    CK(cudaSetDevice(device));
    rb_ctx* c = new rb_ctx();
    c->device = device;
    ctx = c;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (e != cudaSuccess || prop.major < 10) {
        std::string m = e != cudaSuccess ? cudaGetErrorString(e) : "device is not sm_100 (this library ships sm_100a code only)";
        delete c;
        return fail(nullptr, RB_ECUDA, m);
    }
    // Every filter probe is an isolated 32 B sector of a multi-GiB array: ask L2 to fetch exactly that from HBM instead of
    // the whole 128 B line (measured 4x DRAM over-fetch otherwise; profiles/r01_notes.md).  Process-wide device limit.
    if (!getenv("RB_KEEP_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->ev_kernel[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMalloc(&c->scratch, 64);
    if (e != cudaSuccess) { std::string m = cudaGetErrorString(e); delete c; return fail(nullptr, RB_ECUDA, m); }
    c->stream = c->own_stream;
    *out = c;
    return RB_OK;
}
extern "C" int32_t rb_ctx_destroy(rb_ctx* ctx) {
    if (!ctx) return RB_EINVAL;
    {
        LOCK(ctx);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < 32; ++i) if (ctx->stage[i]) cudaFree(ctx->stage[i]);
        for (int i = 0; i < 2; ++i) { if (ctx->ev_kernel[i]) cudaEventDestroy(ctx->ev_kernel[i]); if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]); }
        for (int i = 0; i < 8; ++i) if (ctx->ticket_ev[i]) cudaEventDestroy(ctx->ticket_ev[i]);
        if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
        if (ctx->claim) cudaFree(ctx->claim);
        if (ctx->scratch) cudaFree(ctx->scratch);
        if (ctx->ev0) { cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); }
        for (auto& sp : ctx->prof_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
        for (auto e : ctx->prof_pool) cudaEventDestroy(e);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    }
    delete ctx;
    return RB_OK;
}
extern "C" const char* rb_last_error(rb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
extern "C" int32_t rb_ctx_set_stream(rb_ctx* ctx, void* s) {
    if (!ctx) return RB_EINVAL;
    LOCK(ctx);
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return RB_OK;
}
extern "C" int32_t rb_ctx_sync(rb_ctx* ctx) {
    if (!ctx) return RB_EINVAL;
    LOCK(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}
extern "C" int32_t rb_ctx_set_rng_seed(rb_ctx* ctx, uint64_t seed) { if (!ctx) return RB_EINVAL; ctx->rng_seed = seed; return RB_OK; }
extern "C" int32_t rb_ctx_set_subbatch_kmers(rb_ctx* ctx, int64_t kmers) {
    if (!ctx || kmers < 1024) return RB_EINVAL;
    LOCK(ctx);
    ctx->subbatch_kmers = kmers;
    ctx->subbatch_user_set = true;
    return RB_OK;
}
extern "C" int64_t rb_ctx_kernel_launches(rb_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int32_t rb_timer_start(rb_ctx* ctx) {
    if (!ctx) return RB_EINVAL;
    LOCK(ctx);
    if (!ctx->ev0) { CK(cudaEventCreate(&ctx->ev0)); CK(cudaEventCreate(&ctx->ev1)); }
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    return RB_OK;
}
extern "C" int32_t rb_timer_stop(rb_ctx* ctx, float* ms) {
    if (!ctx || !ms || !ctx->ev0) return RB_EINVAL;
    LOCK(ctx);
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return RB_OK;
}
extern "C" int32_t rb_host_alloc(void** p, int64_t bytes) {
    if (!p || bytes < 0) return RB_EINVAL;
    return cudaHostAlloc(p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault) == cudaSuccess ? RB_OK : RB_ENOMEM;
}
extern "C" int32_t rb_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? RB_OK : RB_ECUDA; }
extern "C" int32_t rb_dev_alloc(rb_ctx* ctx, void** p, int64_t bytes) {
    if (!ctx || !p || bytes < 0) return RB_EINVAL;
    LOCK(ctx);
    CK(cudaMalloc(p, (size_t)std::max<int64_t>(bytes, 16)));
    return RB_OK;
}
extern "C" int32_t rb_dev_free(rb_ctx* ctx, void* p) {
    if (!ctx) return RB_EINVAL;
    LOCK(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(p));
    return RB_OK;
}
extern "C" int32_t rb_memcpy_h2d(rb_ctx* ctx, void* dst, const void* src, int64_t bytes) {
    if (!ctx) return RB_EINVAL;
    LOCK(ctx);
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}
extern "C" int32_t rb_memcpy_d2h(rb_ctx* ctx, void* dst, const void* src, int64_t bytes) {
    if (!ctx) return RB_EINVAL;
    LOCK(ctx);
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}

// grow-only device staging slot
static int32_t stage_get(rb_ctx* ctx, int slot, int64_t bytes, void** p) {
    bytes = std::max<int64_t>(bytes, 256);
    if (ctx->stage_bytes[slot] < bytes) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaStreamSynchronize(ctx->copy_stream));
        if (ctx->stage[slot]) CK(cudaFree(ctx->stage[slot]));
        ctx->stage[slot] = nullptr; ctx->stage_bytes[slot] = 0;
        const int64_t want = bytes + bytes / 4 + 256;
        CK(cudaMalloc(&ctx->stage[slot], (size_t)want));
        ctx->stage_bytes[slot] = want;
    }
    *p = ctx->stage[slot];
    return RB_OK;
}

// Make room in the claim table for `n` more claims (clears it when the load factor would pass 1/2).
static int32_t claim_reserve(rb_ctx* ctx, int64_t n, const void* owner, ClaimTable* ct) {
    if (owner != ctx->claim_owner) { ctx->claim_owner = owner; ctx->claim_used = ctx->claim_cap; }  // claims are per filter
    int64_t need = 1024;
    while (need < 2 * n) need <<= 1;
    if (ctx->claim_cap < need) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->claim) CK(cudaFree(ctx->claim));
        ctx->claim = nullptr; ctx->claim_cap = 0;
        CK(cudaMalloc(&ctx->claim, (size_t)(need + 1) * 8));
        ctx->claim_cap = need;
        ctx->claim_used = ctx->claim_cap;  // force the clear below
    }
    if (ctx->claim_used + n > ctx->claim_cap / 2) {
        CK(cudaMemsetAsync(ctx->claim, 0, (size_t)(ctx->claim_cap + 1) * 8, ctx->stream));
        ctx->claim_used = 0;
    }
    ctx->claim_used += n;
    ct->slots = ctx->claim;
    ct->mask = (uint64_t)ctx->claim_cap - 1;
    int lg = 0;
    while ((1LL << lg) < ctx->claim_cap) ++lg;
    ct->shift = 64 - lg;
    return RB_OK;
}
static void claim_invalidate(rb_ctx* ctx) { ctx->claim_used = ctx->claim_cap; }  // after empty()/upload(): stale claims must go

// ---- host helpers ------------------------------------------------------------------------------------------------------
extern "C" int64_t rb_expected_size(int64_t n, float fpr, int32_t num_hash) {  // BloomFilter.java:196-199
    const double r = (double)(-num_hash) / std::log(1 - std::exp(std::log((double)fpr) / (double)num_hash));
    return (int64_t)std::ceil((double)n * r);
}
extern "C" float rb_minifloat_to_float(int8_t b) { return minifloat_to_float((int)b); }

extern "C" int64_t rb_kmer_offsets(const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int32_t k, int64_t* off) {
    int64_t acc = 0;
    for (int64_t i = 0; i < n_reads; ++i) {
        if (off) off[i] = acc;
        const int32_t len = read_len ? read_len[i] : uniform_len;
        acc += std::max(0, len - k + 1);
    }
    if (off) off[n_reads] = acc;
    return acc;
}
extern "C" int64_t rb_pack_reads_host(const char* bases, const char* quals, const int64_t* ascii_off, int64_t n_reads, int32_t min_qual,
                                      uint64_t* packed, uint32_t* mask, int64_t* out_read_off, int32_t* out_read_len) {
    int64_t word = 0;
    const int qlo = '!' + min_qual;
    for (int64_t r = 0; r < n_reads; ++r) {
        const int64_t a0 = ascii_off[r];
        const int len = (int)(ascii_off[r + 1] - a0);
        out_read_off[r] = word * 32;
        out_read_len[r] = len;
        const int nw = (len + 31) / 32;
        for (int wi = 0; wi < nw; ++wi) {
            uint64_t w = 0; uint32_t m = 0;
            for (int j = 0; j < 32; ++j) {
                const int i = wi * 32 + j;
                if (i >= len) { m |= 1u << j; continue; }
                int code = 0; bool ok = true;
                switch ((unsigned char)bases[a0 + i]) {
                    case 'A': case 'a': code = 0; break;
                    case 'C': case 'c': code = 1; break;
                    case 'G': case 'g': code = 2; break;
                    case 'T': case 't': case 'U': case 'u': code = 3; break;
                    default: ok = false;
                }
                if (quals) { const unsigned char q = (unsigned char)quals[a0 + i]; if (q < qlo || q > '~') ok = false; }
                w |= (uint64_t)code << (2 * j);
                if (!ok) m |= 1u << j;
            }
            packed[word] = w;
            if (mask) mask[word] = m;
            ++word;
        }
    }
    return word * 32;
}

// ---- filters -------------------------------------------------------------------------------------------------------------
static int32_t filter_alloc(rb_ctx* ctx, int kind, int64_t size, int num_hash, int k, rb_filter** out) {
    if (!ctx || !out) return RB_EINVAL;
    if (size <= 0 || num_hash < 1 || num_hash > kMaxHash || k < 1) return fail(ctx, RB_EINVAL, "filter: size/num_hash/k out of range");
    rb_filter* f = new rb_filter();
    f->ctx = ctx; f->kind = kind; f->size = size; f->num_hash = num_hash; f->k = k; f->in_graph = false; f->cs = nullptr;
    f->nbytes = kind == RB_BLOOM ? size / 8 + ((size % 8) ? 1 : 0) : size;  // UnsafeBitBuffer.java:44-49
    f->alloc = div_up(f->nbytes + 16, 256) * 256;
    cudaError_t e = cudaMalloc(&f->dev, (size_t)f->alloc);
    if (e != cudaSuccess) { delete f; return fail(ctx, RB_ENOMEM, std::string("cudaMalloc(filter): ") + cudaGetErrorString(e)); }
    e = cudaMemsetAsync(f->dev, 0, (size_t)f->alloc, ctx->stream);
    ctx->claim_used = ctx->claim_cap;  // a fresh filter may reuse the address of a destroyed one: drop stale claims
    if (e != cudaSuccess) { cudaFree(f->dev); delete f; return fail(ctx, RB_ECUDA, cudaGetErrorString(e)); }
    *out = f;
    return RB_OK;
}
extern "C" int32_t rb_filter_create(rb_ctx* ctx, int32_t kind, int64_t size, int32_t num_hash, int32_t k, rb_filter** out) {
    if (!ctx || !out || (kind != RB_BLOOM && kind != RB_COUNTING)) return RB_EINVAL;
    LOCK(ctx);
    return filter_alloc(ctx, kind, size, num_hash, k, out);
}
static int32_t filter_free(rb_filter* f) {
    rb_ctx* ctx = f->ctx;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(f->dev));
    delete f;
    return RB_OK;
}
extern "C" int32_t rb_filter_destroy(rb_filter* f) {
    if (!f) return RB_EINVAL;
    if (f->in_graph) return fail(f->ctx, RB_ESTATE, "filter belongs to a graph; destroy the graph");
    LOCK(f->ctx);
    return filter_free(f);
}
extern "C" int32_t rb_filter_empty(rb_filter* f) {
    if (!f) return RB_EINVAL;
    rb_ctx* ctx = f->ctx;
    LOCK(ctx);
    { const int32_t rc = filter_ensure_logical(f); if (rc) return rc; }
    CK(cudaMemsetAsync(f->dev, 0, (size_t)f->alloc, ctx->stream));
    claim_invalidate(ctx);
    return RB_OK;
}
extern "C" int64_t rb_filter_size(const rb_filter* f) { return f ? f->size : 0; }
extern "C" int64_t rb_filter_num_bytes(const rb_filter* f) { return f ? f->nbytes : 0; }
extern "C" int32_t rb_filter_num_hash(const rb_filter* f) { return f ? f->num_hash : 0; }
// the logical array; for a graph's filter it is current until the next read-level call on the graph
extern "C" int32_t rb_filter_device_ptr(rb_filter* f, void** p) {
    if (!f || !p) return RB_EINVAL;
    LOCK(f->ctx);
    const int32_t rc = filter_ensure_logical(f);
    if (rc) return rc;
    *p = f->dev;
    return RB_OK;
}

static GraphDev filter_view(rb_filter* f) {  // a lone filter seen through the graph-shaped kernel argument
    GraphDev gd;
    memset(&gd, 0, sizeof gd);
    gd.hm = make_hm(f->k);
    gd.k = f->k;
    gd.rng_seed = f->ctx->rng_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(f->ctx->launches + 1);  // fresh coins every launch
    gd.dbg.words = f->dev; gd.dbg.fm = make_fm(f->size); gd.dbg.num_hash = f->num_hash;
    gd.cbf.words = f->dev; gd.cbf.fm = make_fm(f->size); gd.cbf.num_hash = f->num_hash;
    return gd;
}

template <int OP>
static void launch_hash_op(int maxh, int64_t n, cudaStream_t s, const int64_t* base, const GraphDev& gd, uint8_t* o8, float* of) {
    const int grid = (int)div_up(n, kThreads);
    if (maxh <= 2) RB_LAUNCH(grid, kThreads, 0, s, k_hash_op<2, OP>)(base, n, gd, o8, of);
    else if (maxh <= 3) RB_LAUNCH(grid, kThreads, 0, s, k_hash_op<3, OP>)(base, n, gd, o8, of);
    else if (maxh <= 4) RB_LAUNCH(grid, kThreads, 0, s, k_hash_op<4, OP>)(base, n, gd, o8, of);
    else RB_LAUNCH(grid, kThreads, 0, s, k_hash_op<8, OP>)(base, n, gd, o8, of);
}
static void dispatch_hash_op(int op, int maxh, int64_t n, cudaStream_t s, const int64_t* base, const GraphDev& gd, uint8_t* o8, float* of) {
    switch (op) {
        case OP_BF_ADD: launch_hash_op<OP_BF_ADD>(maxh, n, s, base, gd, o8, of); break;
        case OP_BF_LOOKUP: launch_hash_op<OP_BF_LOOKUP>(maxh, n, s, base, gd, o8, of); break;
        case OP_BF_LTA: launch_hash_op<OP_BF_LTA>(maxh, n, s, base, gd, o8, of); break;
        case OP_CBF_INC: launch_hash_op<OP_CBF_INC>(maxh, n, s, base, gd, o8, of); break;
        case OP_CBF_INC_GET: launch_hash_op<OP_CBF_INC_GET>(maxh, n, s, base, gd, o8, of); break;
        case OP_CBF_COUNT: launch_hash_op<OP_CBF_COUNT>(maxh, n, s, base, gd, o8, of); break;
        case OP_GRAPH_ADD: launch_hash_op<OP_GRAPH_ADD>(maxh, n, s, base, gd, o8, of); break;
        case OP_GRAPH_COUNT_IF_PRESENT: launch_hash_op<OP_GRAPH_COUNT_IF_PRESENT>(maxh, n, s, base, gd, o8, of); break;
        default: launch_hash_op<OP_GRAPH_COUNT>(maxh, n, s, base, gd, o8, of); break;
    }
}

// host-pointer per-hash operator: H2D hashes, run, D2H results, in sub-batches
static int32_t run_hash_op(rb_ctx* ctx, int op, GraphDev gd, int maxh, const int64_t* base, int64_t n, uint8_t* out8, float* outf) {
    if (n < 0 || (n > 0 && !base)) return fail(ctx, RB_EINVAL, "hash op: bad arguments");
    const bool needs_claim = (op == OP_BF_LTA || op == OP_GRAPH_ADD);
    const int64_t step = ctx->subbatch_kmers;
    for (int64_t i0 = 0; i0 < n; i0 += step) {
        const int64_t m = std::min(step, n - i0);
        void *dbase, *dout;
        int32_t rc = stage_get(ctx, 0, m * 8, &dbase); if (rc) return rc;
        rc = stage_get(ctx, 1, m * 4, &dout); if (rc) return rc;
        if (needs_claim) { rc = claim_reserve(ctx, m, gd.dbg.words, &gd.ct); if (rc) return rc; }
        CK(cudaMemcpyAsync(dbase, base + i0, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        dispatch_hash_op(op, maxh, m, ctx->stream, (const int64_t*)dbase, gd, (uint8_t*)dout, (float*)dout);
        LAUNCH_CHECK();
        if (out8) CK(cudaMemcpyAsync(out8 + i0, dout, (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
        if (outf) CK(cudaMemcpyAsync(outf + i0, dout, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return RB_OK;
}
#define FILTER_OP(f, want_kind, op, o8, of)                                                      \
    if (!(f)) return RB_EINVAL;                                                                  \
    rb_ctx* ctx = (f)->ctx;                                                                      \
    LOCK(ctx);                                                                                   \
    if ((f)->kind != (want_kind)) return fail(ctx, RB_EINVAL, "wrong filter kind for this call"); \
    { const int32_t rc_ = filter_ensure_logical(f); if (rc_) return rc_; }                       \
    return run_hash_op(ctx, (op), filter_view(f), (f)->num_hash, base, n, (o8), (of))

extern "C" int32_t rb_filter_add_hashes(rb_filter* f, const int64_t* base, int64_t n) { FILTER_OP(f, RB_BLOOM, OP_BF_ADD, nullptr, nullptr); }
extern "C" int32_t rb_filter_lookup_hashes(rb_filter* f, const int64_t* base, int64_t n, uint8_t* out) { FILTER_OP(f, RB_BLOOM, OP_BF_LOOKUP, out, nullptr); }
extern "C" int32_t rb_filter_lookup_then_add_hashes(rb_filter* f, const int64_t* base, int64_t n, uint8_t* out) { FILTER_OP(f, RB_BLOOM, OP_BF_LTA, out, nullptr); }
extern "C" int32_t rb_cbf_increment_hashes(rb_filter* f, const int64_t* base, int64_t n) { FILTER_OP(f, RB_COUNTING, OP_CBF_INC, nullptr, nullptr); }
extern "C" int32_t rb_cbf_increment_and_get_hashes(rb_filter* f, const int64_t* base, int64_t n, float* out) { FILTER_OP(f, RB_COUNTING, OP_CBF_INC_GET, nullptr, out); }
extern "C" int32_t rb_cbf_count_hashes(rb_filter* f, const int64_t* base, int64_t n, float* out) { FILTER_OP(f, RB_COUNTING, OP_CBF_COUNT, nullptr, out); }

extern "C" int32_t rb_filter_popcount(rb_filter* f, int64_t* out) {
    if (!f || !out) return RB_EINVAL;
    rb_ctx* ctx = f->ctx;
    LOCK(ctx);
    { const int32_t rc = filter_ensure_logical(f); if (rc) return rc; }
    CK(cudaMemsetAsync(ctx->scratch, 0, 8, ctx->stream));
    const int64_t n_vec = f->alloc / 16;  // the padding beyond nbytes is always zero
    const int grid = (int)std::min<int64_t>(div_up(n_vec, kThreads), (int64_t)ctx->sm_count * 16);
    if (f->kind == RB_BLOOM) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_popcount<0>)((const uint4*)f->dev, n_vec, ctx->scratch);
    else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_popcount<1>)((const uint4*)f->dev, n_vec, ctx->scratch);
    LAUNCH_CHECK();
    unsigned long long v = 0;
    CK(cudaMemcpyAsync(&v, ctx->scratch, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out = (int64_t)v;
    return RB_OK;
}
extern "C" int32_t rb_filter_fpr(rb_filter* f, float* out) {  // BloomFilter.java:185-194
    if (!f || !out) return RB_EINVAL;
    int64_t pop = 0;
    const int32_t rc = rb_filter_popcount(f, &pop);
    if (rc) return rc;
    *out = (float)std::pow((double)pop / (double)f->size, f->num_hash);
    return RB_OK;
}
extern "C" int32_t rb_filter_download(rb_filter* f, void* dst, int64_t nbytes) {
    if (!f || !dst) return RB_EINVAL;
    rb_ctx* ctx = f->ctx;
    LOCK(ctx);
    if (nbytes != f->nbytes) return fail(ctx, RB_EINVAL, "download: nbytes must equal rb_filter_num_bytes()");
    { const int32_t rc = filter_ensure_logical(f); if (rc) return rc; }
    CK(cudaMemcpyAsync(dst, f->dev, (size_t)nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}
extern "C" int32_t rb_filter_upload(rb_filter* f, const void* src, int64_t nbytes) {
    if (!f || !src) return RB_EINVAL;
    rb_ctx* ctx = f->ctx;
    LOCK(ctx);
    if (nbytes != f->nbytes) return fail(ctx, RB_EINVAL, "upload: nbytes must equal rb_filter_num_bytes()");
    { const int32_t rc = filter_ensure_logical(f); if (rc) return rc; }
    CK(cudaMemcpyAsync(f->dev, src, (size_t)nbytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    claim_invalidate(ctx);
    return RB_OK;
}

// ---- CascadingBloomFilter (bloom/CascadingBloomFilter.java:34-100): L plain Bloom filters of size / L bits each ---------------------------------
// add: walk the levels with lookupThenAdd, stop at the first level where the key was not already present; lookup: the top level;
// lookupThenAdd: true iff every level already had the key.  A batch goes level by level: the instances a level reports as present
// (duplicates inside the batch included: the claim table makes exactly one instance of a key the first sighting per level) are
// compacted and handed to the next level -- the same final state as the sequential loop for any order of the batch.
struct rb_cascade { rb_ctx* ctx; int levels; std::vector<rb_filter*> bf; };
extern "C" int32_t rb_cascade_create(rb_ctx* ctx, int64_t size, int32_t num_hash, int32_t k, int32_t num_levels, rb_cascade** out) {
    if (!ctx || !out || num_levels < 1 || num_levels > 64) return RB_EINVAL;
    LOCK(ctx);
    rb_cascade* c = new rb_cascade();
    c->ctx = ctx; c->levels = num_levels;
    for (int i = 0; i < num_levels; ++i) {
        rb_filter* f = nullptr;
        const int32_t rc = filter_alloc(ctx, RB_BLOOM, size / num_levels, num_hash, k, &f);   // partitionSize = size / numLevels (:38)
        if (rc) { for (rb_filter* x : c->bf) filter_free(x); delete c; return rc; }
        f->in_graph = true;
        c->bf.push_back(f);
    }
    *out = c;
    return RB_OK;
}
extern "C" int32_t rb_cascade_destroy(rb_cascade* c) {
    if (!c) return RB_EINVAL;
    { LOCK(c->ctx); for (rb_filter* f : c->bf) filter_free(f); }
    delete c;
    return RB_OK;
}
extern "C" int32_t rb_cascade_level(rb_cascade* c, int32_t level, rb_filter** out) {
    if (!c || !out || level < 0 || level >= c->levels) return RB_EINVAL;
    *out = c->bf[(size_t)level];
    return RB_OK;
}
extern "C" int32_t rb_cascade_lookup_hashes(rb_cascade* c, const int64_t* base, int64_t n, uint8_t* out) {   // :79-85: the top level answers
    if (!c) return RB_EINVAL;
    return rb_filter_lookup_hashes(c->bf.back(), base, n, out);
}
// out (nullable): lookupThenAdd's answer per key (:93-100); add (:66-72) is the same walk without the answer
extern "C" int32_t rb_cascade_lookup_then_add_hashes(rb_cascade* c, const int64_t* base, int64_t n, uint8_t* out) {
    if (!c || n < 0 || (n > 0 && !base)) return RB_EINVAL;
    rb_ctx* ctx = c->ctx;
    LOCK(ctx);
    if (n > INT32_MAX) return fail(ctx, RB_EINVAL, "cascade: at most 2^31-1 keys per call");
    if (n == 0) return RB_OK;
    void *k0, *k1, *i0, *i1, *fl, *res = nullptr;
    int32_t rc;
    if ((rc = stage_get(ctx, 0, n * 8, &k0)) || (rc = stage_get(ctx, 1, n * 8, &k1)) || (rc = stage_get(ctx, 2, n * 4, &i0)) ||
        (rc = stage_get(ctx, 3, n * 4, &i1)) || (rc = stage_get(ctx, 4, n, &fl))) return rc;
    if (out && (rc = stage_get(ctx, 5, n, &res))) return rc;
    CK(cudaMemcpyAsync(k0, base, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (out) CK(cudaMemsetAsync(res, 0, (size_t)n, ctx->stream));
    int64_t m = n;
    int64_t *kin = (int64_t*)k0, *kout = (int64_t*)k1;
    int32_t *iin = nullptr, *iout = (int32_t*)i0, *ispare = (int32_t*)i1;
    for (int lv = 0; lv < c->levels && m > 0; ++lv) {
        GraphDev gd = filter_view(c->bf[(size_t)lv]);
        rc = claim_reserve(ctx, m, gd.dbg.words, &gd.ct);
        if (rc) return rc;
        dispatch_hash_op(OP_BF_LTA, c->bf[(size_t)lv]->num_hash, m, ctx->stream, kin, gd, (uint8_t*)fl, nullptr);
        LAUNCH_CHECK();
        CK(cudaMemsetAsync(ctx->scratch, 0, 8, ctx->stream));
        RB_LAUNCH((int)div_up(m, kThreads), kThreads, 0, ctx->stream, k_cascade_survivors)(kin, iin, (const uint8_t*)fl, m, kout, iout, (unsigned int*)ctx->scratch);
        LAUNCH_CHECK();
        unsigned int next = 0;
        CK(cudaMemcpyAsync(&next, ctx->scratch, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        m = next;
        std::swap(kin, kout);
        int32_t* t = iin ? iin : ispare;
        iin = iout; iout = t;
    }
    if (out) {   // the survivors of the last level were present everywhere
        if (m > 0) { RB_LAUNCH((int)div_up(m, kThreads), kThreads, 0, ctx->stream, k_scatter_ones)(iin, m, (uint8_t*)res); LAUNCH_CHECK(); }
        CK(cudaMemcpyAsync(out, res, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return RB_OK;
}
extern "C" int32_t rb_cascade_add_hashes(rb_cascade* c, const int64_t* base, int64_t n) { return rb_cascade_lookup_then_add_hashes(c, base, n, nullptr); }

// Float.toString look-alike for the "fpr:" line of the .desc files (value is informational: the reader skips it,
// bloom/BloomFilter.java:73-88)
static std::string java_float(float v) {
    if (v == 0.f) return "0.0";
    char buf[64];
    int prec = 1;
    for (; prec <= 9; ++prec) { snprintf(buf, sizeof buf, "%.*e", prec - 1, (double)v); if (strtof(buf, nullptr) == v) break; }
    snprintf(buf, sizeof buf, "%.*e", prec - 1, (double)v);
    std::string s(buf);
    const size_t epos = s.find('e');
    std::string mant = s.substr(0, epos);
    const int ex = atoi(s.c_str() + epos + 1);
    std::string digits;
    for (char c : mant) if (c >= '0' && c <= '9') digits.push_back(c);
    const bool neg = v < 0;
    std::string out;
    if (ex >= -3 && ex < 7) {
        if (ex >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)ex + 1));
            while ((int)ip.size() < ex + 1) ip.push_back('0');
            std::string fp = digits.size() > (size_t)ex + 1 ? digits.substr(ex + 1) : "0";
            out = ip + "." + fp;
        } else {
            out = "0." + std::string((size_t)(-ex - 1), '0') + digits;
        }
    } else {
        out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(ex);
    }
    return neg ? "-" + out : out;
}

static int32_t write_file(rb_ctx* ctx, const char* path, const void* data, size_t n) {
    FILE* fp = fopen(path, "wb");
    if (!fp) return fail(ctx, RB_EIO, std::string("cannot open for writing: ") + path);
    const size_t w = n ? fwrite(data, 1, n, fp) : 0;
    if (fclose(fp) != 0 || w != n) return fail(ctx, RB_EIO, std::string("short write: ") + path);
    return RB_OK;
}
extern "C" int32_t rb_filter_save(rb_filter* f, const char* desc_path, const char* bits_path) {  // BloomFilter.java:113-124
    if (!f || !desc_path || !bits_path) return RB_EINVAL;
    rb_ctx* ctx = f->ctx;
    LOCK(ctx);
    float fpr = 0;
    int32_t rc = rb_filter_fpr(f, &fpr);
    if (rc) return rc;
    const std::string desc = "size:" + std::to_string(f->size) + "\nnumhash:" + std::to_string(f->num_hash) + "\nfpr:" + java_float(fpr) + "\n";
    rc = write_file(ctx, desc_path, desc.data(), desc.size());
    if (rc) return rc;
    std::vector<uint8_t> host((size_t)f->nbytes);
    rc = rb_filter_download(f, host.data(), f->nbytes);
    if (rc) return rc;
    return write_file(ctx, bits_path, host.data(), host.size());
}
static int32_t read_desc(rb_ctx* ctx, const char* path, std::vector<std::pair<std::string, std::string>>* kv) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return fail(ctx, RB_EIO, std::string("cannot open: ") + path);
    char line[512];
    while (fgets(line, sizeof line, fp)) {
        std::string s(line);
        while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back();
        const size_t c = s.find(':');
        if (c == std::string::npos) continue;
        kv->emplace_back(s.substr(0, c), s.substr(c + 1));
    }
    fclose(fp);
    return RB_OK;
}
extern "C" int32_t rb_filter_load(rb_ctx* ctx, int32_t kind, const char* desc_path, const char* bits_path, int32_t k, int32_t load_bits,
                                  rb_filter** out) {  // BloomFilter.java:70-106
    if (!ctx || !desc_path || !out) return RB_EINVAL;
    LOCK(ctx);
    std::vector<std::pair<std::string, std::string>> kv;
    int32_t rc = read_desc(ctx, desc_path, &kv);
    if (rc) return rc;
    int64_t size = 0; int num_hash = 0;
    for (auto& e : kv) { if (e.first == "size") size = atoll(e.second.c_str()); else if (e.first == "numhash") num_hash = atoi(e.second.c_str()); }
    rb_filter* f = nullptr;
    rc = filter_alloc(ctx, kind, size, num_hash, k, &f);
    if (rc) return rc;
    if (load_bits) {
        if (!bits_path) { filter_free(f); return RB_EINVAL; }
        FILE* fp = fopen(bits_path, "rb");
        if (!fp) { filter_free(f); return fail(ctx, RB_EIO, std::string("cannot open: ") + bits_path); }
        std::vector<uint8_t> host((size_t)f->nbytes);
        const size_t got = fread(host.data(), 1, host.size(), fp);
        fclose(fp);
        if (got != host.size()) { filter_free(f); return fail(ctx, RB_EIO, std::string("file shorter than the filter: ") + bits_path); }
        rc = rb_filter_upload(f, host.data(), f->nbytes);
        if (rc) { filter_free(f); return rc; }
    }
    *out = f;
    return RB_OK;
}

extern "C" int32_t rb_index_hashes(rb_ctx* ctx, const int64_t* hash, int64_t n, int64_t size, int64_t* out) {
    if (!ctx || n < 0 || size <= 0 || (n > 0 && (!hash || !out))) return RB_EINVAL;
    LOCK(ctx);
    if (n == 0) return RB_OK;
    void *din, *dout;
    int32_t rc = stage_get(ctx, 0, n * 8, &din); if (rc) return rc;
    rc = stage_get(ctx, 1, n * 8, &dout); if (rc) return rc;
    CK(cudaMemcpyAsync(din, hash, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    RB_LAUNCH((int)div_up(n, kThreads), kThreads, 0, ctx->stream, k_index)((const int64_t*)din, n, make_fm(size), (int64_t*)dout);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(out, dout, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}

// ---- ingest plumbing ---------------------------------------------------------------------------------------------------
struct ReadsArg {
    const uint64_t* packed; const uint32_t* mask; const int64_t* read_off; const int32_t* read_len;
    int64_t n_reads; int32_t uniform_len; int64_t uniform_stride;
    bool on_device;
    int64_t read0 = 0;   // uniform layout: index of the first read inside `packed` (a launch handed on as a ReadsArg of its own)
    const uint32_t* rcm = nullptr;   // device reads only: reverse-strand seed plane of unusable bases (Ingest::rcm; the ASCII entry points make it)
};
// One launch worth of reads, all pointers on the device.
struct Launch { Ingest ing; int64_t n_pos; };

typedef int32_t (*LaunchFn)(rb_ctx* ctx, const Ingest& ing, void* user);

// Splits the reads into launches of at most ctx->subbatch_kmers positions (span = bases a position needs: k or k+d),
// stages host data when needed, and calls fn for each launch.  total_out receives the number of positions.
static int32_t for_each_launch(rb_ctx* ctx, const ReadsArg& ra, int span, LaunchFn fn, void* user, int64_t* total_out) {
    if (ra.n_reads < 0 || (ra.n_reads > 0 && !ra.packed)) return fail(ctx, RB_EINVAL, "reads: packed is NULL");
    const bool uniform = ra.read_off == nullptr;
    if (uniform && (ra.uniform_len <= 0 || ra.uniform_stride < ra.uniform_len)) return fail(ctx, RB_EINVAL, "reads: bad uniform layout");
    if (!uniform && !ra.read_len) return fail(ctx, RB_EINVAL, "reads: read_len is NULL");
    int64_t total = 0;
    if (uniform) {
        const int32_t npos = std::max(0, ra.uniform_len - span + 1);
        total = npos * ra.n_reads;
        if (npos > 0) {
            const int64_t reads_per = std::max<int64_t>(1, ctx->subbatch_kmers / npos);
            for (int64_t r0 = 0; r0 < ra.n_reads; r0 += reads_per) {
                const int64_t nr = std::min(reads_per, ra.n_reads - r0);
                Ingest ing;
                memset(&ing, 0, sizeof ing);
                ing.n_reads = nr; ing.n_pos = nr * npos; ing.out_base = r0 * npos; ing.read_base = r0;
                ing.uniform_stride = ra.uniform_stride; ing.uniform_len = ra.uniform_len; ing.uniform_npos = npos;
                const int64_t b_lo = (ra.read0 + r0) * ra.uniform_stride, b_hi = (ra.read0 + r0 + nr - 1) * ra.uniform_stride + ra.uniform_len;
                if (ra.on_device) { ing.packed = ra.packed; ing.mask = ra.mask; ing.rcm = ra.rcm; ing.first_base = b_lo; }
                else {
                    const int64_t w_lo = b_lo >> 5, w_hi = (b_hi + 31) >> 5;
                    void* dp; int32_t rc = stage_get(ctx, 0, (w_hi - w_lo) * 8, &dp); if (rc) return rc;
                    CK(cudaMemcpyAsync(dp, ra.packed + w_lo, (size_t)(w_hi - w_lo) * 8, cudaMemcpyHostToDevice, ctx->stream));
                    ing.packed = (const uint64_t*)dp - w_lo;
                    if (ra.mask) {
                        void* dm; rc = stage_get(ctx, 1, (w_hi - w_lo) * 4, &dm); if (rc) return rc;
                        CK(cudaMemcpyAsync(dm, ra.mask + w_lo, (size_t)(w_hi - w_lo) * 4, cudaMemcpyHostToDevice, ctx->stream));
                        ing.mask = (const uint32_t*)dm - w_lo;
                    }
                    ing.first_base = b_lo;
                }
                const int32_t rc = fn(ctx, ing, user);   // staging reuse is safe: copies and kernels are ordered on one stream
                if (rc) return rc;
            }
        }
    } else {
        // per-read tables are needed on the host to cut launches
        std::vector<int64_t> h_off; std::vector<int32_t> h_len;
        const int64_t* off = ra.read_off; const int32_t* len = ra.read_len;
        if (ra.on_device) {
            h_off.resize((size_t)ra.n_reads); h_len.resize((size_t)ra.n_reads);
            CK(cudaMemcpyAsync(h_off.data(), ra.read_off, (size_t)ra.n_reads * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(h_len.data(), ra.read_len, (size_t)ra.n_reads * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            off = h_off.data(); len = h_len.data();
        }
        std::vector<int64_t> pos_off((size_t)ra.n_reads + 1);
        for (int64_t i = 0; i < ra.n_reads; ++i) { pos_off[(size_t)i] = total; total += std::max(0, len[i] - span + 1); }
        pos_off[(size_t)ra.n_reads] = total;
        int64_t r0 = 0;
        while (r0 < ra.n_reads) {
            int64_t r1 = r0 + 1;
            while (r1 < ra.n_reads && pos_off[(size_t)r1 + 1] - pos_off[(size_t)r0] <= ctx->subbatch_kmers) ++r1;
            const int64_t nr = r1 - r0, npos = pos_off[(size_t)r1] - pos_off[(size_t)r0];
            if (npos > 0) {
                Ingest ing;
                memset(&ing, 0, sizeof ing);
                ing.n_reads = nr; ing.n_pos = npos; ing.out_base = pos_off[(size_t)r0]; ing.pos_bias = pos_off[(size_t)r0]; ing.read_base = r0;
                void *d_po, *d_ro = nullptr, *d_rl = nullptr;
                int32_t rc = stage_get(ctx, 2, (nr + 1) * 8, &d_po); if (rc) return rc;
                CK(cudaMemcpyAsync(d_po, pos_off.data() + r0, (size_t)(nr + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
                ing.pos_off = (const int64_t*)d_po;
                if (ra.on_device) { ing.packed = ra.packed; ing.mask = ra.mask; ing.rcm = ra.rcm; ing.read_off = ra.read_off + r0; ing.read_len = ra.read_len + r0; }
                else {
                    int64_t b_lo = INT64_MAX, b_hi = 0;
                    for (int64_t i = r0; i < r1; ++i) if (len[i] > 0) { b_lo = std::min(b_lo, off[i]); b_hi = std::max(b_hi, off[i] + len[i]); }
                    if (b_lo < 0) return fail(ctx, RB_EINVAL, "reads: negative read_off");
                    const int64_t w_lo = b_lo >> 5, w_hi = (b_hi + 31) >> 5;
                    void* dp; rc = stage_get(ctx, 0, (w_hi - w_lo) * 8, &dp); if (rc) return rc;
                    CK(cudaMemcpyAsync(dp, ra.packed + w_lo, (size_t)(w_hi - w_lo) * 8, cudaMemcpyHostToDevice, ctx->stream));
                    ing.packed = (const uint64_t*)dp - w_lo;
                    if (ra.mask) {
                        void* dm; rc = stage_get(ctx, 1, (w_hi - w_lo) * 4, &dm); if (rc) return rc;
                        CK(cudaMemcpyAsync(dm, ra.mask + w_lo, (size_t)(w_hi - w_lo) * 4, cudaMemcpyHostToDevice, ctx->stream));
                        ing.mask = (const uint32_t*)dm - w_lo;
                    }
                    rc = stage_get(ctx, 3, nr * 8, &d_ro); if (rc) return rc;
                    rc = stage_get(ctx, 4, nr * 4, &d_rl); if (rc) return rc;
                    CK(cudaMemcpyAsync(d_ro, off + r0, (size_t)nr * 8, cudaMemcpyHostToDevice, ctx->stream));
                    CK(cudaMemcpyAsync(d_rl, len + r0, (size_t)nr * 4, cudaMemcpyHostToDevice, ctx->stream));
                    ing.read_off = (const int64_t*)d_ro; ing.read_len = (const int32_t*)d_rl;
                }
                rc = fn(ctx, ing, user);
                if (rc) return rc;
            }
            r0 = r1;
        }
    }
    if (total_out) *total_out = total;
    return RB_OK;
}

// ---- k-merizer ---------------------------------------------------------------------------------------------------------
struct KmerizeUser { int k, mode, d; int64_t *f, *r, *b; int64_t *hf, *hr, *hb; bool pairs; };
static int32_t kmerize_launch(rb_ctx* ctx, const Ingest& ing, void* user) {
    KmerizeUser* u = (KmerizeUser*)user;
    const int grid = (int)div_up(div_up(ing.n_pos, kChunk), kThreads);
    // outputs: device staging slots 5..7 indexed by launch-local position
    Ingest g = ing;
    const int64_t out0 = ing.out_base;
    g.out_base = 0;
    int64_t *df = nullptr, *dr = nullptr, *db = nullptr;
    int32_t rc;
    if (u->hf) { void* p; rc = stage_get(ctx, 5, ing.n_pos * 8, &p); if (rc) return rc; df = (int64_t*)p; }
    if (u->hr) { void* p; rc = stage_get(ctx, 6, ing.n_pos * 8, &p); if (rc) return rc; dr = (int64_t*)p; }
    if (u->hb) { void* p; rc = stage_get(ctx, 7, ing.n_pos * 8, &p); if (rc) return rc; db = (int64_t*)p; }
    if (u->pairs) {
        GraphDev gd; memset(&gd, 0, sizeof gd); gd.k = u->k; gd.hm = make_hm(u->k);
        BitFilter none; memset(&none, 0, sizeof none);
        if (u->mode == RB_MODE_FWD) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<0, 2, 0>)(g, gd, none, u->d, db);
        else if (u->mode == RB_MODE_RC) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<1, 2, 0>)(g, gd, none, u->d, db);
        else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<2, 2, 0>)(g, gd, none, u->d, db);
    } else {
        if (u->mode == RB_MODE_FWD) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_kmerize<0>)(g, u->k, df, dr, db);
        else if (u->mode == RB_MODE_RC) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_kmerize<1>)(g, u->k, df, dr, db);
        else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_kmerize<2>)(g, u->k, df, dr, db);
    }
    LAUNCH_CHECK();
    if (u->hf) CK(cudaMemcpyAsync(u->hf + out0, df, (size_t)ing.n_pos * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (u->hr) CK(cudaMemcpyAsync(u->hr + out0, dr, (size_t)ing.n_pos * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (u->hb) CK(cudaMemcpyAsync(u->hb + out0, db, (size_t)ing.n_pos * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}
extern "C" int32_t rb_kmerize(rb_ctx* ctx, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                              int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int32_t k, int32_t mode, int64_t* fhash,
                              int64_t* rhash, int64_t* base) {
    if (!ctx || k < 1 || mode < 0 || mode > 2) return RB_EINVAL;
    LOCK(ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    KmerizeUser u{k, mode, 0, nullptr, nullptr, nullptr, fhash, rhash, base, false};
    return for_each_launch(ctx, ra, k, kmerize_launch, &u, nullptr);
}
extern "C" int32_t rb_kmerize_pairs(rb_ctx* ctx, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                    int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int32_t k, int32_t d, int32_t mode,
                                    int64_t* pair_base) {
    if (!ctx || k < 1 || d < 1 || mode < 0 || mode > 2 || !pair_base) return RB_EINVAL;
    LOCK(ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    KmerizeUser u{k, mode, d, nullptr, nullptr, nullptr, nullptr, nullptr, pair_base, true};
    return for_each_launch(ctx, ra, k + d, kmerize_launch, &u, nullptr);
}

// MinimizerHashIterator.next() for every window of w k-mers of every read (bloom/hash/MinimizerHashIterator.java:42-101): out has
// max(0, len - k - w + 2) values per read, in read order (rb_kmer_offsets with k + w - 1)
struct MinimizerUser { int k, w, mode; int64_t* out; };
static int32_t minimizer_launch(rb_ctx* ctx, const Ingest& ing, void* user) {
    MinimizerUser* u = (MinimizerUser*)user;
    const int grid = (int)div_up(div_up(ing.n_pos, kChunk), kThreads);
    Ingest g = ing;
    const int64_t out0 = ing.out_base;
    g.out_base = 0;
    void* p;
    const int32_t rc = stage_get(ctx, 5, ing.n_pos * 8, &p);
    if (rc) return rc;
    PROF("k_minimizers");
    if (u->mode == RB_MODE_FWD) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_minimizers<0>)(g, u->k, u->w, (int64_t*)p);
    else if (u->mode == RB_MODE_RC) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_minimizers<1>)(g, u->k, u->w, (int64_t*)p);
    else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_minimizers<2>)(g, u->k, u->w, (int64_t*)p);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(u->out + out0, p, (size_t)ing.n_pos * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}
extern "C" int32_t rb_minimizers(rb_ctx* ctx, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                 int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int32_t k, int32_t w, int32_t mode, int64_t* minimizers) {
    if (!ctx || k < 1 || w < 1 || w > kMaxMinimizerWindow || mode < 0 || mode > 2 || !minimizers) return RB_EINVAL;
    LOCK(ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    MinimizerUser u{k, w, mode, minimizers};
    return for_each_launch(ctx, ra, k + w - 1, minimizer_launch, &u, nullptr);
}

// ---- graph ---------------------------------------------------------------------------------------------------------------
// ---- f4: the screening Bloom filter over whole sequences (RNABloom.java:1680,2530-2536,4264; util/GraphUtils.java:627-650) ----------------
struct SeqUser { rb_filter* f; int mode, op; uint8_t* missing; };
template <int MODE, int MAXH>
static void launch_seq_op(int op, int grid, cudaStream_t s, const Ingest& ing, int k, const BitFilter& bf, const HashMults& hm, uint8_t* missing) {
    if (op == SEQ_ADD) RB_LAUNCH(grid, kThreads, 0, s, k_seq_filter<MODE, MAXH, SEQ_ADD>)(ing, k, bf, hm, missing);
    else if (op == SEQ_CONTAINS_ALL) RB_LAUNCH(grid, kThreads, 0, s, k_seq_filter<MODE, MAXH, SEQ_CONTAINS_ALL>)(ing, k, bf, hm, missing);
    else RB_LAUNCH(grid, kThreads, 0, s, k_seq_filter<MODE, MAXH, SEQ_LOOKUP_AND_ADD_ALL>)(ing, k, bf, hm, missing);
}
template <int MAXH>
static void launch_seq_mode(int mode, int op, int grid, cudaStream_t s, const Ingest& ing, int k, const BitFilter& bf, const HashMults& hm, uint8_t* missing) {
    if (mode == RB_MODE_FWD) launch_seq_op<0, MAXH>(op, grid, s, ing, k, bf, hm, missing);
    else if (mode == RB_MODE_RC) launch_seq_op<1, MAXH>(op, grid, s, ing, k, bf, hm, missing);
    else launch_seq_op<2, MAXH>(op, grid, s, ing, k, bf, hm, missing);
}
static int32_t seq_launch(rb_ctx* ctx, const Ingest& ing, void* user) {
    SeqUser* u = (SeqUser*)user;
    BitFilter bf; bf.words = u->f->dev; bf.fm = make_fm(u->f->size); bf.num_hash = u->f->num_hash;
    const HashMults hm = make_hm(u->f->k);
    const int grid = (int)div_up(div_up(ing.n_pos, kChunk), kThreads);
    PROF("k_seq_filter");
    if (u->f->num_hash <= 2) launch_seq_mode<2>(u->mode, u->op, grid, ctx->stream, ing, u->f->k, bf, hm, u->missing);
    else if (u->f->num_hash <= 3) launch_seq_mode<3>(u->mode, u->op, grid, ctx->stream, ing, u->f->k, bf, hm, u->missing);
    else if (u->f->num_hash <= 4) launch_seq_mode<4>(u->mode, u->op, grid, ctx->stream, ing, u->f->k, bf, hm, u->missing);
    else launch_seq_mode<8>(u->mode, u->op, grid, ctx->stream, ing, u->f->k, bf, hm, u->missing);
    LAUNCH_CHECK();
    return RB_OK;
}
extern "C" int32_t rb_filter_seq_op(rb_filter* f, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                    int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int32_t mode, int32_t op, uint8_t* all_found) {
    if (!f || n_reads < 0 || mode < 0 || mode > 2 || op < SEQ_ADD || op > SEQ_LOOKUP_AND_ADD_ALL) return RB_EINVAL;
    if (op != SEQ_ADD && !all_found) return RB_EINVAL;
    rb_ctx* ctx = f->ctx;
    LOCK(ctx);
    if (f->kind != RB_BLOOM) return fail(ctx, RB_EINVAL, "wrong filter kind for this call");
    if (n_reads == 0) return RB_OK;
    int32_t rc = filter_ensure_logical(f);
    if (rc) return rc;
    void* d_missing = nullptr;
    if (op != SEQ_ADD) {
        rc = stage_get(ctx, 25, n_reads, &d_missing);
        if (rc) return rc;
        CK(cudaMemsetAsync(d_missing, 0, (size_t)n_reads, ctx->stream));
    }
    SeqUser u{f, mode, op, (uint8_t*)d_missing};
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    rc = for_each_launch(ctx, ra, f->k, seq_launch, &u, nullptr);
    if (rc) return rc;
    if (op != SEQ_ADD) {
        CK(cudaMemcpyAsync(all_found, d_missing, (size_t)n_reads, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        // missing -> found; a read without k-mers: containsAllKmers returns false (:629-631), lookupAndAddAllKmers true (empty loop)
        for (int64_t r = 0; r < n_reads; ++r) {
            const int64_t len = read_off ? read_len[r] : uniform_len;
            all_found[r] = len < f->k ? (op == SEQ_LOOKUP_AND_ADD_ALL ? 1 : 0) : (all_found[r] ? 0 : 1);
        }
    } else CK(cudaStreamSynchronize(ctx->stream));
    claim_invalidate(ctx);
    return RB_OK;
}

// ---- f3: k-mer multiplicity histogram by hash sampling (the reference runs the external `ntcard`, RNABloom.java:5745-5768) ----------------
struct rb_card { rb_ctx* ctx; int k, mode; CardTable ct; unsigned long long* hist; };
extern "C" int32_t rb_card_create(rb_ctx* ctx, int32_t k, int32_t stranded, int32_t sample_bits, int64_t table_slots, rb_card** out) {
    if (!ctx || !out || k < 1 || sample_bits < 0 || sample_bits > 30 || table_slots < 64) return RB_EINVAL;
    LOCK(ctx);
    rb_card* c = new rb_card();
    memset(c, 0, sizeof *c);
    c->ctx = ctx; c->k = k; c->mode = stranded ? RB_MODE_FWD : RB_MODE_CANON;
    int lg = 6; while ((1LL << lg) < table_slots && lg < 34) ++lg;
    c->ct.n_slots = 1ULL << lg; c->ct.shift = 64 - lg; c->ct.sample_bits = sample_bits;
    cudaError_t e = cudaMalloc(&c->ct.keys, (size_t)(c->ct.n_slots + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&c->ct.counts, (size_t)(c->ct.n_slots + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&c->ct.totals, 64);
    if (e == cudaSuccess) e = cudaMalloc(&c->hist, (size_t)65536 * 8);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->ct.keys, 0, (size_t)(c->ct.n_slots + 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->ct.counts, 0, (size_t)(c->ct.n_slots + 1) * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->ct.totals, 0, 64, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(c->ct.keys); cudaFree(c->ct.counts); cudaFree(c->ct.totals); cudaFree(c->hist);
        delete c;
        return fail(ctx, RB_ENOMEM, std::string("rb_card_create: ") + cudaGetErrorString(e));
    }
    *out = c;
    return RB_OK;
}
extern "C" int32_t rb_card_destroy(rb_card* c) {
    if (!c) return RB_EINVAL;
    rb_ctx* ctx = c->ctx;
    LOCK(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(c->ct.keys); cudaFree(c->ct.counts); cudaFree(c->ct.totals); cudaFree(c->hist);
    delete c;
    return RB_OK;
}
static int32_t card_launch(rb_ctx* ctx, const Ingest& ing, void* user) {
    rb_card* c = (rb_card*)user;
    const int grid = (int)div_up(div_up(ing.n_pos, kChunk), kThreads);
    PROF("k_card_add");
    if (c->mode == RB_MODE_FWD) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_card_add<0>)(ing, c->k, c->ct);
    else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_card_add<2>)(ing, c->k, c->ct);
    LAUNCH_CHECK();
    return RB_OK;
}
static int32_t card_add(rb_card* c, const ReadsArg& ra, int64_t* n_kmers_out) {
    rb_ctx* ctx = c->ctx;
    const int32_t rc = for_each_launch(ctx, ra, c->k, card_launch, c, n_kmers_out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}
extern "C" int32_t rb_card_add_reads(rb_card* c, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                     int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int64_t* n_kmers_out) {
    if (!c) return RB_EINVAL;
    LOCK(c->ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    return card_add(c, ra, n_kmers_out);
}
extern "C" int32_t rb_card_add_reads_dev(rb_card* c, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                         int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, int64_t* n_kmers_out) {
    if (!c) return RB_EINVAL;
    LOCK(c->ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, true};
    return card_add(c, ra, n_kmers_out);
}
// totals[0] = F1 (usable k-mer instances), [1] = sampled instances, [2] = sampled distinct k-mers, [3] = 2^sample_bits;
// hist[m - 1] = sampled distinct k-mers of multiplicity m (m <= max_mult <= 65535), hist[max_mult] = those above.  F0 ~ totals[2] * totals[3].
extern "C" int32_t rb_card_histogram(rb_card* c, int64_t* totals, int64_t* hist, int32_t max_mult) {
    if (!c || !totals || !hist || max_mult < 1 || max_mult > 65535) return RB_EINVAL;
    rb_ctx* ctx = c->ctx;
    LOCK(ctx);
    unsigned long long t3[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(t3, c->ct.totals, 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemsetAsync(c->hist, 0, (size_t)(max_mult + 1) * 8, ctx->stream));
    const int grid = (int)std::min<int64_t>(div_up((int64_t)c->ct.n_slots + 1, kThreads), (int64_t)ctx->sm_count * 16);
    RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_card_hist)(c->ct, c->hist, max_mult);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(hist, c->hist, (size_t)(max_mult + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (t3[2]) return fail(ctx, RB_ESTATE, "rb_card: the sample table is full -- more sample_bits or more table_slots");
    int64_t distinct = 0;
    for (int i = 0; i <= max_mult; ++i) distinct += hist[i];
    totals[0] = (int64_t)t3[0]; totals[1] = (int64_t)t3[1]; totals[2] = distinct; totals[3] = 1LL << c->ct.sample_bits;
    return RB_OK;
}

static int engine_from_env() {   // RB_ENGINE = direct | sliced | (anything else) auto
    const char* eng = getenv("RB_ENGINE");
    return (eng && !strcmp(eng, "direct")) ? RB_ENGINE_DIRECT : (eng && !strcmp(eng, "sliced")) ? RB_ENGINE_SLICED : RB_ENGINE_AUTO;
}
extern "C" int32_t rb_graph_create(rb_ctx* ctx, int64_t dbg_bits, int64_t cbf_bytes, int64_t pkbf_bits, int32_t hd, int32_t hc, int32_t hp,
                                   int32_t k, int32_t stranded, int32_t use_pairs, rb_graph** out) {
    if (!ctx || !out) return RB_EINVAL;
    LOCK(ctx);
    rb_graph* g = new rb_graph();
    memset(g, 0, sizeof *g);
    g->ctx = ctx; g->k = k; g->stranded = stranded ? 1 : 0; g->hd = hd; g->hc = hc; g->hp = hp; g->hmax = std::max(hd, hc);
    g->d_read = -1; g->d_frag = -1;
    int32_t rc = filter_alloc(ctx, RB_BLOOM, dbg_bits, hd, k, &g->dbg);
    if (!rc) rc = filter_alloc(ctx, RB_COUNTING, cbf_bytes, hc, k, &g->cbf);
    if (!rc && use_pairs) rc = filter_alloc(ctx, RB_BLOOM, pkbf_bits, hp, k, &g->rpk);
    if (rc) {
        if (g->dbg) filter_free(g->dbg);
        if (g->cbf) filter_free(g->cbf);
        if (g->rpk) filter_free(g->rpk);
        delete g;
        return rc;
    }
    g->dbg->in_graph = g->cbf->in_graph = true;
    if (g->rpk) g->rpk->in_graph = true;
    g->engine = engine_from_env();
    *out = g;
    return RB_OK;
}
static void sliced_engine_free(rb_graph* g);
extern "C" int32_t rb_graph_set_engine(rb_graph* g, int32_t engine) {
    if (!g || (engine != RB_ENGINE_DIRECT && engine != RB_ENGINE_SLICED && engine != RB_ENGINE_AUTO)) return RB_EINVAL;
    LOCK(g->ctx);
    g->engine = engine;
    return RB_OK;
}
extern "C" int32_t rb_graph_destroy(rb_graph* g) {
    if (!g) return RB_EINVAL;
    LOCK(g->ctx);
    sliced_engine_free(g);
    if (g->dbg && g->dbg->cs) { cudaStreamSynchronize(g->ctx->stream); cells_destroy(g->dbg->cs); }
    if (g->dbg) filter_free(g->dbg);
    if (g->cbf) filter_free(g->cbf);
    if (g->rpk) filter_free(g->rpk);
    if (g->fpk) filter_free(g->fpk);
    delete g;
    return RB_OK;
}
extern "C" int32_t rb_graph_init_fpkbf(rb_graph* g, int64_t bits, int32_t hp) {  // graph :352-359
    if (!g) return RB_EINVAL;
    LOCK(g->ctx);
    if (g->fpk) return rb_filter_empty(g->fpk);
    const int32_t rc = filter_alloc(g->ctx, RB_BLOOM, bits, hp, g->k, &g->fpk);
    if (!rc) g->fpk->in_graph = true;
    return rc;
}
extern "C" int32_t rb_graph_set_distances(rb_graph* g, int32_t d_read, int32_t d_frag) { if (!g) return RB_EINVAL; g->d_read = d_read; g->d_frag = d_frag; return RB_OK; }
extern "C" int32_t rb_graph_filter(rb_graph* g, int32_t which, rb_filter** out) {
    if (!g || !out) return RB_EINVAL;
    *out = which == RB_DBGBF ? g->dbg : which == RB_CBF ? g->cbf : which == RB_RPKBF ? g->rpk : which == RB_FPKBF ? g->fpk : nullptr;
    return RB_OK;
}
extern "C" int32_t rb_graph_clear(rb_graph* g) {
    if (!g) return RB_EINVAL;
    LOCK(g->ctx);
    int32_t rc = rb_filter_empty(g->dbg);
    if (!rc) rc = rb_filter_empty(g->cbf);
    if (!rc && g->rpk) rc = rb_filter_empty(g->rpk);
    if (!rc && g->fpk) rc = rb_filter_empty(g->fpk);
    return rc;
}
// the graph as the direct kernels see it: the logical arrays.  need_filters = false: the caller's kernel touches neither dbgbf nor cbf (pair
// inserts), so a graph whose current representation is the sliced engine's cell array stays that way
static GraphDev graph_view(rb_graph* g, bool need_filters = true) {
    if (need_filters) (void)cells_ensure_logical(g->ctx, g->dbg->cs);   // a failed launch surfaces at the caller's own launch check
    GraphDev gd;
    memset(&gd, 0, sizeof gd);
    gd.k = g->k; gd.hm = make_hm(g->k);
    gd.rng_seed = g->ctx->rng_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(g->ctx->launches + 1);  // fresh coins every launch
    gd.dbg.words = g->dbg->dev; gd.dbg.fm = make_fm(g->dbg->size); gd.dbg.num_hash = g->hd;
    gd.cbf.words = g->cbf->dev; gd.cbf.fm = make_fm(g->cbf->size); gd.cbf.num_hash = g->hc;
    return gd;
}
static BitFilter bit_view(rb_filter* f) { BitFilter b; b.words = f->dev; b.fm = make_fm(f->size); b.num_hash = f->num_hash; return b; }

template <int MODE, int MAXH>
static void launch_insert(int policy, int grid, cudaStream_t s, const Ingest& ing, const GraphDev& gd) {
    if (policy == POLICY_ADD) RB_LAUNCH(grid, kThreads, 0, s, k_graph_insert<MODE, MAXH, POLICY_ADD>)(ing, gd);
    else if (policy == POLICY_COUNT_IF_PRESENT) RB_LAUNCH(grid, kThreads, 0, s, k_graph_insert<MODE, MAXH, POLICY_COUNT_IF_PRESENT>)(ing, gd);
    else RB_LAUNCH(grid, kThreads, 0, s, k_graph_insert<MODE, MAXH, POLICY_DBG_ONLY>)(ing, gd);
}
template <int MAXH>
static void launch_insert_mode(int mode, int policy, int grid, cudaStream_t s, const Ingest& ing, const GraphDev& gd) {
    if (mode == RB_MODE_FWD) launch_insert<0, MAXH>(policy, grid, s, ing, gd);
    else if (mode == RB_MODE_RC) launch_insert<1, MAXH>(policy, grid, s, ing, gd);
    else launch_insert<2, MAXH>(policy, grid, s, ing, gd);
}
struct InsertUser { rb_graph* g; int mode, policy; int64_t direct_subbatch; bool direct_only; };
static int32_t sliced_insert_round(rb_graph* g, const Ingest& ing, int mode, int policy, bool* fell_back);
static int32_t sliced_count_round(rb_graph* g, const Ingest& ing, int mode, float* counts, int64_t* fh, int64_t* rh, bool* fell_back);
static int64_t sliced_round_kmers(const rb_ctx* ctx, bool host_results);
static bool sliced_supports(const rb_graph* g);
constexpr int64_t kAutoSlicedMinKmers = 1LL << 20;   // RB_ENGINE_AUTO: rounds smaller than this go to the direct kernels (14 launches and
                                                     // 3 stream syncs per sliced round only pay off on large rounds)
static bool use_sliced(const rb_graph* g, int64_t n_pos) {
    if (g->engine == RB_ENGINE_SLICED) return true;
    return g->engine == RB_ENGINE_AUTO && n_pos >= kAutoSlicedMinKmers && sliced_supports(g);
}
// the sliced engine amortises one sweep of the filters over a whole round, so its rounds are much larger than the direct launches
struct RoundSize {
    rb_ctx* c; int64_t keep;
    RoundSize(rb_graph* g, bool host_results) : c(g->ctx), keep(g->ctx->subbatch_kmers) {
        if (g->engine != RB_ENGINE_DIRECT && sliced_supports(g)) c->subbatch_kmers = sliced_round_kmers(c, host_results);
    }
    ~RoundSize() { c->subbatch_kmers = keep; }
};
static int32_t for_each_launch(rb_ctx* ctx, const ReadsArg& ra, int span, LaunchFn fn, void* user, int64_t* total_out);
static int32_t insert_launch(rb_ctx* ctx, const Ingest& ing, void* user) {
    InsertUser* u = (InsertUser*)user;
    if (!u->direct_only && use_sliced(u->g, ing.n_pos)) {
        bool fell_back = false;
        const int32_t rc = sliced_insert_round(u->g, ing, u->mode, u->policy, &fell_back);
        if (rc || !fell_back) return rc;
        if (ing.n_pos > u->direct_subbatch) {
            // a sliced round handed back (skew beyond the spill path, or no memory for the work buffers) is as large as 2^29 k-mers: the
            // direct engine takes it in launches of its own size, so that its claim table stays bounded (8 B x 2 x k-mers per launch)
            ReadsArg sub{ing.packed, ing.mask, ing.read_off, ing.read_len, ing.n_reads, ing.uniform_len, ing.uniform_stride, true};
            sub.rcm = ing.rcm;
            if (!ing.pos_off) sub.read0 = ing.first_base / ing.uniform_stride;
            InsertUser u2 = *u;
            u2.direct_only = true;
            const int64_t keep = ctx->subbatch_kmers;
            ctx->subbatch_kmers = u->direct_subbatch;
            const int32_t rc2 = for_each_launch(ctx, sub, u->g->k, insert_launch, &u2, nullptr);
            ctx->subbatch_kmers = keep;
            return rc2;
        }
    }
    GraphDev gd = graph_view(u->g);
    if (u->policy == POLICY_ADD) { const int32_t rc = claim_reserve(ctx, ing.n_pos, gd.dbg.words, &gd.ct); if (rc) return rc; }
    const int grid = (int)div_up(div_up(ing.n_pos, kChunk), kThreads);
    const int maxh = u->g->hmax;
    PROF("k_graph_insert");
    if (maxh <= 2) launch_insert_mode<2>(u->mode, u->policy, grid, ctx->stream, ing, gd);
    else if (maxh <= 3) launch_insert_mode<3>(u->mode, u->policy, grid, ctx->stream, ing, gd);
    else if (maxh <= 4) launch_insert_mode<4>(u->mode, u->policy, grid, ctx->stream, ing, gd);
    else launch_insert_mode<8>(u->mode, u->policy, grid, ctx->stream, ing, gd);
    LAUNCH_CHECK();
    return RB_OK;
}
struct PairUser { rb_graph* g; rb_filter* pk; int mode, d, op; };
template <int MODE, int MAXH>
static void launch_pairs_op(int op, int grid, cudaStream_t s, const Ingest& ing, const GraphDev& gd, const BitFilter& pk, int d) {
    if (op == 1) RB_LAUNCH(grid, kThreads, 0, s, k_pairs<MODE, MAXH, 1>)(ing, gd, pk, d, nullptr);
    else RB_LAUNCH(grid, kThreads, 0, s, k_pairs<MODE, MAXH, 2>)(ing, gd, pk, d, nullptr);
}
template <int MAXH>
static void launch_pairs_mode(int mode, int op, int grid, cudaStream_t s, const Ingest& ing, const GraphDev& gd, const BitFilter& pk, int d) {
    if (mode == RB_MODE_FWD) launch_pairs_op<0, MAXH>(op, grid, s, ing, gd, pk, d);
    else if (mode == RB_MODE_RC) launch_pairs_op<1, MAXH>(op, grid, s, ing, gd, pk, d);
    else launch_pairs_op<2, MAXH>(op, grid, s, ing, gd, pk, d);
}
static int32_t pairs_launch(rb_ctx* ctx, const Ingest& ing, void* user) {
    PairUser* u = (PairUser*)user;
    const GraphDev gd = graph_view(u->g, u->op == 2);   // only RB_PAIRS_EXISTING_ONLY looks the k-mers up in dbgbf
    const BitFilter pk = bit_view(u->pk);
    const int grid = (int)div_up(div_up(ing.n_pos, kChunk), kThreads);
    const int maxh = std::max(u->g->hd, u->pk->num_hash);
    PROF("k_pairs");
    if (maxh <= 2) launch_pairs_mode<2>(u->mode, u->op, grid, ctx->stream, ing, gd, pk, u->d);
    else if (maxh <= 3) launch_pairs_mode<3>(u->mode, u->op, grid, ctx->stream, ing, gd, pk, u->d);
    else if (maxh <= 4) launch_pairs_mode<4>(u->mode, u->op, grid, ctx->stream, ing, gd, pk, u->d);
    else launch_pairs_mode<8>(u->mode, u->op, grid, ctx->stream, ing, gd, pk, u->d);
    LAUNCH_CHECK();
    return RB_OK;
}
// graph.lookupReadKmerPair / lookupFragmentKmerPair (:526-532) for every pair position of every read: the test inside
// breakWith{Read,Frag}PairedKmers (util/GraphUtils.java:4184-4310).  found has max(0, len - k - d + 1) bytes per read (rb_kmer_offsets with k + d).
struct PairLookupUser { rb_graph* g; rb_filter* pk; int mode, d; uint8_t* found; };
static int32_t pair_lookup_launch(rb_ctx* ctx, const Ingest& ing, void* user) {
    PairLookupUser* u = (PairLookupUser*)user;
    const GraphDev gd = graph_view(u->g, false);
    const BitFilter pk = bit_view(u->pk);
    const int grid = (int)div_up(div_up(ing.n_pos, kChunk), kThreads);
    Ingest g = ing;
    const int64_t out0 = ing.out_base;
    g.out_base = 0;
    void* p;
    const int32_t rc = stage_get(ctx, 26, ing.n_pos, &p);
    if (rc) return rc;
    PROF("k_pairs");
    if (u->pk->num_hash <= 3) {
        if (u->mode == RB_MODE_FWD) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<0, 3, 3>)(g, gd, pk, u->d, (int64_t*)p);
        else if (u->mode == RB_MODE_RC) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<1, 3, 3>)(g, gd, pk, u->d, (int64_t*)p);
        else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<2, 3, 3>)(g, gd, pk, u->d, (int64_t*)p);
    } else {
        if (u->mode == RB_MODE_FWD) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<0, 8, 3>)(g, gd, pk, u->d, (int64_t*)p);
        else if (u->mode == RB_MODE_RC) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<1, 8, 3>)(g, gd, pk, u->d, (int64_t*)p);
        else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_pairs<2, 8, 3>)(g, gd, pk, u->d, (int64_t*)p);
    }
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(u->found + out0, p, (size_t)ing.n_pos, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}
static int graph_mode(const rb_graph* g, uint32_t flags);
extern "C" int32_t rb_graph_lookup_pairs_reads(rb_graph* g, int32_t which, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                               const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, uint32_t flags,
                                               uint8_t* found) {
    if (!g || !found || (which != RB_RPKBF && which != RB_FPKBF)) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    rb_filter* pk = which == RB_RPKBF ? g->rpk : g->fpk;
    const int d = which == RB_RPKBF ? g->d_read : g->d_frag;
    if (!pk || d < 1) return fail(ctx, RB_ESTATE, "pair look-up needs the pair filter and its distance");
    PairLookupUser u{g, pk, graph_mode(g, flags), d, found};
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    return for_each_launch(ctx, ra, g->k + d, pair_lookup_launch, &u, nullptr);
}
static int graph_mode(const rb_graph* g, uint32_t flags) {  // CanonicalHashFunction.java:179-206 ignores reverse-complement
    if (!g->stranded) return RB_MODE_CANON;
    return (flags & RB_REVCOMP) ? RB_MODE_RC : RB_MODE_FWD;
}
static int32_t graph_add_reads(rb_graph* g, const ReadsArg& ra, uint32_t flags, int64_t* n_kmers_out) {
    rb_ctx* ctx = g->ctx;
    const int mode = graph_mode(g, flags);
    int64_t total = 0;
    if (!(flags & RB_PAIRS_EXISTING_ONLY)) {
        InsertUser u{g, mode, (flags & RB_DBG_ONLY) ? POLICY_DBG_ONLY : (flags & RB_ADD_COUNT_IF_PRESENT) ? POLICY_COUNT_IF_PRESENT : POLICY_ADD,
                     ctx->subbatch_kmers, false};
        RoundSize rs(g, false);
        const int32_t rc = for_each_launch(ctx, ra, g->k, insert_launch, &u, &total);
        if (rc) return rc;
    }
    if (flags & (RB_STORE_READ_PAIRS | RB_PAIRS_EXISTING_ONLY)) {
        if (!g->rpk || g->d_read < 1) return fail(ctx, RB_ESTATE, "read-pair insert needs rpkbf and a read pair distance");
        PairUser u{g, g->rpk, mode, g->d_read, (flags & RB_PAIRS_EXISTING_ONLY) ? 2 : 1};
        const int32_t rc = for_each_launch(ctx, ra, g->k + g->d_read, pairs_launch, &u, nullptr);
        if (rc) return rc;
    }
    if (flags & RB_STORE_FRAG_PAIRS) {
        if (!g->fpk || g->d_frag < 1) return fail(ctx, RB_ESTATE, "fragment-pair insert needs fpkbf and a fragment pair distance");
        PairUser u{g, g->fpk, mode, g->d_frag, 1};
        const int32_t rc = for_each_launch(ctx, ra, g->k + g->d_frag, pairs_launch, &u, nullptr);
        if (rc) return rc;
    }
    if (n_kmers_out) *n_kmers_out = total;
    return RB_OK;
}
extern "C" int32_t rb_graph_add_reads(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                      int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, uint32_t flags, int64_t* n_kmers_out) {
    if (!g) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    const int32_t rc = graph_add_reads(g, ra, flags, n_kmers_out);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return RB_OK;
}
extern "C" int32_t rb_graph_add_reads_dev(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                          const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, uint32_t flags,
                                          int64_t* n_kmers_out) {
    if (!g) return RB_EINVAL;
    LOCK(g->ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, true};
    return graph_add_reads(g, ra, flags, n_kmers_out);
}

// ASCII records -> 2-bit codes + unusable-base mask + reverse-seed plane, on the device, in grow-only staging of the context (no
// allocation per call).  The chunk of records a Java worker hands over (RNABloom.java:551-634) arrives here as three host arrays.
static int32_t ascii_pack(rb_ctx* ctx, const char* bases, const char* quals, const int64_t* ascii_off, int64_t n_reads, int32_t min_qual, ReadsArg* out) {
    // records of one length (untrimmed short reads) take the uniform ingest layout: prefix k-merizer, no per-read tables at all; otherwise
    // per-read word offsets (reads start on 32-base word boundaries) -- a host pass over n_reads + 1 integers
    const int64_t len0 = ascii_off[1] - ascii_off[0];
    bool same_len = len0 > 0 && len0 <= INT32_MAX && !getenv("RB_ASCII_RAGGED");
    for (int64_t r = 1; r < n_reads && same_len; ++r) same_len = ascii_off[r + 1] - ascii_off[r] == len0;
    const bool uniform = same_len;
    std::vector<int64_t> word_off, read_off;
    std::vector<int32_t> read_len;
    int64_t words = 0;
    if (uniform) words = n_reads * ((len0 + 31) / 32);
    else {
        word_off.resize((size_t)n_reads + 1); read_off.resize((size_t)n_reads); read_len.resize((size_t)n_reads);
        for (int64_t r = 0; r < n_reads; ++r) {
            const int64_t len = ascii_off[r + 1] - ascii_off[r];
            if (len < 0 || len > INT32_MAX) return fail(ctx, RB_EINVAL, "ascii_off must be non-decreasing");
            word_off[(size_t)r] = words; read_off[(size_t)r] = words * 32; read_len[(size_t)r] = (int32_t)len;
            words += (len + 31) / 32;
        }
        word_off[(size_t)n_reads] = words;
    }
    const int64_t a_lo = ascii_off[0], a_hi = ascii_off[n_reads];
    void *d_b, *d_q = nullptr, *d_ao = nullptr, *d_wo = nullptr, *d_ro = nullptr, *d_rl = nullptr, *d_packed, *d_mask, *d_rcm;
    int32_t rc;
    if ((rc = stage_get(ctx, 16, (a_hi - a_lo) + 16, &d_b))) return rc;
    if (quals && (rc = stage_get(ctx, 17, (a_hi - a_lo) + 16, &d_q))) return rc;
    if (!uniform) {
        if ((rc = stage_get(ctx, 18, (n_reads + 1) * 8, &d_ao))) return rc;
        if ((rc = stage_get(ctx, 19, (n_reads + 1) * 8, &d_wo))) return rc;
        if ((rc = stage_get(ctx, 20, n_reads * 8, &d_ro))) return rc;
        if ((rc = stage_get(ctx, 21, n_reads * 4, &d_rl))) return rc;
    }
    if ((rc = stage_get(ctx, 22, (words + 2) * 8, &d_packed))) return rc;
    if ((rc = stage_get(ctx, 23, (words + 2) * 4, &d_mask))) return rc;
    if ((rc = stage_get(ctx, 24, (words + 2) * 4, &d_rcm))) return rc;
    CK(cudaMemcpyAsync(d_b, bases + a_lo, (size_t)(a_hi - a_lo), cudaMemcpyHostToDevice, ctx->stream));
    if (quals) CK(cudaMemcpyAsync(d_q, quals + a_lo, (size_t)(a_hi - a_lo), cudaMemcpyHostToDevice, ctx->stream));
    if (!uniform) {
        CK(cudaMemcpyAsync(d_ao, ascii_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_wo, word_off.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_ro, read_off.data(), (size_t)n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_rl, read_len.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (words > 0) {
        RB_LAUNCH((int)div_up(words, kThreads), kThreads, 0, ctx->stream, k_pack_ascii)((const char*)d_b - a_lo, d_q ? (const char*)d_q - a_lo : nullptr, (const int64_t*)d_ao,
                                                                                 (const int64_t*)d_wo, n_reads, words, min_qual, (uint64_t*)d_packed,
                                                                                 (uint32_t*)d_mask, (uint32_t*)d_rcm, uniform ? (int)len0 : 0, a_lo);
        LAUNCH_CHECK();
    }
    CK(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope
    if (uniform) *out = ReadsArg{(const uint64_t*)d_packed, (const uint32_t*)d_mask, nullptr, nullptr, n_reads, (int32_t)len0, ((len0 + 31) / 32) * 32, true};
    else *out = ReadsArg{(const uint64_t*)d_packed, (const uint32_t*)d_mask, (const int64_t*)d_ro, (const int32_t*)d_rl, n_reads, 0, 0, true};
    out->rcm = (const uint32_t*)d_rcm;
    return RB_OK;
}
extern "C" int32_t rb_graph_add_reads_ascii(rb_graph* g, const char* bases, const char* quals, const int64_t* ascii_off, int64_t n_reads,
                                            int32_t min_qual, uint32_t flags, int64_t* n_kmers_out) {
    if (!g || !bases || !ascii_off || n_reads < 0) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    if (n_kmers_out) *n_kmers_out = 0;
    if (n_reads == 0) return RB_OK;
    ReadsArg ra{};
    int32_t rc = ascii_pack(ctx, bases, quals, ascii_off, n_reads, min_qual, &ra);
    if (rc) return rc;
    rc = graph_add_reads(g, ra, flags, n_kmers_out);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (!rc && e != cudaSuccess) rc = fail(ctx, RB_ECUDA, cudaGetErrorString(e));
    return rc;
}

// ---- f2: the reference's .2bit fragment records (stage-3 rebuild: FragmentsToGraphWorker, RNABloom.java:1489-1516 reads them with
// io/NucleotideBitsReader.java:39-49) ------------------------------------------------------------------------------------------------------------
// Host helpers: the codec itself (deterministic for ACGTU input; the reference writes a RANDOM base for anything else, SeqBitsUtils.java:138-156)
extern "C" int64_t rb_2bit_record_bytes(int32_t seq_len) { return 4 + (int64_t)(seq_len / 4 + (seq_len % 4 ? 1 : 0)); }
extern "C" int64_t rb_2bit_encode_records(const char* bases, const int64_t* ascii_off, int64_t n_reads, uint8_t* out) {
    int64_t at = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        const int64_t a0 = ascii_off[r];
        const int len = (int)(ascii_off[r + 1] - a0);
        if (out) { out[at] = (uint8_t)(len >> 24); out[at + 1] = (uint8_t)(len >> 16); out[at + 2] = (uint8_t)(len >> 8); out[at + 3] = (uint8_t)len; }   // intToFourBytes :209
        at += 4;
        for (int i = 0; i < len; i += 4) {
            int v = 0;
            for (int j = 0; j < 4; ++j) {
                int code = 0;   // bases past the end count as A (seqToBits :236-243)
                if (i + j < len) {
                    switch (bases[a0 + i + j]) {
                        case 'C': case 'c': code = 1; break;
                        case 'G': case 'g': code = 2; break;
                        case 'T': case 't': case 'U': case 'u': code = 3; break;
                        default: code = 0;
                    }
                }
                v = v * 4 + code;
            }
            if (out) out[at] = (uint8_t)(v - 128);
            ++at;
        }
    }
    return at;
}
// number of complete records in `records` (n_bytes bytes); fills data_off / read_len when non-NULL (capacity max_reads).  -1: truncated stream
extern "C" int64_t rb_2bit_index_records(const uint8_t* records, int64_t n_bytes, int64_t* data_off, int32_t* read_len, int64_t max_reads) {
    int64_t at = 0, n = 0;
    while (at < n_bytes) {
        if (at + 4 > n_bytes) return -1;
        const int64_t len = ((int64_t)records[at] << 24) | ((int64_t)records[at + 1] << 16) | ((int64_t)records[at + 2] << 8) | (int64_t)records[at + 3];   // fourBytesToInt :213
        const int64_t nb = len / 4 + (len % 4 ? 1 : 0);
        if (len < 0 || len > INT32_MAX || at + 4 + nb > n_bytes) return -1;
        if (data_off && n < max_reads) { data_off[n] = at + 4; read_len[n] = (int32_t)len; }
        at += 4 + nb;
        ++n;
    }
    return n;
}
// graph.add & co. for a buffer of .2bit records in host memory: the tetramer bytes are re-packed on the GPU
extern "C" int32_t rb_graph_add_reads_2bit(rb_graph* g, const uint8_t* records, int64_t n_bytes, uint32_t flags, int64_t* n_reads_out, int64_t* n_kmers_out) {
    if (!g || (n_bytes > 0 && !records) || n_bytes < 0) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    if (n_reads_out) *n_reads_out = 0;
    if (n_kmers_out) *n_kmers_out = 0;
    const int64_t n_reads = rb_2bit_index_records(records, n_bytes, nullptr, nullptr, 0);
    if (n_reads < 0) return fail(ctx, RB_EINVAL, "2bit records: truncated or corrupt stream");
    if (n_reads == 0) return RB_OK;
    std::vector<int64_t> data_off((size_t)n_reads), word_off((size_t)n_reads + 1), read_off((size_t)n_reads);
    std::vector<int32_t> read_len((size_t)n_reads);
    rb_2bit_index_records(records, n_bytes, data_off.data(), read_len.data(), n_reads);
    int64_t words = 0;
    for (int64_t r = 0; r < n_reads; ++r) { word_off[(size_t)r] = words; read_off[(size_t)r] = words * 32; words += (read_len[(size_t)r] + 31) / 32; }
    word_off[(size_t)n_reads] = words;
    void *d_rec, *d_do, *d_wo, *d_ro, *d_rl, *d_packed;
    int32_t rc;
    if ((rc = stage_get(ctx, 16, n_bytes + 16, &d_rec)) || (rc = stage_get(ctx, 18, n_reads * 8, &d_do)) || (rc = stage_get(ctx, 19, (n_reads + 1) * 8, &d_wo)) ||
        (rc = stage_get(ctx, 20, n_reads * 8, &d_ro)) || (rc = stage_get(ctx, 21, n_reads * 4, &d_rl)) || (rc = stage_get(ctx, 22, (words + 2) * 8, &d_packed))) return rc;
    CK(cudaMemcpyAsync(d_rec, records, (size_t)n_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_do, data_off.data(), (size_t)n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_wo, word_off.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_ro, read_off.data(), (size_t)n_reads * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_rl, read_len.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (words > 0) {
        RB_LAUNCH((int)div_up(words, kThreads), kThreads, 0, ctx->stream, k_unpack_2bit)((const uint8_t*)d_rec, (const int64_t*)d_do, (const int32_t*)d_rl, (const int64_t*)d_wo,
                                                                                  n_reads, words, (uint64_t*)d_packed);
        LAUNCH_CHECK();
    }
    CK(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope
    ReadsArg ra{(const uint64_t*)d_packed, nullptr, (const int64_t*)d_ro, (const int32_t*)d_rl, n_reads, 0, 0, true};
    rc = graph_add_reads(g, ra, flags, n_kmers_out);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (!rc && e != cudaSuccess) rc = fail(ctx, RB_ECUDA, cudaGetErrorString(e));
    if (!rc && n_reads_out) *n_reads_out = n_reads;
    return rc;
}

struct CountUser { rb_graph* g; int mode; float* counts; int64_t *fh, *rh; bool on_device; };
template <int MAXH>
static void launch_count_mode(int mode, int grid, cudaStream_t s, const Ingest& ing, const GraphDev& gd, float* c, int64_t* f, int64_t* r) {
    if (mode == RB_MODE_FWD) RB_LAUNCH(grid, kThreads, 0, s, k_graph_count<0, MAXH>)(ing, gd, c, f, r);
    else RB_LAUNCH(grid, kThreads, 0, s, k_graph_count<2, MAXH>)(ing, gd, c, f, r);
}
static int32_t count_launch(rb_ctx* ctx, const Ingest& ing_in, void* user) {
    CountUser* u = (CountUser*)user;
    Ingest ing = ing_in;
    float* dc = u->counts; int64_t *df = u->fh, *dr = u->rh;
    const int64_t out0 = ing.out_base;
    const int par = (int)(ctx->count_launch_index++ & 1);
    if (!u->on_device) {  // results go through device staging (double buffered), indexed by launch-local position
        ing.out_base = 0;
        int32_t rc; void* p;
        const int s0 = par ? 8 : 5;
        if (u->counts) { rc = stage_get(ctx, s0, ing.n_pos * 4, &p); if (rc) return rc; dc = (float*)p; }
        if (u->fh) { rc = stage_get(ctx, s0 + 1, ing.n_pos * 8, &p); if (rc) return rc; df = (int64_t*)p; }
        if (u->rh) { rc = stage_get(ctx, s0 + 2, ing.n_pos * 8, &p); if (rc) return rc; dr = (int64_t*)p; }
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[par], 0));   // the previous D2H out of this staging set is done
    }
    bool sliced_done = false;
    if (dc && use_sliced(u->g, ing.n_pos)) {
        bool fell_back = false;
        const int32_t rc = sliced_count_round(u->g, ing, u->mode, dc, df, dr, &fell_back);
        if (rc) return rc;
        sliced_done = !fell_back;
    }
    const int grid = (int)div_up(div_up(ing.n_pos, kChunk), kThreads);
    const int maxh = u->g->hmax;
    GraphDev gd;
    if (!sliced_done) gd = graph_view(u->g);
    if (!sliced_done) PROF("k_graph_count");
    if (sliced_done) { /* counts are already in dc */ }
    else if (maxh <= 2) launch_count_mode<2>(u->mode, grid, ctx->stream, ing, gd, dc, df, dr);
    else if (maxh <= 3) launch_count_mode<3>(u->mode, grid, ctx->stream, ing, gd, dc, df, dr);
    else if (maxh <= 4) launch_count_mode<4>(u->mode, grid, ctx->stream, ing, gd, dc, df, dr);
    else launch_count_mode<8>(u->mode, grid, ctx->stream, ing, gd, dc, df, dr);
    if (!sliced_done) LAUNCH_CHECK();
    if (!u->on_device) {   // D2H on the copy stream: overlaps the next launch's kernel
        CK(cudaEventRecord(ctx->ev_kernel[par], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_kernel[par], 0));
        // in pieces: when the driver puts this stream and the compute stream's H2D copies on the same copy engine, a 2 GB D2H in one
        // piece keeps the next call's 160 MB of reads (and with them its kernels) waiting for all of it (measured: e2e 119 -> 154 ms per step
        // on such a box); between pieces the engine takes the other stream's copy
        auto d2h = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
            const size_t piece = (size_t)32 << 20;
            for (size_t o = 0; o < bytes; o += piece) {
                const cudaError_t e = cudaMemcpyAsync((char*)dst + o, (const char*)src + o, std::min(piece, bytes - o), cudaMemcpyDeviceToHost, ctx->copy_stream);
                if (e != cudaSuccess) return e;
            }
            return cudaSuccess;
        };
        if (u->counts) CK(d2h(u->counts + out0, dc, (size_t)ing.n_pos * 4));
        if (u->fh) CK(d2h(u->fh + out0, df, (size_t)ing.n_pos * 8));
        if (u->rh) CK(d2h(u->rh + out0, dr, (size_t)ing.n_pos * 8));
        CK(cudaEventRecord(ctx->ev_copy[par], ctx->copy_stream));
    }
    return RB_OK;
}
// big_rounds: host results, but the caller does not wait for them (rb_graph_count_reads_async): one round as large as for device results,
// its results parked in device staging while the copy stream drains them behind the caller's next calls
static int32_t graph_count_reads(rb_graph* g, const ReadsArg& ra, float* counts, int64_t* fh, int64_t* rh, int64_t* n_out, bool results_on_device,
                                 bool big_rounds = false) {
    CountUser u{g, g->stranded ? RB_MODE_FWD : RB_MODE_CANON, counts, fh, g->stranded ? nullptr : rh, results_on_device};
    RoundSize rs(g, !results_on_device && !big_rounds);
    return for_each_launch(g->ctx, ra, g->k, count_launch, &u, n_out);
}
extern "C" int32_t rb_graph_count_reads(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                        int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, float* counts, int64_t* fhash, int64_t* rhash,
                                        int64_t* n_kmers_out) {
    if (!g) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    const int32_t rc = graph_count_reads(g, ra, counts, fhash, rhash, n_kmers_out, false);
    if (rc) { cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream); return rc; }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    return RB_OK;
}
// The same without waiting for the results: the call returns when the last round's kernels are queued; the device->host copies run on
// the copy stream behind whatever the caller does next (typically the next rb_graph_add_reads: 2 GB of counts per 504 M k-mers take
// longer over PCIe than the look-up kernels themselves).  `packed` / `mask` and the result buffers (pinned host memory) must stay
// untouched until rb_ctx_wait(ticket) returns.
extern "C" int32_t rb_graph_count_reads_async(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                              int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, float* counts, int64_t* fhash, int64_t* rhash,
                                              int64_t* n_kmers_out, int64_t* ticket) {
    if (!g || !ticket) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, false};
    const int32_t rc = graph_count_reads(g, ra, counts, fhash, rhash, n_kmers_out, false, true);
    if (rc) { cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream); return rc; }
    const int slot = (int)(ctx->ticket_seq % 8);
    if (!ctx->ticket_ev[slot]) CK(cudaEventCreateWithFlags(&ctx->ticket_ev[slot], cudaEventDisableTiming));
    // the copy stream has every D2H of this call queued behind its kernels (count_launch); a call without any launch completes at once
    CK(cudaEventRecord(ctx->ticket_ev[slot], ctx->copy_stream));
    *ticket = ++ctx->ticket_seq;
    return RB_OK;
}
// Blocks until the results of that call (and of every earlier asynchronous call) are in host memory.  Takes no lock: a thread may wait
// while another one is inside an insert call of the same context.
extern "C" int32_t rb_ctx_wait(rb_ctx* ctx, int64_t ticket) {
    if (!ctx || ticket < 1 || ticket > ctx->ticket_seq) return RB_EINVAL;
    cudaEvent_t ev = ctx->ticket_ev[(ticket - 1) % 8];   // a slot recorded again by a later ticket completes later on the same stream: still correct
    if (!ev) return RB_EINVAL;
    const cudaError_t e = cudaEventSynchronize(ev);
    return e == cudaSuccess ? RB_OK : RB_ECUDA;
}
extern "C" int32_t rb_graph_count_reads_dev(rb_graph* g, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off,
                                            const int32_t* read_len, int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, float* counts,
                                            int64_t* fhash, int64_t* rhash, int64_t* n_kmers_out) {
    if (!g) return RB_EINVAL;
    LOCK(g->ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, true};
    return graph_count_reads(g, ra, counts, fhash, rhash, n_kmers_out, true);
}
// graph.getKmers(String) for a chunk of sequences (graph :1224-1234; bloom/hash/HashFunction.java:55-85): counts and hashes of every
// k-mer window; windows over a non-ACGTU character get count 0 but are still hashed through exactly as NTHash does (forward row 0,
// reverse row c & 0x07).  Results in host memory.
extern "C" int32_t rb_graph_count_reads_ascii(rb_graph* g, const char* bases, const int64_t* ascii_off, int64_t n_reads, float* counts, int64_t* fhash,
                                              int64_t* rhash, int64_t* n_kmers_out) {
    if (!g || !bases || !ascii_off || n_reads < 0) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    if (n_kmers_out) *n_kmers_out = 0;
    if (n_reads == 0) return RB_OK;
    ReadsArg ra{};
    int32_t rc = ascii_pack(ctx, bases, nullptr, ascii_off, n_reads, 0, &ra);
    if (!rc) rc = graph_count_reads(g, ra, counts, fhash, rhash, n_kmers_out, false);
    const cudaError_t e1 = cudaStreamSynchronize(ctx->stream), e2 = cudaStreamSynchronize(ctx->copy_stream);
    if (!rc && (e1 != cudaSuccess || e2 != cudaSuccess)) rc = fail(ctx, RB_ECUDA, cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return rc;
}
// NTHashIterator / ReverseComplement... / Canonical... over ASCII sequences: hashes of every window, bit-exact for EVERY byte value
extern "C" int32_t rb_kmerize_ascii(rb_ctx* ctx, const char* bases, const int64_t* ascii_off, int64_t n_reads, int32_t k, int32_t mode, int64_t* fhash,
                                    int64_t* rhash, int64_t* base) {
    if (!ctx || !bases || !ascii_off || n_reads < 0 || k < 1 || mode < 0 || mode > 2) return RB_EINVAL;
    LOCK(ctx);
    if (n_reads == 0) return RB_OK;
    ReadsArg ra{};
    const int32_t rc = ascii_pack(ctx, bases, nullptr, ascii_off, n_reads, 0, &ra);
    if (rc) return rc;
    KmerizeUser u{k, mode, 0, nullptr, nullptr, nullptr, fhash, rhash, base, false};
    return for_each_launch(ctx, ra, k, kmerize_launch, &u, nullptr);
}
extern "C" int32_t rb_graph_add_hashes(rb_graph* g, const int64_t* base, int64_t n, uint32_t flags) {
    if (!g) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    if (flags & RB_DBG_ONLY) return run_hash_op(ctx, OP_BF_ADD, graph_view(g), g->hmax, base, n, nullptr, nullptr);
    return run_hash_op(ctx, (flags & RB_ADD_COUNT_IF_PRESENT) ? OP_GRAPH_COUNT_IF_PRESENT : OP_GRAPH_ADD, graph_view(g), g->hmax, base, n, nullptr, nullptr);
}
extern "C" int32_t rb_graph_count_hashes(rb_graph* g, const int64_t* base, int64_t n, float* counts) {
    if (!g || !counts) return RB_EINVAL;
    LOCK(g->ctx);
    return run_hash_op(g->ctx, OP_GRAPH_COUNT, graph_view(g), g->hmax, base, n, nullptr, counts);
}
extern "C" int32_t rb_graph_add_pair_hashes(rb_graph* g, int32_t which, const int64_t* pair_hash, int64_t n) {
    if (!g) return RB_EINVAL;
    rb_filter* f = which == RB_RPKBF ? g->rpk : which == RB_FPKBF ? g->fpk : nullptr;
    if (!f) return fail(g->ctx, RB_ESTATE, "pair filter not initialised");
    return rb_filter_add_hashes(f, pair_hash, n);
}
extern "C" int32_t rb_graph_lookup_pair_hashes(rb_graph* g, int32_t which, const int64_t* pair_hash, int64_t n, uint8_t* out) {
    if (!g) return RB_EINVAL;
    rb_filter* f = which == RB_RPKBF ? g->rpk : which == RB_FPKBF ? g->fpk : nullptr;
    if (!f) return fail(g->ctx, RB_ESTATE, "pair filter not initialised");
    return rb_filter_lookup_hashes(f, pair_hash, n, out);
}

// ---- f1: batched neighbour query (graph/Kmer.java:213-253, graph/CanonicalKmer.java:232-271) ------------------------------------------------
// out arrays: [n][2][4] -- per k-mer the 4 successors (A,C,G,T) then the 4 predecessors
extern "C" int32_t rb_graph_neighbor_counts(rb_graph* g, const int64_t* fhash, const int64_t* rhash, const uint8_t* first_base, const uint8_t* last_base,
                                            int64_t n, float* counts, int64_t* nbr_fhash, int64_t* nbr_rhash) {
    if (!g || n < 0 || (n > 0 && (!fhash || !first_base || !last_base || !counts))) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    const int canonical = g->stranded ? 0 : 1;
    if (canonical && n > 0 && !rhash) return fail(ctx, RB_EINVAL, "neighbours of canonical k-mers need the reverse-strand hashes");
    const GraphDev gd = graph_view(g);
    const int64_t step = std::max<int64_t>(1024, ctx->subbatch_kmers / 8);
    for (int64_t i0 = 0; i0 < n; i0 += step) {
        const int64_t m = std::min(step, n - i0);
        void *df, *dr = nullptr, *d1, *d2, *dc, *dnf = nullptr, *dnr = nullptr;
        int32_t rc = stage_get(ctx, 0, m * 8, &df); if (rc) return rc;
        if (canonical) { rc = stage_get(ctx, 1, m * 8, &dr); if (rc) return rc; }
        rc = stage_get(ctx, 2, m, &d1); if (rc) return rc;
        rc = stage_get(ctx, 3, m, &d2); if (rc) return rc;
        rc = stage_get(ctx, 5, m * 32, &dc); if (rc) return rc;
        if (nbr_fhash) { rc = stage_get(ctx, 6, m * 64, &dnf); if (rc) return rc; }
        if (nbr_rhash && canonical) { rc = stage_get(ctx, 7, m * 64, &dnr); if (rc) return rc; }
        CK(cudaMemcpyAsync(df, fhash + i0, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (canonical) CK(cudaMemcpyAsync(dr, rhash + i0, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d1, first_base + i0, (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d2, last_base + i0, (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        const int grid = (int)div_up(2 * m, kThreads);
        PROF("k_neighbors");
        if (g->hmax <= 3) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_neighbors<3>)((const int64_t*)df, (const int64_t*)dr, (const uint8_t*)d1, (const uint8_t*)d2, m, gd, canonical, (float*)dc, (int64_t*)dnf, (int64_t*)dnr);
        else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_neighbors<8>)((const int64_t*)df, (const int64_t*)dr, (const uint8_t*)d1, (const uint8_t*)d2, m, gd, canonical, (float*)dc, (int64_t*)dnf, (int64_t*)dnr);
        LAUNCH_CHECK();
        CK(cudaMemcpyAsync(counts + i0 * 8, dc, (size_t)m * 32, cudaMemcpyDeviceToHost, ctx->stream));
        if (dnf) CK(cudaMemcpyAsync(nbr_fhash + i0 * 8, dnf, (size_t)m * 64, cudaMemcpyDeviceToHost, ctx->stream));
        if (dnr) CK(cudaMemcpyAsync(nbr_rhash + i0 * 8, dnr, (size_t)m * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return RB_OK;
}

// variants of the first / last base (graph/Kmer.java:357-405): out arrays [n][2][4] -- per k-mer the 4 left variants (A,C,G,T at position 0) then
// the 4 right variants (position k-1); the entry of the k-mer's own base is the k-mer itself
extern "C" int32_t rb_graph_variant_counts(rb_graph* g, const int64_t* fhash, const int64_t* rhash, const uint8_t* first_base, const uint8_t* last_base,
                                           int64_t n, float* counts, int64_t* var_fhash, int64_t* var_rhash) {
    if (!g || n < 0 || (n > 0 && (!fhash || !first_base || !last_base || !counts))) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    const int canonical = g->stranded ? 0 : 1;
    if (canonical && n > 0 && !rhash) return fail(ctx, RB_EINVAL, "variants of canonical k-mers need the reverse-strand hashes");
    const GraphDev gd = graph_view(g);
    const int64_t step = std::max<int64_t>(1024, ctx->subbatch_kmers / 8);
    for (int64_t i0 = 0; i0 < n; i0 += step) {
        const int64_t m = std::min(step, n - i0);
        void *df, *dr = nullptr, *d1, *d2, *dc, *dnf = nullptr, *dnr = nullptr;
        int32_t rc = stage_get(ctx, 0, m * 8, &df); if (rc) return rc;
        if (canonical) { rc = stage_get(ctx, 1, m * 8, &dr); if (rc) return rc; }
        rc = stage_get(ctx, 2, m, &d1); if (rc) return rc;
        rc = stage_get(ctx, 3, m, &d2); if (rc) return rc;
        rc = stage_get(ctx, 5, m * 32, &dc); if (rc) return rc;
        if (var_fhash) { rc = stage_get(ctx, 6, m * 64, &dnf); if (rc) return rc; }
        if (var_rhash && canonical) { rc = stage_get(ctx, 7, m * 64, &dnr); if (rc) return rc; }
        CK(cudaMemcpyAsync(df, fhash + i0, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (canonical) CK(cudaMemcpyAsync(dr, rhash + i0, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d1, first_base + i0, (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d2, last_base + i0, (size_t)m, cudaMemcpyHostToDevice, ctx->stream));
        const int grid = (int)div_up(2 * m, kThreads);
        PROF("k_variants");
        if (g->hmax <= 3) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_variants<3>)((const int64_t*)df, (const int64_t*)dr, (const uint8_t*)d1, (const uint8_t*)d2, m, gd, canonical, (float*)dc, (int64_t*)dnf, (int64_t*)dnr);
        else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_variants<8>)((const int64_t*)df, (const int64_t*)dr, (const uint8_t*)d1, (const uint8_t*)d2, m, gd, canonical, (float*)dc, (int64_t*)dnf, (int64_t*)dnr);
        LAUNCH_CHECK();
        CK(cudaMemcpyAsync(counts + i0 * 8, dc, (size_t)m * 32, cudaMemcpyDeviceToHost, ctx->stream));
        if (dnf) CK(cudaMemcpyAsync(var_fhash + i0 * 8, dnf, (size_t)m * 64, cudaMemcpyDeviceToHost, ctx->stream));
        if (dnr) CK(cudaMemcpyAsync(var_rhash + i0 * 8, dnr, (size_t)m * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return RB_OK;
}
// Kmer.getMaxCovSuccessor / getMaxCovPredecessor (graph/Kmer.java:301-355) for a batch: best[n][2] = base code (0..3) of the successor /
// predecessor with the largest count >= min_cov, the first of equal ones in A, C, G, T order; -1 if there is none.  best_count[n][2] likewise.
extern "C" int32_t rb_graph_max_cov_neighbors(rb_graph* g, const int64_t* fhash, const int64_t* rhash, const uint8_t* first_base, const uint8_t* last_base,
                                              int64_t n, float min_cov, int8_t* best, float* best_count, int64_t* best_fhash, int64_t* best_rhash) {
    if (!g || n < 0 || (n > 0 && !best)) return RB_EINVAL;
    std::vector<float> counts((size_t)n * 8);
    std::vector<int64_t> nf((size_t)(best_fhash ? n * 8 : 0)), nr((size_t)(best_rhash ? n * 8 : 0));
    const int32_t rc = rb_graph_neighbor_counts(g, fhash, rhash, first_base, last_base, n, counts.data(), best_fhash ? nf.data() : nullptr,
                                                best_rhash ? nr.data() : nullptr);
    if (rc) return rc;
    for (int64_t i = 0; i < n * 2; ++i) {
        float bc = -1.f; int b = -1;
        for (int c = 0; c < 4; ++c) { const float v = counts[(size_t)i * 4 + c]; if (v >= min_cov && v > bc) { bc = v; b = c; } }   // :312-317
        best[i] = (int8_t)b;
        if (best_count) best_count[i] = b < 0 ? 0.f : bc;
        if (best_fhash) best_fhash[i] = b < 0 ? 0 : nf[(size_t)i * 4 + b];
        if (best_rhash) best_rhash[i] = b < 0 ? 0 : (g->stranded ? 0 : nr[(size_t)i * 4 + b]);
    }
    return RB_OK;
}
// GraphUtils.greedyExtendRight / greedyExtendLeft with lookahead <= 1 for a batch of start k-mers (util/GraphUtils.java:1961-1976): kmer_bits =
// 2 x uint64 per k-mer (2-bit codes, base i at bits 2 * (i & 31) of word i >> 5), k <= 64.  ext_codes[n][bound] receives the added bases in
// walking order (to the left: the bases as they are prepended), ext_len[n] how many.
extern "C" int32_t rb_graph_greedy_extend(rb_graph* g, const uint64_t* kmer_bits, const int64_t* fhash, const int64_t* rhash, int64_t n, int32_t right,
                                          int32_t bound, float min_cov, int32_t* ext_len, uint8_t* ext_codes, float* ext_counts) {
    if (!g || n < 0 || bound < 1 || (n > 0 && (!kmer_bits || !fhash || !ext_len || !ext_codes))) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    if (g->k > 64) return fail(ctx, RB_EINVAL, "greedy extension carries the k-mer in 128 bits: k <= 64");
    const int canonical = g->stranded ? 0 : 1;
    if (canonical && n > 0 && !rhash) return fail(ctx, RB_EINVAL, "canonical k-mers need the reverse-strand hashes");
    const GraphDev gd = graph_view(g);
    const int64_t step = std::max<int64_t>(1024, std::min<int64_t>(1 << 22, (1LL << 30) / bound));
    for (int64_t i0 = 0; i0 < n; i0 += step) {
        const int64_t m = std::min(step, n - i0);
        void *dk, *df, *dr = nullptr, *dl, *dc, *dn = nullptr;
        int32_t rc;
        if ((rc = stage_get(ctx, 0, m * 16, &dk)) || (rc = stage_get(ctx, 1, m * 8, &df)) || (rc = stage_get(ctx, 3, m * 4, &dl)) ||
            (rc = stage_get(ctx, 5, m * bound, &dc))) return rc;
        if (canonical && (rc = stage_get(ctx, 2, m * 8, &dr))) return rc;
        if (ext_counts && (rc = stage_get(ctx, 6, m * bound * 4, &dn))) return rc;
        CK(cudaMemcpyAsync(dk, kmer_bits + 2 * i0, (size_t)m * 16, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(df, fhash + i0, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (canonical) CK(cudaMemcpyAsync(dr, rhash + i0, (size_t)m * 8, cudaMemcpyHostToDevice, ctx->stream));
        const int grid = (int)div_up(m, kThreads);
        PROF("k_greedy_extend");
        if (g->hmax <= 3) RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_greedy_extend<3>)((const uint64_t*)dk, (const int64_t*)df, (const int64_t*)dr, m, gd, canonical, right ? 1 : 0, bound, min_cov, (int32_t*)dl, (uint8_t*)dc, (float*)dn);
        else RB_LAUNCH(grid, kThreads, 0, ctx->stream, k_greedy_extend<8>)((const uint64_t*)dk, (const int64_t*)df, (const int64_t*)dr, m, gd, canonical, right ? 1 : 0, bound, min_cov, (int32_t*)dl, (uint8_t*)dc, (float*)dn);
        LAUNCH_CHECK();
        CK(cudaMemcpyAsync(ext_len + i0, dl, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ext_codes + i0 * bound, dc, (size_t)m * bound, cudaMemcpyDeviceToHost, ctx->stream));
        if (ext_counts) CK(cudaMemcpyAsync(ext_counts + i0 * bound, dn, (size_t)m * bound * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return RB_OK;
}

// ---- persistence (graph :297-339, file ctor :121-189) -------------------------------------------------------------------------
extern "C" int32_t rb_graph_save(rb_graph* g, const char* path) {
    if (!g || !path) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    const std::string p(path);
    const std::string desc = "dbgbfCbfMaxNumHash:" + std::to_string(g->hmax) + "\nstranded:" + (g->stranded ? "true" : "false") + "\nk:" +
                             std::to_string(g->k) + "\nreadPairedKmersDistance:" + std::to_string(g->d_read) +
                             "\nfragmentPairedKmersDistance:" + std::to_string(g->d_frag) + "\n";
    int32_t rc = write_file(ctx, path, desc.data(), desc.size());
    if (!rc) rc = rb_filter_save(g->dbg, (p + ".dbgbf.desc").c_str(), (p + ".dbgbf").c_str());
    if (!rc) rc = rb_filter_save(g->cbf, (p + ".cbf.desc").c_str(), (p + ".cbf").c_str());
    if (!rc && g->rpk) rc = rb_filter_save(g->rpk, (p + ".rpkbf.desc").c_str(), (p + ".rpkbf").c_str());
    if (!rc && g->fpk) rc = rb_filter_save(g->fpk, (p + ".fpkbf.desc").c_str(), (p + ".fpkbf").c_str());
    return rc;
}
static bool file_exists(const std::string& p) { FILE* f = fopen(p.c_str(), "rb"); if (f) fclose(f); return f != nullptr; }
extern "C" int32_t rb_graph_load(rb_ctx* ctx, const char* path, int32_t load_dbgbf, int32_t load_fpkbf, rb_graph** out) {
    if (!ctx || !path || !out) return RB_EINVAL;
    LOCK(ctx);
    std::vector<std::pair<std::string, std::string>> kv;
    int32_t rc = read_desc(ctx, path, &kv);
    if (rc) return rc;
    rb_graph* g = new rb_graph();
    memset(g, 0, sizeof *g);
    g->ctx = ctx; g->d_read = -1; g->d_frag = -1;
    for (auto& e : kv) {
        if (e.first == "dbgbfCbfMaxNumHash") g->hmax = atoi(e.second.c_str());
        else if (e.first == "stranded") g->stranded = e.second == "true";
        else if (e.first == "k") g->k = atoi(e.second.c_str());
        else if (e.first == "readPairedKmersDistance") g->d_read = atoi(e.second.c_str());
        else if (e.first == "fragmentPairedKmersDistance") g->d_frag = atoi(e.second.c_str());
    }
    const std::string p(path);
    rc = rb_filter_load(ctx, RB_BLOOM, (p + ".dbgbf.desc").c_str(), (p + ".dbgbf").c_str(), g->k, load_dbgbf, &g->dbg);
    if (!rc) rc = rb_filter_load(ctx, RB_COUNTING, (p + ".cbf.desc").c_str(), (p + ".cbf").c_str(), g->k, 1, &g->cbf);
    if (!rc && file_exists(p + ".rpkbf.desc")) rc = rb_filter_load(ctx, RB_BLOOM, (p + ".rpkbf.desc").c_str(), (p + ".rpkbf").c_str(), g->k, 1, &g->rpk);
    if (!rc && load_fpkbf && file_exists(p + ".fpkbf.desc"))
        rc = rb_filter_load(ctx, RB_BLOOM, (p + ".fpkbf.desc").c_str(), (p + ".fpkbf").c_str(), g->k, 1, &g->fpk);
    if (rc) {
        if (g->dbg) filter_free(g->dbg);
        if (g->cbf) filter_free(g->cbf);
        if (g->rpk) filter_free(g->rpk);
        if (g->fpk) filter_free(g->fpk);
        delete g;
        return rc;
    }
    g->hd = g->dbg->num_hash; g->hc = g->cbf->num_hash; g->hp = g->rpk ? g->rpk->num_hash : (g->fpk ? g->fpk->num_hash : 0);
    // the file's dbgbfCbfMaxNumHash is max(h_d, h_c) by construction (graph :83); a file that says otherwise is inconsistent
    if (g->hmax != 0 && g->hmax != std::max(g->hd, g->hc)) {
        filter_free(g->dbg); filter_free(g->cbf);
        if (g->rpk) filter_free(g->rpk);
        if (g->fpk) filter_free(g->fpk);
        delete g;
        return fail(ctx, RB_EIO, "graph desc: dbgbfCbfMaxNumHash does not match the filters' numhash");
    }
    g->hmax = std::max(g->hd, g->hc);
    g->engine = engine_from_env();
    g->dbg->in_graph = g->cbf->in_graph = true;
    if (g->rpk) g->rpk->in_graph = true;
    if (g->fpk) g->fpk->in_graph = true;
    *out = g;
    return RB_OK;
}

// ---- host mirror: per-k-mer Java calls (GraphUtils: getCount, contains, neighbour iterators ...) keep running on the host
// buffers of the inherited BloomFilter / CountingBloomFilter objects, which are byte-identical to the device arrays after this call
extern "C" int32_t rb_graph_sync(rb_graph* g) {
    if (!g) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    return RB_OK;
}
extern "C" int32_t rb_graph_sync_to_host(rb_graph* g, void* dbgbf, void* cbf, void* rpkbf, void* fpkbf) {
    if (!g) return RB_EINVAL;
    rb_ctx* ctx = g->ctx;
    LOCK(ctx);
    if ((rpkbf && !g->rpk) || (fpkbf && !g->fpk)) return fail(ctx, RB_ESTATE, "sync_to_host: the graph has no such pair filter");
    { const int32_t rc = cells_ensure_logical(ctx, g->dbg->cs); if (rc) return rc; }
    // the four copies are queued back to back (pageable destinations are staged by the driver; pinned ones run at PCIe speed)
    if (dbgbf) CK(cudaMemcpyAsync(dbgbf, g->dbg->dev, (size_t)g->dbg->nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (cbf) CK(cudaMemcpyAsync(cbf, g->cbf->dev, (size_t)g->cbf->nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (rpkbf) CK(cudaMemcpyAsync(rpkbf, g->rpk->dev, (size_t)g->rpk->nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (fpkbf) CK(cudaMemcpyAsync(fpkbf, g->fpk->dev, (size_t)g->fpk->nbytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    return RB_OK;
}

// ---- synthetic workload ----------------------------------------------------------------------------------------------------
extern "C" int32_t rb_synth_reads_dev(rb_ctx* ctx, uint64_t seed, uint64_t genome_len, uint64_t first_read, int64_t n_reads, int32_t L,
                                      uint32_t err_ppm, int64_t stride_bases, uint64_t* packed_dev) {
    if (!ctx || !packed_dev || n_reads < 0 || L < 1 || stride_bases < L || (stride_bases & 31) || genome_len < (uint64_t)L) return RB_EINVAL;
    LOCK(ctx);
    const int64_t threads = n_reads * (stride_bases >> 5);
    if (threads == 0) return RB_OK;
    RB_LAUNCH((int)div_up(threads, kThreads), kThreads, 0, ctx->stream, k_synth_reads)(seed, genome_len, first_read, n_reads, L, err_ppm, stride_bases, packed_dev);
    LAUNCH_CHECK();
    return RB_OK;
}

extern "C" int32_t rb_synth_long_read_len(uint64_t seed, uint64_t read) { return synth_long_len(seed, read); }
extern "C" int32_t rb_synth_long_reads_dev(rb_ctx* ctx, uint64_t seed, uint64_t genome_len, uint64_t first_read, int64_t n_reads, uint32_t sub_ppm,
                                           uint32_t ins_ppm, uint32_t del_ppm, const int64_t* read_off_dev, uint64_t* packed_dev) {
    if (!ctx || !packed_dev || !read_off_dev || n_reads < 0 || genome_len < 16384) return RB_EINVAL;
    LOCK(ctx);
    if (n_reads == 0) return RB_OK;
    RB_LAUNCH((int)div_up(n_reads, kThreads), kThreads, 0, ctx->stream, k_synth_long_reads)(seed, genome_len, first_read, n_reads, sub_ppm, ins_ppm, del_ppm,
                                                                                        read_off_dev, packed_dev);
    LAUNCH_CHECK();
    return RB_OK;
}

#include "rb_sliced_host.inl"
#include "rb_mgraph_host.inl"

