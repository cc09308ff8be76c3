// rb_mgraph_host.inl -- the hash-sharded graph: one process per GPU, filters split by index range, exchanges owned by the library.
// Included by rnabloom_gpu.cu (kernels: rb_sliced.cuh).  SURVEY.md section 8e; graph/BloomFilterDeBruijnGraph.java:75-104,405-412,562-570.
//
// The tile sort of the sliced engine *is* the routing step: a producer sorts its records by (owner rank, filter slice of the owner), so
// the regions of one destination rank are one contiguous piece of the send arena and every exchange is an all-to-all with equal split
// sizes (regions travel with their capacity, their counts beside them).  The owner consumes the regions it received slice by slice (all
// sources of a slice together, so the slice stays L2-resident), writes one answer byte per probe at the probe's own position, and the
// answers go back with the mirror-image all-to-all: they land exactly where the producer's tile metadata points.  Per round and rank:
//   lookup: route -> | probes -> apply -> | answers back -> combine
//   insert: keys -> | (home rank of the key's hash range) -> split + dedup -> probes -> | -> apply (test-and-set) -> | answers back ->
//           combine (raise bytes over the answers) -> | raise bytes to the owners -> apply raises     ( | = one all-to-all )
// The raise bytes travel like the probes did (same regions, one byte per record) to the owner, which still holds the probe records.
// With paired probe records (cbf_bytes = 2^c dividing dbg_bits, h_d >= h_c) a rank owns whole paired slices: its counters are a
// contiguous range of the cbf, its bits are that range inside every chunk of cbf_bytes bits (rb_mgraph_layout tells how to reassemble).
//
// The exchange is a transport of two calls (rb_transport in the header): the built-in one drives NCCL (dlopen of libnccl.so.2:
// ncclAlltoAll where the library has it, grouped ncclSend / ncclRecv otherwise) on the context's stream; a host can plug in its own
// (the CPU tests run the same orchestrator over gloo with the emulated kernels).  No host synchronisation inside a round: region
// overflows raise a device flag that is max-reduced over the ranks and makes the kernels that would modify filters return at once; the
// host reads the flags once, after the round.

#ifndef RB_EMU
#include <dlfcn.h>
#endif

// ---- NCCL, bound at run time (no link-time dependency: a single-GPU host never needs it) ---------------------------------------------------
struct NcclUid { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int version = 0;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(NcclUid*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*AlltoAll)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;   // NCCL >= 2.28
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
};
enum { kNcclUint8 = 1, kNcclInt32 = 2, kNcclMax = 2 };
static NcclApi g_nccl;
static std::mutex g_nccl_mu;
static bool nccl_load(std::string* why) {
#ifdef RB_EMU
    *why = "NCCL is not available in the host emulation";
    return false;
#else
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return true;
    const char* names[] = {getenv("RB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
    if (!h) { *why = std::string("cannot load libnccl.so.2 (set RB_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return false; }
    NcclApi a;
    a.lib = h;
#define NCCL_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
    NCCL_SYM(GetVersion, "ncclGetVersion"); NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); NCCL_SYM(CommInitRank, "ncclCommInitRank");
    NCCL_SYM(CommDestroy, "ncclCommDestroy"); NCCL_SYM(GetErrorString, "ncclGetErrorString"); NCCL_SYM(AlltoAll, "ncclAlltoAll");
    NCCL_SYM(AllReduce, "ncclAllReduce"); NCCL_SYM(AllGather, "ncclAllGather"); NCCL_SYM(GroupStart, "ncclGroupStart"); NCCL_SYM(GroupEnd, "ncclGroupEnd");
    NCCL_SYM(Send, "ncclSend"); NCCL_SYM(Recv, "ncclRecv");
#undef NCCL_SYM
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GroupStart || !a.GroupEnd || !a.Send || !a.Recv) {
        *why = "libnccl.so.2 lacks a required symbol";
        return false;
    }
    if (a.GetVersion) a.GetVersion(&a.version);
    g_nccl = a;
    return true;
#endif
}
extern "C" int32_t rb_nccl_unique_id(void* id, int64_t len) {
    if (!id || len < 128) return RB_EINVAL;
    std::string why;
    if (!nccl_load(&why)) return fail(nullptr, RB_ENCCL, why);
    NcclUid u;
    const int r = g_nccl.GetUniqueId(&u);
    if (r != 0) return fail(nullptr, RB_ENCCL, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    memcpy(id, &u, 128);
    return RB_OK;
}
struct NcclTransport { void* comm; int W, rank; rb_ctx* ctx; };
static int32_t nccl_fail(rb_ctx* ctx, const char* what, int r) {
    return fail(ctx, RB_ENCCL, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error"));
}
static int32_t nccl_all_to_all(void* user, const void* send, void* recv, int64_t bytes_per_rank, void* stream) {
    NcclTransport* t = (NcclTransport*)user;
    cudaStream_t s = (cudaStream_t)stream;
    int r;
    if (g_nccl.AlltoAll && !getenv("RB_NCCL_SENDRECV")) {
        r = g_nccl.AlltoAll(send, recv, (size_t)bytes_per_rank, kNcclUint8, t->comm, s);
        return r ? nccl_fail(t->ctx, "ncclAlltoAll", r) : RB_OK;
    }
    if ((r = g_nccl.GroupStart())) return nccl_fail(t->ctx, "ncclGroupStart", r);
    for (int p = 0; p < t->W; ++p) {
        if ((r = g_nccl.Send((const char*)send + (size_t)p * bytes_per_rank, (size_t)bytes_per_rank, kNcclUint8, p, t->comm, s))) return nccl_fail(t->ctx, "ncclSend", r);
        if ((r = g_nccl.Recv((char*)recv + (size_t)p * bytes_per_rank, (size_t)bytes_per_rank, kNcclUint8, p, t->comm, s))) return nccl_fail(t->ctx, "ncclRecv", r);
    }
    if ((r = g_nccl.GroupEnd())) return nccl_fail(t->ctx, "ncclGroupEnd", r);
    return RB_OK;
}
static int32_t nccl_all_reduce_max(void* user, int32_t* buf, int64_t n, void* stream) {
    NcclTransport* t = (NcclTransport*)user;
    const int r = g_nccl.AllReduce(buf, buf, (size_t)n, kNcclInt32, kNcclMax, t->comm, (cudaStream_t)stream);
    return r ? nccl_fail(t->ctx, "ncclAllReduce", r) : RB_OK;
}

static int32_t nccl_all_gather(void* user, const void* send, void* recv, int64_t bytes, void* stream) {
    NcclTransport* t = (NcclTransport*)user;
    if (!g_nccl.AllGather) return fail(t->ctx, RB_ENCCL, "libnccl has no ncclAllGather");
    const int r = g_nccl.AllGather(send, recv, (size_t)bytes, kNcclUint8, t->comm, (cudaStream_t)stream);
    return r ? nccl_fail(t->ctx, "ncclAllGather", r) : RB_OK;
}

// ---- the sharded graph --------------------------------------------------------------------------------------------------------------------------
struct rb_mgraph {
    rb_ctx* ctx;
    int W, rank, hd, hc, k, stranded;
    int64_t dbg_bits, cbf_bytes;      // global sizes
    SlGeom sg_route, sg_apply;        // producer view (global regions) / consumer view (local regions, region_div = W)
    bool paired;
    int R, KR;                        // per rank: probe regions, key ranges
    int lg1, sub_bits;
    int64_t n_max, n_dense;           // instances per rank and round; capacity of the dense distinct-key arrays
    uint32_t probe_cap, key_cap, sub_cap;
    rb_filter *dbg, *cbf;             // local shares
    unsigned int *probe_cursor, *key_cursor, *cons_cursor, *sub_cursor, *n_distinct;
    uint32_t *cons_rlo, *pos;
    uint2* tile_meta;
    unsigned long long *sub_data, *dkey;
    unsigned int* dmult;
    int* chunk_prefix;
    int* flags;                       // device: [0] a region overflowed before anything was modified, [2] never set (barrier word)
    int64_t n_items;                  // instances of the last route_lookup
    bool lookup_fast;                 // which route kernel (and so which combine mapping) the last route_lookup used
    // exchange
    rb_transport tr;
    NcclTransport nccl;               // tr.user when the built-in transport is used
    bool own_comm;
    uint32_t *send32, *recv32, *cnt_s, *cnt_r;
    unsigned long long *send64, *recv64;
    uint8_t *ans, *home_ans;
    int64_t n_probe, n_key;           // records of one send buffer (all destinations)
    int64_t exchanged_bytes, rounds;
    // peer-to-peer mode (GPUs of one box, CUDA IPC): a consumer kernel reads the producers' send arenas directly over NVLink and writes
    // its answers straight into the producers' answer arrays -- the transfer overlaps the filter work tile by tile and no separate
    // exchange step (nor a receive buffer) exists; the transport is only used for barriers / flag agreement
    bool p2p;
    uint32_t* cnt_k;                  // packed key-range counts (its own array: peers may still read it while the probes are counted)
    void** peer_open;                 // host: W * 5 pointers opened with cudaIpcOpenMemHandle (nullptr for this rank's own)
    void **d_peer32, **d_peer64, **d_peer_ans, **d_peer_cnt, **d_peer_cntk;   // device tables [W]
};

extern "C" int32_t rb_mgraph_destroy(rb_mgraph* mg) {
    if (!mg) return RB_EINVAL;
    rb_ctx* ctx = mg->ctx;
    LOCK(ctx);
    cudaStreamSynchronize(ctx->stream);
#ifndef RB_EMU
    if (mg->own_comm && mg->nccl.comm) g_nccl.CommDestroy(mg->nccl.comm);
#endif
    if (mg->dbg && mg->dbg->cs) cells_destroy(mg->dbg->cs);
    if (mg->dbg) filter_free(mg->dbg);
    if (mg->cbf) filter_free(mg->cbf);
    cudaFree(mg->probe_cursor); cudaFree(mg->key_cursor); cudaFree(mg->cons_cursor); cudaFree(mg->sub_cursor);
    cudaFree(mg->n_distinct); cudaFree(mg->cons_rlo); cudaFree(mg->pos); cudaFree(mg->tile_meta); cudaFree(mg->sub_data); cudaFree(mg->dkey);
    cudaFree(mg->dmult); cudaFree(mg->chunk_prefix); cudaFree(mg->flags);
#ifndef RB_EMU
    if (mg->peer_open) {
        for (int i = 0; i < mg->W * 5; ++i) if (mg->peer_open[i]) cudaIpcCloseMemHandle(mg->peer_open[i]);
        free(mg->peer_open);
    }
#endif
    cudaFree(mg->d_peer32); cudaFree(mg->d_peer64); cudaFree(mg->d_peer_ans); cudaFree(mg->d_peer_cnt); cudaFree(mg->d_peer_cntk);
    cudaFree(mg->send32); cudaFree(mg->send64); cudaFree(mg->home_ans); cudaFree(mg->cnt_s); cudaFree(mg->cnt_k);
    if (mg->W > 1 && !mg->p2p) { cudaFree(mg->recv32); cudaFree(mg->recv64); cudaFree(mg->ans); cudaFree(mg->cnt_r); }
    delete mg;
    return RB_OK;
}

// Peer-to-peer mode: every rank exports its send arenas, its answer array and its count arrays with CUDA IPC, the handles travel through
// the transport's all_gather, every rank maps the others'.  All ranks agree (max-reduce of a failure flag) so that a box without
// peer access falls back to the staged exchange everywhere.  RB_MGRAPH_P2P=0 forces the staged exchange.
static int32_t mg_setup_p2p(rb_mgraph* mg) {
    mg->p2p = false;
#ifndef RB_EMU
    rb_ctx* ctx = mg->ctx;
    const int W = mg->W;
    if (W == 1 || !mg->tr.all_gather || !env_int("RB_MGRAPH_P2P", 1, 0, 1)) return RB_OK;
    void* mine[5] = {mg->send32, mg->send64, mg->home_ans, mg->cnt_s, mg->cnt_k};
    std::vector<cudaIpcMemHandle_t> hs(5), all((size_t)W * 5);
    int failed = 0;
    for (int i = 0; i < 5; ++i) if (cudaIpcGetMemHandle(&hs[i], mine[i]) != cudaSuccess) failed = 1;
    cudaGetLastError();
    void *d_h = nullptr, *d_all = nullptr;
    int* d_flag = nullptr;
    const size_t hb = sizeof(cudaIpcMemHandle_t) * 5;
    CK(cudaMalloc(&d_h, hb)); CK(cudaMalloc(&d_all, hb * W)); CK(cudaMalloc(&d_flag, 64));
    CK(cudaMemcpyAsync(d_h, hs.data(), hb, cudaMemcpyHostToDevice, ctx->stream));
    int32_t rc = mg->tr.all_gather(mg->tr.user, d_h, d_all, (int64_t)hb, ctx->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(all.data(), d_all, hb * W, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    mg->peer_open = (void**)calloc((size_t)W * 5, sizeof(void*));
    std::vector<void*> tab((size_t)W * 5, nullptr);
    for (int p = 0; p < W && !failed; ++p)
        for (int i = 0; i < 5; ++i) {
            if (p == mg->rank) { tab[(size_t)p * 5 + i] = mine[i]; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[(size_t)p * 5 + i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { failed = 1; cudaGetLastError(); break; }
            mg->peer_open[(size_t)p * 5 + i] = ptr;
            tab[(size_t)p * 5 + i] = ptr;
        }
    CK(cudaMemcpyAsync(d_flag, &failed, 4, cudaMemcpyHostToDevice, ctx->stream));
    rc = mg->tr.all_reduce_max(mg->tr.user, d_flag, 1, ctx->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(&failed, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_h); cudaFree(d_all); cudaFree(d_flag);
    if (failed) {   // somewhere a mapping failed: nobody uses peer memory
        for (int i = 0; i < W * 5; ++i) if (mg->peer_open[i]) { cudaIpcCloseMemHandle(mg->peer_open[i]); mg->peer_open[i] = nullptr; }
        return RB_OK;
    }
    void*** dst[5] = {&mg->d_peer32, &mg->d_peer64, &mg->d_peer_ans, &mg->d_peer_cnt, &mg->d_peer_cntk};
    for (int i = 0; i < 5; ++i) {
        std::vector<void*> col((size_t)W);
        for (int p = 0; p < W; ++p) col[(size_t)p] = tab[(size_t)p * 5 + i];
        CK(cudaMalloc(dst[i], sizeof(void*) * W));
        CK(cudaMemcpyAsync(*dst[i], col.data(), sizeof(void*) * W, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    mg->p2p = true;
#endif
    return RB_OK;
}

static int32_t mgraph_create(rb_ctx* ctx, int32_t n_ranks, int32_t rank, const rb_transport* tr, const void* nccl_id, int64_t dbg_bits, int64_t cbf_bytes,
                             int32_t hd, int32_t hc, int32_t k, int32_t stranded, int64_t max_kmers, rb_mgraph** out) {
    if (!ctx || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks || max_kmers < 1) return RB_EINVAL;
    if ((n_ranks & (n_ranks - 1)) != 0) return fail(ctx, RB_EINVAL, "mgraph: the number of ranks must be a power of two");
    if (hd < 1 || hd > kSlMaxH || hc < 1 || hc > kSlMaxH || dbg_bits < 1 || cbf_bytes < 1) return fail(ctx, RB_EINVAL, "mgraph: needs 1..3 hashes per filter");
    if (n_ranks > 1 && !tr && !nccl_id) return fail(ctx, RB_EINVAL, "mgraph: more than one rank needs a transport or an NCCL unique id");
    rb_mgraph* mg = new rb_mgraph();
    memset(mg, 0, sizeof *mg);
    mg->ctx = ctx; mg->W = n_ranks; mg->rank = rank; mg->hd = hd; mg->hc = hc; mg->k = k; mg->stranded = stranded ? 1 : 0;
    mg->dbg_bits = dbg_bits; mg->cbf_bytes = cbf_bytes;
    const int W = n_ranks;
    SlGeom sg;
    memset(&sg, 0, sizeof sg);
    sg.dbg_fm = make_fm(dbg_bits); sg.cbf_fm = make_fm(cbf_bytes);
    sg.hd = hd; sg.hc = hc;
    mg->paired = sl_pair_geometry(dbg_bits, cbf_bytes, hd, hc, W, &sg);
    sg.dbg_log2 = env_int("RB_SLICE_BITS_LOG2", 29, 5, 31);
    sg.cbf_log2 = env_int("RB_SLICE_BYTES_LOG2", 26, 2, 31);
    int64_t tot_d, tot_c;
    for (;;) {   // slices of every rank: all regions of a producer must fit the tile sort's bucket range
        tot_d = div_up(dbg_bits, 1LL << sg.dbg_log2); tot_c = div_up(cbf_bytes, 1LL << sg.cbf_log2);
        if ((div_up(tot_d, W) + div_up(tot_c, W)) * W <= kSlMaxRegions) break;
        if (tot_d >= tot_c && sg.dbg_log2 < 31) ++sg.dbg_log2; else if (sg.cbf_log2 < 31) ++sg.cbf_log2; else break;
    }
    sg.shard_d = (int)div_up(tot_d, W); sg.shard_c = (int)div_up(tot_c, W);
    if (mg->paired) {   // the counters of a rank are its paired slices
        sg.shard_p = sg.n_pair / W;                                   // also for one rank (sl_pair_geometry leaves it 0 there)
        sg.pair_local_c = (uint64_t)sg.shard_p << sg.pair_log2;
        sg.cbf_log2 = sg.pair_log2;
        sg.shard_c = sg.shard_p;
        sg.shard_d = 0;   // unused: a paired region is the global slice number
    }
    sg.region_div = 1;
    mg->R = mg->paired ? sg.shard_p : sg.shard_d + sg.shard_c;
    if ((int64_t)mg->R * W > kSlMaxRegions) { delete mg; return fail(ctx, RB_EINVAL, "mgraph: too many filter slices for this many ranks"); }
    if (!mg->paired) { sg.n_dbg = sg.shard_d * W; sg.n_cbf = sg.shard_c * W; }
    mg->sg_route = sg;
    mg->sg_apply = sg;
    mg->sg_apply.region_div = W;
    if (!mg->paired) { mg->sg_apply.n_dbg = sg.shard_d; mg->sg_apply.n_cbf = sg.shard_c; }
    // the shard_* fields make the region functions of the producer interleave (owner, local slice); the consumer sees local regions
    // local shares: whole slices, so that the ranks' shares reassemble to the global arrays (rb_mgraph_layout)
    int64_t local_d, local_c;
    if (mg->paired) {
        local_c = (int64_t)sg.pair_local_c;
        local_d = (dbg_bits / cbf_bytes) * local_c;
    } else {
        const int64_t share_d = (int64_t)sg.shard_d << sg.dbg_log2, share_c = (int64_t)sg.shard_c << sg.cbf_log2;
        local_d = std::max<int64_t>(0, std::min<int64_t>(share_d, dbg_bits - share_d * rank));
        local_c = std::max<int64_t>(0, std::min<int64_t>(share_c, cbf_bytes - share_c * rank));
    }
    // capacities follow the caller's round size (no power-of-two rounding: every record of slack travels through the exchanges); the
    // distinct keys a home rank can see exceed its share of the instances only by the imbalance of the hash ranges (~1e-4 at 10^8 keys)
    const int64_t n_max = div_up(std::max<int64_t>(max_kmers, 1024), 4096) * 4096;
    mg->n_max = n_max; mg->n_dense = n_max + n_max / 32 + 4096;
    const int lgSub = env_int("RB_SLICED_SUBRANGE_LOG2", 10, 4, 11);
    int lgW = 0; while ((1 << lgW) < W) ++lgW;
    int lgS = 0; while (((n_max * W) >> lgSub) > (1LL << lgS)) ++lgS;       // sub-ranges over all ranks
    mg->lg1 = std::max(lgW, std::min((lgS + 1) / 2, 11));
    mg->sub_bits = std::max(0, std::min(lgS - mg->lg1, 11));
    mg->KR = (1 << mg->lg1) / W;
    const double keys_per_sub = (double)n_max * W / (double)(1LL << (mg->lg1 + mg->sub_bits));
    mg->sub_cap = (uint32_t)sl_capacity(keys_per_sub);
    mg->key_cap = (uint32_t)sl_capacity((double)n_max / (double)(1 << mg->lg1));
    if (mg->paired) {
        mg->probe_cap = (uint32_t)sl_capacity((double)mg->n_dense * hd / (double)sg.n_pair);
    } else {
        const double slices_d = std::max(1.0, (double)dbg_bits / (double)(1LL << sg.dbg_log2)), slices_c = std::max(1.0, (double)cbf_bytes / (double)(1LL << sg.cbf_log2));
        mg->probe_cap = (uint32_t)sl_capacity(std::max((double)mg->n_dense * hd / slices_d, (double)mg->n_dense * hc / slices_c));
    }
    const int64_t n_sub_regions = (int64_t)mg->KR << mg->sub_bits;
    if (mg->sub_cap >= (uint32_t)kSlDedupSlots || (int64_t)mg->R * W * mg->probe_cap >= (1LL << 32) - (1LL << 20) ||
        n_sub_regions * mg->sub_cap >= (1LL << 32) - (1LL << 20)) {
        delete mg;
        return fail(ctx, RB_EINVAL, "mgraph: max_kmers_per_round too large for 32-bit record positions");
    }
    mg->n_probe = (int64_t)W * mg->R * mg->probe_cap;
    mg->n_key = (int64_t)W * mg->KR * mg->key_cap;
    int32_t rc = filter_alloc(ctx, RB_BLOOM, std::max<int64_t>(local_d, 32), hd, k, &mg->dbg);
    if (!rc) rc = filter_alloc(ctx, RB_COUNTING, std::max<int64_t>(local_c, 4), hc, k, &mg->cbf);
    cudaError_t er = cudaSuccess;
    if (!rc) {
        const int nj = mg->paired ? 3 : kSlNJ;
        const int maxB = std::max(mg->R, mg->KR) * W;
        const int64_t n_tiles = mg->n_dense / (mg->paired ? SlShape<3>::TILE : SlShape<6>::TILE) + 8;
        const size_t n32 = (size_t)mg->n_probe + kSlSpill;
        er = cudaMalloc(&mg->probe_cursor, (size_t)mg->R * W * kSlPad * 4);
        if (er == cudaSuccess) er = cudaMalloc(&mg->key_cursor, (size_t)mg->KR * W * kSlPad * 4);
        if (er == cudaSuccess) er = cudaMalloc(&mg->cons_cursor, (size_t)maxB * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&mg->cons_rlo, (size_t)maxB * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&mg->sub_cursor, (size_t)n_sub_regions * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&mg->sub_data, ((size_t)n_sub_regions * mg->sub_cap + kSlSpill) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&mg->n_distinct, 64);
        if (er == cudaSuccess) er = cudaMalloc(&mg->pos, ((size_t)mg->n_dense + 8) * nj * 4);
        if (er == cudaSuccess) er = cudaMalloc(&mg->tile_meta, (size_t)n_tiles * ((size_t)mg->R * W + 1) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&mg->dkey, ((size_t)mg->n_dense + 8) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&mg->dmult, ((size_t)mg->n_dense + 8) * 4);
        if (er == cudaSuccess) er = cudaMalloc(&mg->chunk_prefix, (size_t)(maxB + 2) * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&mg->flags, 64);
        if (er == cudaSuccess) er = cudaMemsetAsync(mg->flags, 0, 64, ctx->stream);
        if (er == cudaSuccess) er = cudaMalloc(&mg->send32, n32 * 4);
        if (er == cudaSuccess) er = cudaMalloc(&mg->send64, ((size_t)mg->n_key + kSlSpill) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&mg->home_ans, (size_t)mg->n_probe + kSlSpill);
        if (er == cudaSuccess) er = cudaMalloc(&mg->cnt_s, (size_t)maxB * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&mg->cnt_k, (size_t)maxB * 4 + 64);
        mg->recv32 = mg->send32; mg->recv64 = mg->send64; mg->ans = mg->home_ans; mg->cnt_r = mg->cnt_s;   // W == 1, and until decided otherwise
    }
    if (rc || er != cudaSuccess) {
        if (!rc) rc = fail(ctx, RB_ENOMEM, std::string("mgraph alloc: ") + cudaGetErrorString(er));
        rb_mgraph_destroy(mg);
        return rc;
    }
    mg->dbg->in_graph = mg->cbf->in_graph = true;
    // paired slices with at most 8 chunks: the owner works on co-located cells (SlGeom::cells); rb_mgraph_filter hands out the logical
    // arrays, converted back on demand
    if (mg->paired && cells_create(ctx, mg->dbg, mg->cbf)) mg->sg_apply.cells = 1;
    const int maxB_all = std::max(mg->R, mg->KR) * W;
    const size_t n32_all = (size_t)mg->n_probe + kSlSpill;
    if (tr) mg->tr = *tr;
    else if (W > 1) {
#ifdef RB_EMU
        rb_mgraph_destroy(mg);
        return fail(ctx, RB_ENCCL, "NCCL is not available in the host emulation");
#else
        std::string why;
        if (!nccl_load(&why)) { rb_mgraph_destroy(mg); return fail(ctx, RB_ENCCL, why); }
        NcclUid u;
        memcpy(&u, nccl_id, 128);
        void* comm = nullptr;
        const int r = g_nccl.CommInitRank(&comm, W, u, rank);
        if (r != 0) { rb_mgraph_destroy(mg); return nccl_fail(ctx, "ncclCommInitRank", r); }
        mg->nccl.comm = comm; mg->nccl.W = W; mg->nccl.rank = rank; mg->nccl.ctx = ctx;
        mg->own_comm = true;
        mg->tr.user = &mg->nccl; mg->tr.all_to_all = nccl_all_to_all; mg->tr.all_reduce_max = nccl_all_reduce_max; mg->tr.all_gather = nccl_all_gather;
#endif
    }
    if (W > 1) {
        rc = mg_setup_p2p(mg);
        if (rc) { rb_mgraph_destroy(mg); return rc; }
        if (!mg->p2p) {   // staged exchange: receive buffers and an owner-side answer array
            mg->recv32 = nullptr; mg->recv64 = nullptr; mg->ans = nullptr; mg->cnt_r = nullptr;   // they aliased the send side until here
            er = cudaMalloc(&mg->recv32, n32_all * 4);
            if (er == cudaSuccess) er = cudaMalloc(&mg->recv64, ((size_t)mg->n_key + kSlSpill) * 8);
            if (er == cudaSuccess) er = cudaMalloc(&mg->ans, (size_t)mg->n_probe + kSlSpill);
            if (er == cudaSuccess) er = cudaMalloc(&mg->cnt_r, (size_t)maxB_all * 4 + 64);
            if (er != cudaSuccess) { rb_mgraph_destroy(mg); return fail(ctx, RB_ENOMEM, std::string("mgraph exchange buffers: ") + cudaGetErrorString(er)); }
        }
    }
    *out = mg;
    return RB_OK;
}
extern "C" int32_t rb_mgraph_create_nccl(rb_ctx* ctx, int32_t n_ranks, int32_t rank, const void* nccl_unique_id, int64_t dbg_bits, int64_t cbf_bytes,
                                         int32_t hd, int32_t hc, int32_t k, int32_t stranded, int64_t max_kmers_per_round, rb_mgraph** out) {
    if (!ctx) return RB_EINVAL;
    LOCK(ctx);
    return mgraph_create(ctx, n_ranks, rank, nullptr, nccl_unique_id, dbg_bits, cbf_bytes, hd, hc, k, stranded, max_kmers_per_round, out);
}
extern "C" int32_t rb_mgraph_create(rb_ctx* ctx, int32_t n_ranks, int32_t rank, const rb_transport* transport, int64_t dbg_bits, int64_t cbf_bytes,
                                    int32_t hd, int32_t hc, int32_t k, int32_t stranded, int64_t max_kmers_per_round, rb_mgraph** out) {
    if (!ctx || (n_ranks > 1 && (!transport || !transport->all_to_all || !transport->all_reduce_max))) return RB_EINVAL;
    LOCK(ctx);
    return mgraph_create(ctx, n_ranks, rank, transport, nullptr, dbg_bits, cbf_bytes, hd, hc, k, stranded, max_kmers_per_round, out);
}
// layout[0] paired (0/1)  [1] dbgbf bits of a full share (unpaired: rank r holds bits [r * this, ...))  [2] cbf bytes of a full share
//       [3] local dbgbf bits  [4] local cbf bytes  [5] chunks = dbg_bits / cbf_bytes (paired)  [6] max k-mers per round and rank
//       [7] bytes one insert round sends per rank  [8] bytes one lookup round sends per rank
// paired: global bit c * cbf_bytes + r * layout[2] + x  (x < layout[2])  is local bit  c * layout[2] + x  of rank r
extern "C" int32_t rb_mgraph_layout(rb_mgraph* mg, int64_t* layout) {
    if (!mg || !layout) return RB_EINVAL;
    const SlGeom& sg = mg->sg_route;
    layout[0] = mg->paired ? 1 : 0;
    layout[1] = mg->paired ? 0 : (int64_t)sg.shard_d << sg.dbg_log2;
    layout[2] = mg->paired ? (int64_t)sg.pair_local_c : (int64_t)sg.shard_c << sg.cbf_log2;
    layout[3] = mg->dbg->size; layout[4] = mg->cbf->size;
    layout[5] = mg->paired ? mg->dbg_bits / mg->cbf_bytes : 0;
    layout[6] = mg->n_max;
    const int64_t cnt = 4 * (int64_t)mg->W;
    layout[7] = mg->n_key * 8 + mg->n_probe * 6 + cnt * (mg->KR + mg->R);   // keys, probes, answers back, raise bytes
    layout[8] = mg->n_probe * 5 + cnt * mg->R;
    return RB_OK;
}
extern "C" int32_t rb_mgraph_filter(rb_mgraph* mg, int32_t which, rb_filter** out) {
    if (!mg || !out) return RB_EINVAL;
    *out = which == RB_DBGBF ? mg->dbg : which == RB_CBF ? mg->cbf : nullptr;
    return RB_OK;
}
extern "C" int32_t rb_mgraph_peer_to_peer(rb_mgraph* mg) { return mg && mg->p2p ? 1 : 0; }
extern "C" int32_t rb_mgraph_stats(rb_mgraph* mg, int64_t* exchanged_bytes, int64_t* rounds) {
    if (!mg) return RB_EINVAL;
    if (exchanged_bytes) *exchanged_bytes = mg->exchanged_bytes;
    if (rounds) *rounds = mg->rounds;
    return RB_OK;
}

// ---- phases ---------------------------------------------------------------------------------------------------------------------------------
static SlArena mg_producer(rb_mgraph* mg, void* data, unsigned int* cursor, int per_rank, uint32_t cap) {
    SlArena a = sl_arena(data, cursor, nullptr, per_rank * mg->W, sl_chunk());
    a.cap = cap;
    return a;
}
// regions received from every rank, in the consumer's order (local region first, source rank second)
// p2p mode: peer_data / peer_cnt are the device tables of the producers' arenas / packed counts (data and recv_cnt are not used)
static int32_t mg_consumer(rb_mgraph* mg, void* data, const uint32_t* recv_cnt, void** peer_data, void** peer_cnt, int per_rank, uint32_t cap, int chunk,
                           SlArena* out, int passes = 1) {
    rb_ctx* ctx = mg->ctx;
    const int n = per_rank * mg->W;
    if (mg->p2p) RB_LAUNCH((int)div_up(n, kSlThreads), kSlThreads, 0, ctx->stream, ks_order_counts_p2p)((const uint32_t* const*)peer_cnt, mg->W, mg->rank, per_rank, cap,
                                                                                                      mg->cons_cursor, mg->cons_rlo);
    else RB_LAUNCH((int)div_up(n, kSlThreads), kSlThreads, 0, ctx->stream, ks_order_counts)(recv_cnt, mg->W, per_rank, cap, mg->cons_cursor, mg->cons_rlo);
    LAUNCH_CHECK();
    SlArena a = sl_arena(data, mg->cons_cursor, nullptr, n, chunk);
    a.cap = cap; a.cursor_stride = 1; a.rlo = mg->cons_rlo;
    if (mg->p2p) { a.data = nullptr; a.peer_data = peer_data; a.n_peers = mg->W; }
    a.passes = passes;
    sl_apply_stage(&a);
    *out = a;
    RB_LAUNCH(1, kSlThreads, ((size_t)((a.B + 3) & ~3) + 296) * 4, ctx->stream, ks_chunk_prefix)(a, mg->chunk_prefix);
    LAUNCH_CHECK();
    return RB_OK;
}
static int32_t mg_pack_counts(rb_mgraph* mg, const SlArena& a, uint32_t* dense) {
    rb_ctx* ctx = mg->ctx;
    RB_LAUNCH((int)div_up(a.B, kSlThreads), kSlThreads, 0, ctx->stream, ks_pack_counts)(a.cursor, a.cursor_stride, a.cap, a.B, dense);
    LAUNCH_CHECK();
    return RB_OK;
}
// one exchange: the counts of the regions, then the regions (equal split: every rank sends `per_rank` regions of `cap` records to every rank)
static int32_t mg_barrier(rb_mgraph* mg);
static int32_t mg_exchange(rb_mgraph* mg, const void* send, void* recv, int per_rank, uint32_t cap, int rec_bytes, bool with_counts) {
    if (mg->W == 1) return RB_OK;   // recv aliases send
    if (mg->p2p) return mg_barrier(mg);   // the consumer reads the producers' arenas in place: it only has to know that they are complete
    rb_ctx* ctx = mg->ctx;
    int32_t rc;
    if (with_counts) {
        PROF("exchange");
        rc = mg->tr.all_to_all(mg->tr.user, mg->cnt_s, mg->cnt_r, (int64_t)per_rank * 4, ctx->stream);
        if (ctx->prof_pending) prof_end(ctx);
        if (rc) return rc;
        mg->exchanged_bytes += (int64_t)per_rank * 4 * mg->W;
    }
    const int64_t bytes = (int64_t)per_rank * cap * rec_bytes;
    PROF("exchange");
    rc = mg->tr.all_to_all(mg->tr.user, send, recv, bytes, ctx->stream);
    if (ctx->prof_pending) prof_end(ctx);
    if (rc) return rc;
    mg->exchanged_bytes += bytes * mg->W;
    return RB_OK;
}
static int32_t mg_agree(rb_mgraph* mg, int first, int n) {   // every rank learns whether a region overflowed anywhere
    if (mg->W == 1) return RB_OK;
    return mg->tr.all_reduce_max(mg->tr.user, mg->flags + first, n, mg->ctx->stream);
}

// flags[2] is never set: its max-reduction is a barrier on the streams of all ranks (the collective completes on a rank only after every
// rank has reached it in stream order)
static int32_t mg_barrier(rb_mgraph* mg) {
    rb_ctx* ctx = mg->ctx;
    PROF("barrier");
    const int32_t rc = mg->tr.all_reduce_max(mg->tr.user, mg->flags + 2, 1, ctx->stream);
    if (ctx->prof_pending) prof_end(ctx);
    return rc;
}
struct MgRouteUser { rb_mgraph* mg; int mode; bool lookup; int64_t *fh, *rh; int launches; };
template <int NJ>
static int32_t mg_route_lookup_t(rb_mgraph* mg, const Ingest& ing, int mode, int64_t* fh, int64_t* rh) {
    rb_ctx* ctx = mg->ctx;
    int32_t rc;
    constexpr int TILE = SlShape<NJ>::TILE, KPT = SlShape<NJ>::KPT;
    const HashMults hm = make_hm(mg->k);
    const SlArena probes = mg_producer(mg, mg->send32, mg->probe_cursor, mg->R, mg->probe_cap);
    CK(cudaMemsetAsync(probes.cursor, 0, (size_t)probes.B * kSlPad * 4, ctx->stream));
    const size_t sm_sort = TileSort<uint32_t, KPT * NJ>::smem_bytes(probes.B);
    const bool fast = sl_uniform_fast_probes<NJ>(ing, mg->k);
    mg->n_items = ing.n_pos; mg->lookup_fast = fast;
    if (fast) {
        const int grid = (int)div_up(ing.n_pos, (int64_t)TILE);
        const size_t sm = std::max(sm_sort, PrefixKmerizer<SlShape<NJ>::PFX_PER, TILE>::smem_bytes());
        auto kf = ks_route_lookup_u<0, NJ>; auto kc = ks_route_lookup_u<2, NJ>;
        if (mode == RB_MODE_FWD) SL_LAUNCH("ks_route_lookup_u<0>", kf, grid, sm, ing, mg->k, hm, mg->sg_route, probes, mg->pos, mg->tile_meta, fh, rh, mg->flags);
        else SL_LAUNCH("ks_route_lookup_u<2>", kc, grid, sm, ing, mg->k, hm, mg->sg_route, probes, mg->pos, mg->tile_meta, fh, rh, mg->flags);
    } else {
        const int grid = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kChunk);
        auto kf = ks_route_lookup<0, NJ>; auto kc = ks_route_lookup<2, NJ>;
        if (mode == RB_MODE_FWD) SL_LAUNCH("ks_route_lookup<0>", kf, grid, sm_sort, ing, mg->k, hm, mg->sg_route, probes, mg->pos, mg->tile_meta, fh, rh, mg->flags);
        else SL_LAUNCH("ks_route_lookup<2>", kc, grid, sm_sort, ing, mg->k, hm, mg->sg_route, probes, mg->pos, mg->tile_meta, fh, rh, mg->flags);
    }
    return RB_OK;
}
static int32_t mg_route_launch(rb_ctx* ctx, const Ingest& ing_in, void* user) {
    MgRouteUser* u = (MgRouteUser*)user;
    rb_mgraph* mg = u->mg;
    if (++u->launches > 1 || ing_in.n_pos > mg->n_max) return fail(ctx, RB_EINVAL, "mgraph: the reads of one round exceed max_kmers_per_round");
    Ingest ing = ing_in;
    ing.out_base = 0;
    int32_t rc;
    if (u->lookup) return mg->paired ? mg_route_lookup_t<3>(mg, ing, u->mode, u->fh, u->rh) : mg_route_lookup_t<6>(mg, ing, u->mode, u->fh, u->rh);
    const SlArena keys = mg_producer(mg, mg->send64, mg->key_cursor, mg->KR, mg->key_cap);
    CK(cudaMemsetAsync(keys.cursor, 0, (size_t)keys.B * kSlPad * 4, ctx->stream));
    const int n_ranges = 1 << mg->lg1, shift = 64 - mg->lg1;
    if (sl_uniform_fast_keys(ing, mg->k)) {
        const int grid = (int)div_up(ing.n_pos, (int64_t)kKeyTile);
        const size_t sm = std::max(TileSort<unsigned long long, kKeyE, true>::smem_bytes(keys.B), KeyKmerizer::smem_bytes());
        if (u->mode == RB_MODE_FWD) SL_LAUNCH("ks_route_keys_u<0>", ks_route_keys_u<0>, grid, sm, ing, mg->k, n_ranges, shift, keys, mg->flags);
        else if (u->mode == RB_MODE_RC) SL_LAUNCH("ks_route_keys_u<1>", ks_route_keys_u<1>, grid, sm, ing, mg->k, n_ranges, shift, keys, mg->flags);
        else SL_LAUNCH("ks_route_keys_u<2>", ks_route_keys_u<2>, grid, sm, ing, mg->k, n_ranges, shift, keys, mg->flags);
    } else {
        const int grid = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kChunk);
        const size_t sm = TileSort<unsigned long long, kChunk, true>::smem_bytes(keys.B);
        if (u->mode == RB_MODE_FWD) SL_LAUNCH("ks_route_keys<0>", ks_route_keys<0>, grid, sm, ing, mg->k, n_ranges, shift, keys, mg->flags);
        else if (u->mode == RB_MODE_RC) SL_LAUNCH("ks_route_keys<1>", ks_route_keys<1>, grid, sm, ing, mg->k, n_ranges, shift, keys, mg->flags);
        else SL_LAUNCH("ks_route_keys<2>", ks_route_keys<2>, grid, sm, ing, mg->k, n_ranges, shift, keys, mg->flags);
    }
    return RB_OK;
}
// route this rank's reads of one round (a single launch) and pack the region counts into cnt_s
static int32_t mg_route(rb_mgraph* mg, const ReadsArg& ra, int mode, bool lookup, int64_t* fh, int64_t* rh, int64_t* n_out) {
    rb_ctx* ctx = mg->ctx;
    MgRouteUser u{mg, mode, lookup, fh, rh, 0};
    const int64_t keep = ctx->subbatch_kmers;
    ctx->subbatch_kmers = INT64_MAX / 4;      // one round = one launch
    int64_t n = 0;
    const int32_t rc = for_each_launch(ctx, ra, mg->k, mg_route_launch, &u, &n);
    ctx->subbatch_kmers = keep;
    if (rc) return rc;
    if (n_out) *n_out = n;
    const SlArena a = lookup ? mg_producer(mg, mg->send32, mg->probe_cursor, mg->R, mg->probe_cap) : mg_producer(mg, mg->send64, mg->key_cursor, mg->KR, mg->key_cap);
    if (u.launches == 0) {   // no k-mer at all on this rank: the exchange still happens, with empty regions
        CK(cudaMemsetAsync(a.cursor, 0, (size_t)a.B * kSlPad * 4, ctx->stream));
        if (lookup) mg->n_items = 0;
    }
    return mg_pack_counts(mg, a, (!lookup && mg->p2p) ? mg->cnt_k : mg->cnt_s);
}
// owner side: probes received from every rank -> answers at the same positions of `ans`
static int32_t mg_apply(rb_mgraph* mg, bool set_bits) {
    rb_ctx* ctx = mg->ctx;
    SlArena a;
    int32_t rc = mg_consumer(mg, mg->recv32, mg->cnt_r, mg->d_peer32, mg->d_peer_cnt, mg->R, mg->probe_cap, sl_chunk(), &a,
                             mg->paired ? 1 << mg->sg_apply.pair_sub_log2 : 1);
    if (rc) return rc;
    if (mg->p2p) a.peer_ans = (uint8_t* const*)mg->d_peer_ans;
    uint32_t *fd = mg->dbg->dev, *fc = mg->cbf->dev;
    if (mg->sg_apply.cells) {
        rc = cells_ensure_cells(ctx, mg->dbg->cs);
        if (rc) return rc;
        fd = fc = mg->dbg->cs->cells;
    }
    return set_bits ? sl_launch_apply<1>(ctx, a, mg->chunk_prefix, mg->sg_apply, fd, fc, mg->ans, (const int*)mg->flags)
                    : sl_launch_apply<0>(ctx, a, mg->chunk_prefix, mg->sg_apply, fd, fc, mg->ans, (const int*)mg->flags);
}
// home side: keys of this rank's hash ranges from every rank -> distinct keys with multiplicities -> their probes (send32 / cnt_s)
static int32_t mg_dedup_emit(rb_mgraph* mg, bool with_cbf) {
    rb_ctx* ctx = mg->ctx;
    SlArena keys;
    int32_t rc = mg_consumer(mg, mg->recv64, mg->cnt_r, mg->d_peer64, mg->d_peer_cntk, mg->KR, mg->key_cap, kSlThreads * kKeyE, &keys);
    if (rc) return rc;
    const int n_sub = 1 << mg->sub_bits;
    const int n_sub_regions = mg->KR << mg->sub_bits;
    SlArena subs = sl_arena(mg->sub_data, mg->sub_cursor, nullptr, n_sub_regions, 0);
    subs.cap = mg->sub_cap; subs.cursor_stride = 1;
    CK(cudaMemsetAsync(subs.cursor, 0, (size_t)n_sub_regions * 4, ctx->stream));
    int grid = 0;
    const size_t sm_split = TileSort<unsigned long long, kKeyE, true>::smem_bytes(n_sub) + (size_t)(keys.B + 1) * 4;
    rc = sl_stream_grid(ctx, ks_split_keys, sm_split, &grid);
    if (rc) return rc;
    SL_LAUNCH("ks_split_keys", ks_split_keys, grid, sm_split, keys, mg->chunk_prefix, mg->sub_bits, 64 - mg->lg1 - mg->sub_bits, mg->W, subs, mg->flags);
    CK(cudaMemsetAsync(mg->n_distinct, 0, 4, ctx->stream));
    const size_t sm_dedup = (size_t)kSlDedupSlots * 12;
    rc = sl_stream_grid(ctx, ks_dedup, sm_dedup, &grid);
    if (rc) return rc;
    SL_LAUNCH("ks_dedup", ks_dedup, std::min(grid, n_sub_regions), sm_dedup, subs, n_sub_regions, mg->lg1 + mg->sub_bits, mg->dkey, mg->dmult, mg->n_distinct,
              (unsigned int)mg->n_dense, mg->flags, SpillTable{nullptr, nullptr, 0, 0});
    const HashMults hm = make_hm(mg->k);
    const SlArena probes = mg_producer(mg, mg->send32, mg->probe_cursor, mg->R, mg->probe_cap);
    CK(cudaMemsetAsync(probes.cursor, 0, (size_t)probes.B * kSlPad * 4, ctx->stream));
    const size_t sm_sort = TileSort<uint32_t, kSlTileRecords>::smem_bytes(probes.B);
    const int grid_d = (int)div_up(mg->n_dense, (int64_t)(mg->paired ? SlShape<3>::TILE : SlShape<6>::TILE));
    if (mg->paired) SL_LAUNCH("ks_emit_probes", ks_emit_probes<3>, grid_d, sm_sort, mg->dkey, mg->n_distinct, hm, mg->sg_route, (int)with_cbf, probes, mg->pos, mg->tile_meta, mg->flags);
    else SL_LAUNCH("ks_emit_probes", ks_emit_probes<6>, grid_d, sm_sort, mg->dkey, mg->n_distinct, hm, mg->sg_route, (int)with_cbf, probes, mg->pos, mg->tile_meta, mg->flags);
    return mg_pack_counts(mg, probes, mg->cnt_s);
}
static int32_t mg_combine_lookup(rb_mgraph* mg, float* counts) {
    rb_ctx* ctx = mg->ctx;
    if (mg->n_items == 0) return RB_OK;
    int32_t rc;
    const int B = mg->R * mg->W;
    const size_t sm_ans = TileAnswers::smem_bytes(B, kSlThreads * kSlTileRecords);
    const int tile = mg->paired ? SlShape<3>::TILE : SlShape<6>::TILE;
    const int grid = mg->lookup_fast ? (int)div_up(mg->n_items, (int64_t)tile) : (int)div_up(mg->n_items, (int64_t)kSlThreads * kChunk);
    if (mg->paired) {
        auto k1 = ks_combine_lookup<1, 3>; auto k0 = ks_combine_lookup<0, 3>;
        if (mg->lookup_fast) SL_LAUNCH("ks_combine_lookup<1>", k1, grid, sm_ans, mg->pos, mg->tile_meta, B, mg->home_ans, mg->n_items, mg->hd, mg->hc, counts, (int64_t)0);
        else SL_LAUNCH("ks_combine_lookup<0>", k0, grid, sm_ans, mg->pos, mg->tile_meta, B, mg->home_ans, mg->n_items, mg->hd, mg->hc, counts, (int64_t)0);
    } else {
        auto k1 = ks_combine_lookup<1, 6>; auto k0 = ks_combine_lookup<0, 6>;
        if (mg->lookup_fast) SL_LAUNCH("ks_combine_lookup<1>", k1, grid, sm_ans, mg->pos, mg->tile_meta, B, mg->home_ans, mg->n_items, mg->hd, mg->hc, counts, (int64_t)0);
        else SL_LAUNCH("ks_combine_lookup<0>", k0, grid, sm_ans, mg->pos, mg->tile_meta, B, mg->home_ans, mg->n_items, mg->hd, mg->hc, counts, (int64_t)0);
    }
    return RB_OK;
}
// the raise phase: combine at the home rank (raise bytes over the answers) -> | -> the owner sweeps its probe regions again
static int32_t mg_raises(rb_mgraph* mg, int policy, uint64_t seed) {
    rb_ctx* ctx = mg->ctx;
    int32_t rc;
    const int B = mg->R * mg->W;
    const size_t sm_r = TileAnswers::smem_bytes(B, kSlThreads * kSlTileRecords);
    const int grid_d = (int)div_up(mg->n_dense, (int64_t)(mg->paired ? SlShape<3>::TILE : SlShape<6>::TILE));
    if (mg->paired) SL_LAUNCH("ks_combine_insert", ks_combine_insert<3>, grid_d, sm_r, mg->dkey, mg->dmult, mg->n_distinct, mg->pos, mg->tile_meta, B, mg->home_ans, mg->sg_route,
                              policy, seed, (const int*)mg->flags);
    else SL_LAUNCH("ks_combine_insert", ks_combine_insert<6>, grid_d, sm_r, mg->dkey, mg->dmult, mg->n_distinct, mg->pos, mg->tile_meta, B, mg->home_ans, mg->sg_route,
                   policy, seed, (const int*)mg->flags);
    // staged: the raise bytes travel to the owners like the probes did; p2p: a barrier, then the owners read them in place
    rc = mg_exchange(mg, mg->home_ans, mg->ans, mg->R, mg->probe_cap, 1, false);
    if (rc) return rc;
    SlArena a;   // the regions mg_apply consumed (their counts are still where they were)
    rc = mg_consumer(mg, mg->recv32, mg->cnt_r, mg->d_peer32, mg->d_peer_cnt, mg->R, mg->probe_cap, sl_chunk(), &a,
                     mg->paired ? 1 << mg->sg_apply.pair_sub_log2 : 1);
    if (rc) return rc;
    if (mg->p2p) a.peer_ans = (uint8_t* const*)mg->d_peer_ans;
    uint32_t* fc = mg->cbf->dev;
    if (mg->sg_apply.cells) {   // still in cells: nothing but mg_apply ran since
        rc = cells_ensure_cells(ctx, mg->dbg->cs);
        if (rc) return rc;
        fc = mg->dbg->cs->cells;
    }
    return sl_launch_raises(ctx, a, mg->chunk_prefix, mg->sg_apply, fc, (const uint8_t*)mg->ans, (const int*)mg->flags);
}
static int32_t mg_read_flags(rb_mgraph* mg, int* f2) {   // the one host synchronisation of a round
    rb_ctx* ctx = mg->ctx;
    f2[1] = 0;
    CK(cudaMemcpyAsync(f2, mg->flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (f2[0]) CK(cudaMemsetAsync(mg->flags, 0, 4, ctx->stream));
    return RB_OK;
}

// graph.add / addCountIfPresent / addDbgOnly for this rank's reads of one round; collective: every rank calls it (n_reads may be 0)
extern "C" int32_t rb_mgraph_add_round_dev(rb_mgraph* mg, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                           int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, uint32_t flags, int64_t* n_kmers_out) {
    if (!mg) return RB_EINVAL;
    rb_ctx* ctx = mg->ctx;
    LOCK(ctx);
    const int mode = !mg->stranded ? RB_MODE_CANON : ((flags & RB_REVCOMP) ? RB_MODE_RC : RB_MODE_FWD);
    const int policy = (flags & RB_DBG_ONLY) ? POLICY_DBG_ONLY : (flags & RB_ADD_COUNT_IF_PRESENT) ? POLICY_COUNT_IF_PRESENT : POLICY_ADD;
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, true};
    int32_t rc = RB_OK;
    if (mg->p2p) { rc = mg_barrier(mg); if (rc) return rc; }   // every rank is done reading this rank's arenas of the previous round
    rc = mg_route(mg, ra, mode, false, nullptr, nullptr, n_kmers_out);
    if (rc) return rc;
    rc = mg_exchange(mg, mg->send64, mg->recv64, mg->KR, mg->key_cap, 8, true);
    if (rc) return rc;
    rc = mg_dedup_emit(mg, policy != POLICY_DBG_ONLY);
    if (rc) return rc;
    rc = mg_agree(mg, 0, 1);   // p2p mode: also the barrier after which the probe arenas are complete everywhere
    if (rc) return rc;
    if (!mg->p2p) { rc = mg_exchange(mg, mg->send32, mg->recv32, mg->R, mg->probe_cap, 4, true); if (rc) return rc; }
    rc = mg_apply(mg, policy != POLICY_COUNT_IF_PRESENT);
    if (rc) return rc;
    ++mg->rounds;
    int f2[2] = {0, 0};
    if (policy != POLICY_DBG_ONLY) {
        rc = mg_exchange(mg, mg->ans, mg->home_ans, mg->R, mg->probe_cap, 1, false);
        if (rc) return rc;
        const uint64_t seed = ctx->rng_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(ctx->launches + 1);
        rc = mg_raises(mg, policy, seed);
        if (rc) return rc;
        rc = mg_read_flags(mg, f2);
        if (rc) return rc;
    } else {
        rc = mg_read_flags(mg, f2);
        if (rc) return rc;
    }
    claim_invalidate(ctx);
    if (f2[0]) return fail(ctx, RB_ESTATE, "mgraph: a region overflowed while the round was routed (skewed hashes); nothing was modified -- lower max_kmers_per_round");
    return RB_OK;
}
// graph.getKmers counts (+ hashes) for this rank's reads of one round; collective
extern "C" int32_t rb_mgraph_count_round_dev(rb_mgraph* mg, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                             int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, float* counts, int64_t* fhash, int64_t* rhash,
                                             int64_t* n_kmers_out) {
    if (!mg || !counts) return RB_EINVAL;
    rb_ctx* ctx = mg->ctx;
    LOCK(ctx);
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, true};
    int32_t rc = RB_OK;
    if (mg->p2p) { rc = mg_barrier(mg); if (rc) return rc; }
    rc = mg_route(mg, ra, mg->stranded ? RB_MODE_FWD : RB_MODE_CANON, true, fhash, mg->stranded ? nullptr : rhash, n_kmers_out);
    if (rc) return rc;
    rc = mg_agree(mg, 0, 1);
    if (rc) return rc;
    if (!mg->p2p) { rc = mg_exchange(mg, mg->send32, mg->recv32, mg->R, mg->probe_cap, 4, true); if (rc) return rc; }
    rc = mg_apply(mg, false);
    if (rc) return rc;
    rc = mg_exchange(mg, mg->ans, mg->home_ans, mg->R, mg->probe_cap, 1, false);
    if (rc) return rc;
    rc = mg_combine_lookup(mg, counts);
    if (rc) return rc;
    ++mg->rounds;
    int f2[2] = {0, 0};
    rc = mg_read_flags(mg, f2);
    if (rc) return rc;
    if (f2[0]) return fail(ctx, RB_ESTATE, "mgraph: a region overflowed while the look-up round was routed (skewed hashes) -- lower max_kmers_per_round");
    return RB_OK;
}
