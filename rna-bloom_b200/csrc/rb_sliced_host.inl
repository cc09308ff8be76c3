// rb_sliced_host.inl -- buffers, geometry and round drivers of the sliced engine (kernels: rb_sliced.cuh).  Included by rnabloom_gpu.cu.

struct SlicedEngine {
    int64_t n_max;                // k-mer instances per round the buffers hold
    SlGeom sg;
    bool unsupported;             // this graph cannot use the engine (hash counts / filter sizes): the direct engine serves it
    bool paired;                  // one record per hash serves both filters (rb_sliced.cuh SlShape<3>); else one record per probe
    // probes: 4-byte slice-local indices, one answer byte per probe, 6 remembered positions per k-mer (or distinct key)
    uint32_t* probe_data; uint8_t* ans; unsigned int* probe_cursor; uint32_t* probe_roff; int probe_B;
    uint32_t* pos;
    uint2* tile_meta;             // per tile sort of probes: where each bucket's run went (TileSort::run), B + 1 entries per tile
    // insert: keys by range, hash table, dense distinct keys (raises need no buffers: they ride the answer bytes of the probe records)
    unsigned long long* key_data; unsigned int* key_cursor; uint32_t* key_roff; int key_B; int key_shift;   // level 1: key_B ranges
    unsigned long long* sub_data; unsigned int* sub_cursor; int sub_bits; uint32_t sub_cap;                   // level 2: key_B << sub_bits sub-ranges
    unsigned long long* dkey; unsigned int* dmult; unsigned int* n_distinct;
    // heavy hitters (RB_SLICED_SPILL=1, off by default until it has been measured): keys that overflow their range / sub-range
    bool spill_on; unsigned long long* spill_keys; unsigned int* spill_cursor; uint32_t spill_cap;
    unsigned long long* htab_keys; unsigned int* htab_counts; int64_t htab_slots; int htab_shift;
    int* chunk_prefix;
    int* overflow;
};
static void sliced_engine_free(rb_graph* g) {
    SlicedEngine* e = g->se;
    if (!e) return;
    cudaStreamSynchronize(g->ctx->stream);
    cudaFree(e->probe_data); cudaFree(e->ans); cudaFree(e->probe_cursor); cudaFree(e->probe_roff); cudaFree(e->pos); cudaFree(e->tile_meta);
    cudaFree(e->key_data); cudaFree(e->key_cursor); cudaFree(e->key_roff);
    cudaFree(e->spill_keys); cudaFree(e->spill_cursor); cudaFree(e->htab_keys); cudaFree(e->htab_counts);
    cudaFree(e->sub_data); cudaFree(e->sub_cursor); cudaFree(e->dkey); cudaFree(e->dmult); cudaFree(e->n_distinct);
    cudaFree(e->chunk_prefix); cudaFree(e->overflow);
    delete e;
    g->se = nullptr;
}
static int env_int(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    const int x = atoi(v);
    return x < lo ? lo : (x > hi ? hi : x);
}
// Consumers walk an arena region by region with a window of grid * chunk records in flight; the window has to stay a small
// fraction of a region or several filter / table slices are live at once and fall out of L2.
static bool sliced_supports(const rb_graph* g) { return g->hd <= kSlMaxH && g->hc <= kSlMaxH && !(g->se && g->se->unsupported); }
static int sl_chunk() { return env_int("RB_SLICED_CHUNK", 2048, 256, 1 << 16) & ~15; }   // a multiple of 16 records: the apply kernels stage work items with 16-byte cp.async
static int sl_consumer_occ() { return env_int("RB_SLICED_CONSUMER_OCC", 8, 1, 8); }
static int64_t sl_pow2_at_least(int64_t v) { int64_t p = 1024; while (p < v) p <<= 1; return p; }
// capacity of a region that expects `expected` records from uniform hashes: 4 % + 8 sigma + a constant
static int64_t sl_capacity(double expected) { return (((int64_t)(expected * 1.04 + 8.0 * std::sqrt(expected + 1.0)) + 2048) + 15) & ~(int64_t)15; }   // multiple of 16: regions start 16-byte aligned (cp.async staging)
// Rounds are as large as the 32-bit record positions allow (one sweep of the filters is amortised over the round); look-ups whose
// results go back to host memory use smaller rounds so that the D2H copy of one round overlaps the kernels of the next.
static int64_t sliced_round_kmers(const rb_ctx* ctx, bool host_results) {
    if (ctx->subbatch_user_set) return std::min<int64_t>(ctx->subbatch_kmers, 1LL << 29);
    if (host_results) return 1LL << env_int("RB_SLICED_HOST_ROUND_LOG2", 27, 10, 29);
    return 1LL << env_int("RB_SLICED_ROUND_LOG2", 29, 10, 29);
}

// Uploads region offsets (B + 1 values) for uniform or two-kind capacities; returns the total number of records.
static int32_t sl_make_roff(rb_ctx* ctx, const std::vector<int64_t>& caps, uint32_t** d_roff, int64_t* total_out) {
    std::vector<uint32_t> roff(caps.size() + 1);
    int64_t acc = 0;
    for (size_t b = 0; b < caps.size(); ++b) { roff[b] = (uint32_t)acc; acc += caps[b]; }
    if (acc >= (1LL << 32) - (1LL << 20)) return fail(ctx, RB_EINVAL, "sliced engine: round too large for 32-bit record positions");
    roff[caps.size()] = (uint32_t)acc;
    CK(cudaMalloc(d_roff, roff.size() * 4));
    CK(cudaMemcpyAsync(*d_roff, roff.data(), roff.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // roff (host vector) goes out of scope
    *total_out = acc;
    return RB_OK;
}

// Paired records need cbf_bytes = 2^c dividing dbg_bits and h_d >= h_c (a paired record always tests a bit).  A slice is W = 2^w
// counters plus dbg_bits / cbf_bytes chunks of W bits; w is chosen so that one slice is about 64 MiB (it has to stay L2-resident
// while it is consumed).  n_ranks > 1: every rank must own the same whole number of slices.  Fills the pair_* fields of sg.
static bool sl_pair_geometry(int64_t dbg_bits, int64_t cbf_bytes, int hd, int hc, int n_ranks, SlGeom* sg) {
    sg->paired = 0; sg->pair_log2 = 0; sg->pair_sub_log2 = 0; sg->cbf_size_log2 = 0; sg->n_pair = 0; sg->pair_local_c = 0; sg->shard_p = 0;
    if (!env_int("RB_SLICED_PAIRED", 1, 0, 1)) return false;
    if (hd < hc || cbf_bytes < 4 || (cbf_bytes & (cbf_bytes - 1)) != 0 || dbg_bits % cbf_bytes != 0) return false;
    const int64_t q = dbg_bits / cbf_bytes;
    int c = 0; while ((1LL << c) < cbf_bytes) ++c;
    int w = 2;
    // 64 MiB slices when the regions are (owner, slice) pairs of a sharded graph (the tile sorts pay for every region); 32 MiB on one GPU:
    // measured (profiles/r02_notes.md) the apply kernels fetch a 64 MiB slice 2.3 times from DRAM, a 32 MiB slice stays put
    const double slice_bytes = (n_ranks > 1 ? 64.0 : 32.0) * 1024 * 1024;
    while (w < c && (double)(1LL << (w + 1)) * (1.0 + (double)q / 8.0) <= slice_bytes) ++w;
    w = env_int("RB_SLICE_PAIR_LOG2", w, 2, 31);
    if (w > c) w = c;
    // RB_SLICE_REGION_TARGET < the number of slices: widen the regions to 2^p sub-slices each, consumed in 2^p passes
    // (SlGeom::pair_sub_log2).  Measured at N = 4 / 8 (profiles/r02_notes.md): the tile sorts do get cheaper with fewer regions, but the
    // consumer streams every region's records once per pass -- over NVLink in peer-to-peer mode -- and loses far more (N = 8: 26.2 ->
    // 16.3 G k-mers/s with 4 passes), so the default keeps one pass; the mechanism stays for geometries that need it (> 2048 slices).
    const int64_t target = env_int("RB_SLICE_REGION_TARGET", kSlMaxRegions, 1, kSlMaxRegions);
    int p = 0;
    while ((cbf_bytes >> (w + p)) > target && p < 3 && w + p < c && (q - 1) < (1LL << (32 - (w + p + 1)))) ++p;
    w += p;
    while ((cbf_bytes >> w) > kSlMaxRegions && w < c) ++w;
    if ((cbf_bytes >> w) > kSlMaxRegions || w > 31 || (w < 32 && (q - 1) >= (1LL << (32 - w)))) return false;
    const int64_t n_pair = cbf_bytes >> w;
    if (n_ranks > 1 && (n_pair % n_ranks) != 0) return false;
    sg->paired = 1; sg->pair_log2 = w; sg->pair_sub_log2 = p; sg->cbf_size_log2 = c; sg->n_pair = (int)n_pair;
    sg->shard_p = n_ranks > 1 ? (int)(n_pair / n_ranks) : 0;
    sg->pair_local_c = n_ranks > 1 ? (uint64_t)sg->shard_p << w : (uint64_t)cbf_bytes;
    return true;
}

static int32_t sliced_engine_get(rb_graph* g, int64_t n_round, SlicedEngine** out) {
    rb_ctx* ctx = g->ctx;
    if (g->se && (g->se->unsupported || g->se->n_max >= n_round)) { *out = g->se; return RB_OK; }
    sliced_engine_free(g);
    SlicedEngine* e = new SlicedEngine();
    memset(e, 0, sizeof *e);
    g->se = e;
    *out = e;
    // ---- geometry ----
    SlGeom& sg = e->sg;
    sg.dbg_fm = make_fm(g->dbg->size); sg.cbf_fm = make_fm(g->cbf->size);
    sg.hd = g->hd; sg.hc = g->hc;
    sg.dbg_log2 = env_int("RB_SLICE_BITS_LOG2", 29, 5, 31);     // 64 MiB of bits
    sg.cbf_log2 = env_int("RB_SLICE_BYTES_LOG2", 26, 2, 31);    // 64 MiB of counters
    for (;;) {
        sg.n_dbg = (int)std::min<int64_t>(div_up(g->dbg->size, 1LL << sg.dbg_log2), 1 << 20);
        sg.n_cbf = (int)std::min<int64_t>(div_up(g->cbf->size, 1LL << sg.cbf_log2), 1 << 20);
        if (sg.n_dbg + sg.n_cbf <= kSlMaxRegions) break;
        if (sg.n_dbg >= sg.n_cbf && sg.dbg_log2 < 31) ++sg.dbg_log2;
        else if (sg.cbf_log2 < 31) ++sg.cbf_log2;
        else break;
    }
    sg.shard_d = sg.shard_c = 0; sg.region_div = 1;
    e->paired = sl_pair_geometry(g->dbg->size, g->cbf->size, g->hd, g->hc, 1, &sg);
    if (g->hd > kSlMaxH || g->hc > kSlMaxH || sg.n_dbg + sg.n_cbf > kSlMaxRegions) { e->unsupported = true; return RB_OK; }
    const int64_t n_max = sl_pow2_at_least(n_round);
    e->n_max = n_max;
    // duplicates of a key are found in sub-ranges of ~2^RB_SLICED_SUBRANGE_LOG2 keys (two tile sorts: key_B ranges x 2^sub_bits each)
    const int lgSub = env_int("RB_SLICED_SUBRANGE_LOG2", 10, 4, 11);
    int lgS = 0; while ((n_max >> lgSub) > (1LL << lgS)) ++lgS;
    const int lg1 = std::min((lgS + 1) / 2, 11);
    e->sub_bits = std::min(lgS - lg1, 11);
    e->key_B = 1 << lg1;
    e->key_shift = 64 - lg1;
    e->sub_cap = (uint32_t)sl_capacity((double)n_max / (double)(1LL << (lg1 + e->sub_bits)));
    if (getenv("RB_SLICED_SUBCAP")) e->sub_cap = (uint32_t)env_int("RB_SLICED_SUBCAP", (int)e->sub_cap, 8, kSlDedupSlots - 1);   // tests: force spills
    // heavy hitters (one k-mer with thousands of copies in a round: poly-A, very highly expressed transcripts): the copies that do not
    // fit their range / sub-range go to a spill list (up to half a round) and are aggregated in a global table of up to 2^26 slots;
    // a round that exceeds either is handed to the direct engine before anything is modified
    e->spill_on = env_int("RB_SLICED_SPILL", 1, 0, 1) != 0;
    e->spill_cap = (uint32_t)std::min<int64_t>(n_max / 2 + 65536, 1LL << 30);
    e->htab_slots = std::min<int64_t>(sl_pow2_at_least(2 * (int64_t)e->spill_cap), 1LL << 26);
    { int lg = 0; while ((1LL << lg) < e->htab_slots) ++lg; e->htab_shift = 64 - lg; }
    if (e->sub_cap >= (uint32_t)kSlDedupSlots) { e->unsupported = true; return RB_OK; }   // cannot happen with lgSub <= 11, n_max <= 2^29
    e->probe_B = e->paired ? sg.n_pair : sg.n_dbg + sg.n_cbf;
    sg.cells = 0;   // decided per round (sl_round_layout)
    if (e->paired && !g->cells_tried) { g->cells_tried = true; cells_create(ctx, g->dbg, g->cbf); }
    // ---- capacities ----
    const double dbg_slices = std::max(1.0, (double)g->dbg->size / (double)(1LL << sg.dbg_log2));
    const double cbf_slices = std::max(1.0, (double)g->cbf->size / (double)(1LL << sg.cbf_log2));
    std::vector<int64_t> caps;
    if (e->paired) {
        for (int b = 0; b < sg.n_pair; ++b) caps.push_back(sl_capacity((double)n_max * g->hd / (double)sg.n_pair));
    } else {
        for (int b = 0; b < sg.n_dbg; ++b) caps.push_back(sl_capacity((double)n_max * g->hd / dbg_slices));
        for (int b = 0; b < sg.n_cbf; ++b) caps.push_back(sl_capacity((double)n_max * g->hc / cbf_slices));
    }
    int64_t probe_slots = 0, key_slots = 0;
    int32_t rc = sl_make_roff(ctx, caps, &e->probe_roff, &probe_slots);
    if (rc) { sliced_engine_free(g); return rc; }
    caps.assign((size_t)e->key_B, getenv("RB_SLICED_KEYCAP") ? (int64_t)env_int("RB_SLICED_KEYCAP", 1 << 20, 8, 1 << 30) : sl_capacity((double)n_max / e->key_B));
    rc = sl_make_roff(ctx, caps, &e->key_roff, &key_slots);
    if (rc) { sliced_engine_free(g); return rc; }
    const int maxB = std::max(e->probe_B, e->key_B);
    const int nj = e->paired ? 3 : kSlNJ;
    const int64_t n_tiles = n_max / (e->paired ? SlShape<3>::TILE : SlShape<6>::TILE) + 8;
    cudaError_t er = cudaMalloc(&e->probe_data, ((size_t)probe_slots + kSlSpill) * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->ans, (size_t)probe_slots + kSlSpill);
    if (er == cudaSuccess) er = cudaMalloc(&e->tile_meta, (size_t)n_tiles * (e->probe_B + 1) * 8);
    if (er == cudaSuccess) er = cudaMalloc(&e->probe_cursor, (size_t)e->probe_B * kSlPad * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->pos, ((size_t)n_max + 8) * nj * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->key_data, ((size_t)key_slots + kSlSpill) * 8);
    if (er == cudaSuccess) er = cudaMalloc(&e->key_cursor, (size_t)e->key_B * kSlPad * 4);
    const int64_t n_sub_regions = (int64_t)e->key_B << e->sub_bits;
    if (n_sub_regions * e->sub_cap >= (1LL << 32) - (1LL << 20)) { sliced_engine_free(g); return fail(ctx, RB_EINVAL, "sliced engine: round too large for 32-bit record positions"); }
    if (er == cudaSuccess) er = cudaMalloc(&e->sub_data, ((size_t)n_sub_regions * e->sub_cap + kSlSpill) * 8);
    if (er == cudaSuccess) er = cudaMalloc(&e->sub_cursor, (size_t)n_sub_regions * 4 + 64);
    if (er == cudaSuccess && e->spill_on) {
        er = cudaMalloc(&e->spill_keys, ((size_t)e->spill_cap + kSlSpill) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&e->spill_cursor, 64);
        if (er == cudaSuccess) er = cudaMalloc(&e->htab_keys, (size_t)(e->htab_slots + 1) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&e->htab_counts, (size_t)(e->htab_slots + 1) * 4);
    }
    if (er == cudaSuccess) er = cudaMalloc(&e->dkey, ((size_t)n_max + 8) * 8);
    if (er == cudaSuccess) er = cudaMalloc(&e->dmult, ((size_t)n_max + 8) * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->n_distinct, 64);
    if (er == cudaSuccess) er = cudaMalloc(&e->chunk_prefix, (size_t)(maxB + 2) * 4 + 64);
    if (er == cudaSuccess) er = cudaMalloc(&e->overflow, 64);
    if (er == cudaSuccess) er = cudaMemsetAsync(e->overflow, 0, 4, ctx->stream);
    if (er != cudaSuccess) {
        // not enough device memory for the work buffers of a round this large: the direct engine needs none and serves the graph
        cudaGetLastError();
        sliced_engine_free(g);
        SlicedEngine* stub = new SlicedEngine();
        memset(stub, 0, sizeof *stub);
        stub->unsupported = true;
        g->se = stub;
        *out = stub;
        return RB_OK;
    }
    return RB_OK;
}

static int32_t sl_read_flag(rb_ctx* ctx, int* dev_flag, int* host) {
    CK(cudaMemcpyAsync(host, dev_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (*host) CK(cudaMemsetAsync(dev_flag, 0, 4, ctx->stream));
    return RB_OK;
}
template <typename K>
static int32_t sl_persistent_grid(rb_ctx* ctx, K kernel, size_t smem, int* grid) {
    if (smem > 32 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kSlThreads, smem));
    if (occ < 1) return fail(ctx, RB_ECUDA, "sliced engine: kernel does not fit on an SM");
    *grid = ctx->sm_count * std::min(occ, sl_consumer_occ());
    return RB_OK;
}
template <typename K>
static int32_t sl_allow_smem(rb_ctx* ctx, K kernel, size_t smem) {
    if (smem > 32 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return RB_OK;
}
// work list of an arena (one small CTA)
static int32_t sl_chunk_prefix(rb_ctx* ctx, SlicedEngine* e, const SlArena& a) {
    const size_t sm = ((size_t)((a.B + 3) & ~3) + 296) * 4;
    int32_t rc = sl_allow_smem(ctx, ks_chunk_prefix, sm);
    if (rc) return rc;
    PROF("ks_chunk_prefix");
    RB_LAUNCH(1, kSlThreads, sm, ctx->stream, ks_chunk_prefix)(a, e->chunk_prefix);
    LAUNCH_CHECK();
    return RB_OK;
}
static SlArena sl_arena(void* data, unsigned int* cursor, const uint32_t* roff, int B, int chunk) {
    SlArena a;
    a.data = data; a.cursor = cursor; a.roff = roff; a.B = B; a.chunk = chunk; a.cap = 0; a.cursor_stride = kSlPad; a.rlo = nullptr;
    a.spill_data = nullptr; a.spill_cursor = nullptr; a.spill_cap = 0;
    a.peer_data = nullptr; a.peer_ans = nullptr; a.n_peers = 1; a.passes = 1; a.stage = 0;
    return a;
}
static void sl_apply_stage(SlArena* a);
static SlArena sl_probe_arena(SlicedEngine* e) {
    SlArena a = sl_arena(e->probe_data, e->probe_cursor, e->probe_roff, e->probe_B, sl_chunk());
    if (e->paired) a.passes = 1 << e->sg.pair_sub_log2;
    sl_apply_stage(&a);
    return a;
}

// the prefix k-merizer covers a CTA's 1024 positions with one span of at most kPfxSpan bases of the packed stream (uniform layout only)
static bool sl_uniform_fast(const Ingest& ing, int k, int tile, int span_cap) {
    const char* v = getenv("RB_SLICED_KMERIZER");
    if (v && !strcmp(v, "walker")) return false;
    if (ing.pos_off || ing.uniform_npos <= 0) return false;
    const int64_t span = ((int64_t)(tile - 1) / ing.uniform_npos + 1) * ing.uniform_stride + ing.uniform_npos - 1 + k;
    return span <= span_cap;
}
static bool sl_uniform_fast_keys(const Ingest& ing, int k) { return sl_uniform_fast(ing, k, kSlTile, KeyKmerizer::kSpan); }
template <int NJ>
static bool sl_uniform_fast_probes(const Ingest& ing, int k) {
    return sl_uniform_fast(ing, k, SlShape<NJ>::TILE, PrefixKmerizer<SlShape<NJ>::PFX_PER, SlShape<NJ>::TILE>::kSpan);
}
// grid of a kernel that streams over an arena with no residency window to respect
template <typename K>
static int32_t sl_stream_grid(rb_ctx* ctx, K kernel, size_t smem, int* grid) {
    if (smem > 32 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kSlThreads, smem));
    if (occ < 1) return fail(ctx, RB_ECUDA, "sliced engine: kernel does not fit on an SM");
    *grid = ctx->sm_count * std::min(occ, 8);
    return RB_OK;
}
#define SL_LAUNCH(name, kern, grid, smem, ...)                                   \
    do {                                                                         \
        rc = sl_allow_smem(ctx, kern, smem);                                     \
        if (rc) return rc;                                                       \
        PROF(name);                                                              \
        RB_LAUNCH(grid, kSlThreads, smem, ctx->stream, kern)(__VA_ARGS__);       \
        LAUNCH_CHECK();                                                          \
    } while (0)

// Which representation of dbgbf + cbf this round works on.  The cell array (one access per paired record) when the graph has one and is
// already in it, or the round is large enough to pay for the conversion (one streaming pass over both filters), or RB_SLICED_CELLS=1.
struct SlLayout { SlGeom sg; uint32_t *dbg, *cbf; };
static int32_t sl_round_layout(rb_graph* g, SlicedEngine* e, int64_t n_pos, SlLayout* out) {
    out->sg = e->sg; out->sg.cells = 0; out->dbg = g->dbg->dev; out->cbf = g->cbf->dev;
    CellStore* cs = e->paired ? g->dbg->cs : nullptr;
    if (!cs) return RB_OK;
    const char* v = getenv("RB_SLICED_CELLS");
    const bool forced = v && !strcmp(v, "1");
    if (!(cs->in_cells || forced || n_pos >= (1LL << 24))) return RB_OK;
    const int32_t rc = cells_ensure_cells(g->ctx, cs);
    if (rc) return rc;
    out->sg.cells = 1; out->dbg = cs->cells; out->cbf = cs->cells;
    return RB_OK;
}

// dynamic shared memory of the apply kernels: the work-list prefix table, then (when the work items fit) two staging buffers of records
// and, for the raise sweep, two of raise bytes
static size_t sl_apply_smem(const SlArena& a, bool with_raise_bytes) {
    size_t sm = ((size_t)(a.B + 1) * 4 + 15) & ~(size_t)15;
    if (a.stage) sm += (size_t)2 * a.chunk * (with_raise_bytes ? 5 : 4);
    return sm;
}
// cp.async staging of the work items: when the records are read out of a peer's arena (NVLink latency), or when asked for (tests)
static void sl_apply_stage(SlArena* a) {
    a->stage = (a->peer_data != nullptr || env_int("RB_SLICED_STAGE", 0, 0, 1)) && a->chunk <= kSlStageRecords ? 1 : 0;
}

// the apply kernels, with or without cp.async staging of the work items (SlArena::stage)
template <int SET>
static int32_t sl_launch_apply(rb_ctx* ctx, const SlArena& a, int* chunk_prefix, const SlGeom& sg, uint32_t* dbg, const uint32_t* cbf, uint8_t* ans, const int* abort) {
    const size_t sm = sl_apply_smem(a, false);
    const char* name = SET ? "ks_apply_probes<1>" : "ks_apply_probes<0>";
    int grid = 0;
    int32_t rc;
    auto ks = ks_apply_probes<SET, true>;
    auto kd = ks_apply_probes<SET, false>;
    if (a.stage) { rc = sl_persistent_grid(ctx, ks, sm, &grid); if (rc) return rc; SL_LAUNCH(name, ks, grid, sm, a, chunk_prefix, sg, dbg, cbf, ans, abort); }
    else { rc = sl_persistent_grid(ctx, kd, sm, &grid); if (rc) return rc; SL_LAUNCH(name, kd, grid, sm, a, chunk_prefix, sg, dbg, cbf, ans, abort); }
    return RB_OK;
}
static int32_t sl_launch_raises(rb_ctx* ctx, const SlArena& a, int* chunk_prefix, const SlGeom& sg, uint32_t* cbf, const uint8_t* raise, const int* abort) {
    const size_t sm = sl_apply_smem(a, true);
    int grid = 0;
    int32_t rc;
    auto ks = ks_apply_raises<true>;
    auto kd = ks_apply_raises<false>;
    if (a.stage) { rc = sl_persistent_grid(ctx, ks, sm, &grid); if (rc) return rc; SL_LAUNCH("ks_apply_raises", ks, grid, sm, a, chunk_prefix, sg, cbf, raise, abort); }
    else { rc = sl_persistent_grid(ctx, kd, sm, &grid); if (rc) return rc; SL_LAUNCH("ks_apply_raises", kd, grid, sm, a, chunk_prefix, sg, cbf, raise, abort); }
    return RB_OK;
}

// S1..S3
template <int NJ>
static int32_t sliced_count_round_t(rb_graph* g, SlicedEngine* e, const Ingest& ing, int mode, float* counts, int64_t* fh, int64_t* rh, bool* fell_back) {
    rb_ctx* ctx = g->ctx;
    int32_t rc;
    constexpr int TILE = SlShape<NJ>::TILE, KPT = SlShape<NJ>::KPT;
    const HashMults hm = make_hm(g->k);
    const SlArena probes = sl_probe_arena(e);
    CK(cudaMemsetAsync(probes.cursor, 0, (size_t)probes.B * kSlPad * 4, ctx->stream));
    const size_t sm_sort = TileSort<uint32_t, KPT * NJ>::smem_bytes(probes.B);
    const bool fast = sl_uniform_fast_probes<NJ>(ing, g->k);
    int grid_pos;
    if (fast) {
        grid_pos = (int)div_up(ing.n_pos, (int64_t)TILE);
        const size_t sm = std::max(sm_sort, PrefixKmerizer<SlShape<NJ>::PFX_PER, TILE>::smem_bytes());
        auto kf = ks_route_lookup_u<0, NJ>;
        auto kc = ks_route_lookup_u<2, NJ>;
        if (mode == RB_MODE_FWD) SL_LAUNCH("ks_route_lookup_u<0>", kf, grid_pos, sm, ing, g->k, hm, e->sg, probes, e->pos, e->tile_meta, fh, rh, e->overflow);
        else SL_LAUNCH("ks_route_lookup_u<2>", kc, grid_pos, sm, ing, g->k, hm, e->sg, probes, e->pos, e->tile_meta, fh, rh, e->overflow);
    } else {
        grid_pos = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kChunk);
        auto kf = ks_route_lookup<0, NJ>;
        auto kc = ks_route_lookup<2, NJ>;
        if (mode == RB_MODE_FWD) SL_LAUNCH("ks_route_lookup<0>", kf, grid_pos, sm_sort, ing, g->k, hm, e->sg, probes, e->pos, e->tile_meta, fh, rh, e->overflow);
        else SL_LAUNCH("ks_route_lookup<2>", kc, grid_pos, sm_sort, ing, g->k, hm, e->sg, probes, e->pos, e->tile_meta, fh, rh, e->overflow);
    }
    int flag = 0;
    rc = sl_read_flag(ctx, e->overflow, &flag);
    if (rc) return rc;
    if (flag) { *fell_back = true; return RB_OK; }   // skewed hashes (one k-mer dominating the batch): the direct engine redoes the round
    rc = sl_chunk_prefix(ctx, e, probes);
    if (rc) return rc;
    SlLayout lay;
    rc = sl_round_layout(g, e, ing.n_pos, &lay);
    if (rc) return rc;
    rc = sl_launch_apply<0>(ctx, probes, e->chunk_prefix, lay.sg, lay.dbg, lay.cbf, e->ans, nullptr);
    if (rc) return rc;
    // same CTA -> k-mer mapping as the route kernel
    const size_t sm_ans = TileAnswers::smem_bytes(probes.B, TILE * NJ);
    auto k1 = ks_combine_lookup<1, NJ>;
    auto k0 = ks_combine_lookup<0, NJ>;
    if (fast) SL_LAUNCH("ks_combine_lookup<1>", k1, grid_pos, sm_ans, e->pos, e->tile_meta, probes.B, e->ans, ing.n_pos, g->hd, g->hc, counts, ing.out_base);
    else SL_LAUNCH("ks_combine_lookup<0>", k0, grid_pos, sm_ans, e->pos, e->tile_meta, probes.B, e->ans, ing.n_pos, g->hd, g->hc, counts, ing.out_base);
    return RB_OK;
}
static int32_t sliced_count_round(rb_graph* g, const Ingest& ing, int mode, float* counts, int64_t* fh, int64_t* rh, bool* fell_back) {
    SlicedEngine* e = nullptr;
    int32_t rc = sliced_engine_get(g, ing.n_pos, &e);
    if (rc) return rc;
    if (e->unsupported) { *fell_back = true; return RB_OK; }
    return e->paired ? sliced_count_round_t<3>(g, e, ing, mode, counts, fh, rh, fell_back) : sliced_count_round_t<6>(g, e, ing, mode, counts, fh, rh, fell_back);
}

// I1..I7
static int32_t sliced_insert_round(rb_graph* g, const Ingest& ing, int mode, int policy, bool* fell_back) {
    rb_ctx* ctx = g->ctx;
    SlicedEngine* e = nullptr;
    int32_t rc = sliced_engine_get(g, ing.n_pos, &e);
    if (rc) return rc;
    if (e->unsupported) { *fell_back = true; return RB_OK; }
    const HashMults hm = make_hm(g->k);
    // I1 keys by range
    SlArena keys = sl_arena(e->key_data, e->key_cursor, e->key_roff, e->key_B, kSlThreads * kKeyE);
    CK(cudaMemsetAsync(keys.cursor, 0, (size_t)keys.B * kSlPad * 4, ctx->stream));
    if (e->spill_on) {
        keys.spill_data = e->spill_keys; keys.spill_cursor = e->spill_cursor; keys.spill_cap = e->spill_cap;
        CK(cudaMemsetAsync(e->spill_cursor, 0, 4, ctx->stream));
    }
    if (sl_uniform_fast_keys(ing, g->k)) {
        const int grid_pos = (int)div_up(ing.n_pos, (int64_t)kKeyTile);
        const size_t sm = std::max(TileSort<unsigned long long, kKeyE, true>::smem_bytes(keys.B), KeyKmerizer::smem_bytes());
        if (mode == RB_MODE_FWD) SL_LAUNCH("ks_route_keys_u<0>", ks_route_keys_u<0>, grid_pos, sm, ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
        else if (mode == RB_MODE_RC) SL_LAUNCH("ks_route_keys_u<1>", ks_route_keys_u<1>, grid_pos, sm, ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
        else SL_LAUNCH("ks_route_keys_u<2>", ks_route_keys_u<2>, grid_pos, sm, ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
    } else {
        const int grid_pos = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kChunk);
        const size_t sm_keys = TileSort<unsigned long long, kChunk, true>::smem_bytes(keys.B);
        if (mode == RB_MODE_FWD) SL_LAUNCH("ks_route_keys<0>", ks_route_keys<0>, grid_pos, sm_keys, ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
        else if (mode == RB_MODE_RC) SL_LAUNCH("ks_route_keys<1>", ks_route_keys<1>, grid_pos, sm_keys, ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
        else SL_LAUNCH("ks_route_keys<2>", ks_route_keys<2>, grid_pos, sm_keys, ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
    }
    // I2 second-level split (no flag read in between: an overflow of level 1 only drops keys, level 2 then sees fewer)
    const int n_sub = 1 << e->sub_bits;
    const int n_sub_regions = e->key_B << e->sub_bits;
    SlArena subs = sl_arena(e->sub_data, e->sub_cursor, nullptr, n_sub_regions, 0);
    subs.cap = e->sub_cap; subs.cursor_stride = 1;
    subs.spill_data = keys.spill_data; subs.spill_cursor = keys.spill_cursor; subs.spill_cap = keys.spill_cap;
    CK(cudaMemsetAsync(subs.cursor, 0, (size_t)n_sub_regions * 4, ctx->stream));
    rc = sl_chunk_prefix(ctx, e, keys);
    if (rc) return rc;
    int grid = 0;
    const size_t sm_split = TileSort<unsigned long long, kKeyE, true>::smem_bytes(n_sub) + (size_t)(keys.B + 1) * 4;
    rc = sl_stream_grid(ctx, ks_split_keys, sm_split, &grid);
    if (rc) return rc;
    SL_LAUNCH("ks_split_keys", ks_split_keys, grid, sm_split, keys, e->chunk_prefix, e->sub_bits, 64 - (64 - e->key_shift) - e->sub_bits, 1, subs, e->overflow);
    int flag = 0;
    rc = sl_read_flag(ctx, e->overflow, &flag);
    if (rc) return rc;
    if (flag) { *fell_back = true; return RB_OK; }   // key skew (one k-mer dominating the batch): nothing modified yet
    // heavy hitters: what did not fit its range / sub-range is aggregated in a global table that ks_dedup merges from
    SpillTable spill;
    memset(&spill, 0, sizeof spill);
    if (e->spill_on) {
        unsigned int n_spill = 0;
        CK(cudaMemcpyAsync(&n_spill, e->spill_cursor, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if ((int64_t)n_spill > e->htab_slots / 2) { *fell_back = true; return RB_OK; }   // more spilled copies than the table is sure to hold
        if (n_spill) {
            spill.keys = e->htab_keys; spill.counts = e->htab_counts; spill.n_slots = (uint64_t)e->htab_slots; spill.shift = e->htab_shift;
            CK(cudaMemsetAsync(spill.keys, 0, (size_t)(e->htab_slots + 1) * 8, ctx->stream));
            CK(cudaMemsetAsync(spill.counts, 0, (size_t)(e->htab_slots + 1) * 4, ctx->stream));
            const int grid_s = (int)std::min<int64_t>(div_up((int64_t)n_spill, kSlThreads), (int64_t)ctx->sm_count * 8);
            SL_LAUNCH("ks_spill_aggregate", ks_spill_aggregate, grid_s, 0, e->spill_keys, e->spill_cursor, e->spill_cap, spill);
        }
    }
    // I3 distinct keys and their multiplicities
    CK(cudaMemsetAsync(e->n_distinct, 0, 4, ctx->stream));
    const size_t sm_dedup = (size_t)kSlDedupSlots * 12;
    rc = sl_stream_grid(ctx, ks_dedup, sm_dedup, &grid);
    if (rc) return rc;
    SL_LAUNCH("ks_dedup", ks_dedup, std::min(grid, n_sub_regions), sm_dedup, subs, n_sub_regions, (64 - e->key_shift) + e->sub_bits, e->dkey, e->dmult, e->n_distinct,
              (unsigned int)std::min<int64_t>(e->n_max + 8, 0xFFFFFFFFLL), e->overflow, spill);
    if (spill.keys) {
        const int grid_a = (int)std::min<int64_t>(div_up(e->htab_slots + 1, kSlThreads), (int64_t)ctx->sm_count * 8);
        SL_LAUNCH("ks_spill_append", ks_spill_append, grid_a, 0, spill, e->dkey, e->dmult, e->n_distinct, (unsigned int)std::min<int64_t>(e->n_max + 8, 0xFFFFFFFFLL), e->overflow);
    }
    // I4 probes by filter slice
    const SlArena probes = sl_probe_arena(e);
    CK(cudaMemsetAsync(probes.cursor, 0, (size_t)probes.B * kSlPad * 4, ctx->stream));
    const int with_cbf = policy != POLICY_DBG_ONLY;
    const size_t sm_sort = TileSort<uint32_t, kSlTileRecords>::smem_bytes(probes.B);
    const int tile_d = e->paired ? SlShape<3>::TILE : SlShape<6>::TILE;
    const int grid_d = (int)div_up(ing.n_pos, (int64_t)tile_d);   // distinct keys <= instances
    if (e->paired) SL_LAUNCH("ks_emit_probes", ks_emit_probes<3>, grid_d, sm_sort, e->dkey, e->n_distinct, hm, e->sg, with_cbf, probes, e->pos, e->tile_meta, e->overflow);
    else SL_LAUNCH("ks_emit_probes", ks_emit_probes<6>, grid_d, sm_sort, e->dkey, e->n_distinct, hm, e->sg, with_cbf, probes, e->pos, e->tile_meta, e->overflow);
    rc = sl_read_flag(ctx, e->overflow, &flag);
    if (rc) return rc;
    if (flag) { *fell_back = true; return RB_OK; }   // still nothing modified
    // I5 apply
    rc = sl_chunk_prefix(ctx, e, probes);
    if (rc) return rc;
    SlLayout lay;
    rc = sl_round_layout(g, e, ing.n_pos, &lay);
    if (rc) return rc;
    rc = policy != POLICY_COUNT_IF_PRESENT ? sl_launch_apply<1>(ctx, probes, e->chunk_prefix, lay.sg, lay.dbg, lay.cbf, e->ans, nullptr)
                                            : sl_launch_apply<0>(ctx, probes, e->chunk_prefix, lay.sg, lay.dbg, lay.cbf, e->ans, nullptr);
    if (rc) return rc;
    if (with_cbf) {
        // I6: the new counter values go back over the answer bytes of the probe records; I7: a second sweep over the same regions applies
        // them.  Nothing is sorted and nothing can overflow here: once the probes are routed the round always completes.
        const uint64_t seed = ctx->rng_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(ctx->launches + 1);
        const size_t sm_r = TileAnswers::smem_bytes(probes.B, kSlThreads * kSlTileRecords);
        if (e->paired) SL_LAUNCH("ks_combine_insert", ks_combine_insert<3>, grid_d, sm_r, e->dkey, e->dmult, e->n_distinct, e->pos, e->tile_meta, probes.B, e->ans, e->sg,
                                 policy, seed, (const int*)nullptr);
        else SL_LAUNCH("ks_combine_insert", ks_combine_insert<6>, grid_d, sm_r, e->dkey, e->dmult, e->n_distinct, e->pos, e->tile_meta, probes.B, e->ans, e->sg,
                       policy, seed, (const int*)nullptr);
        rc = sl_chunk_prefix(ctx, e, probes);   // the same work list again (the consumers' counter starts from 0)
        if (rc) return rc;
        rc = sl_launch_raises(ctx, probes, e->chunk_prefix, lay.sg, lay.cbf, (const uint8_t*)e->ans, nullptr);
        if (rc) return rc;
    }
    claim_invalidate(ctx);   // bits were set without going through the claim table
    return RB_OK;
}
