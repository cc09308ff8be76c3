// rb_sliced_host.inl -- buffers, geometry and round drivers of the sliced engine (kernels: rb_sliced.cuh).  Included by rnabloom_gpu.cu.

struct SlicedEngine {
    int64_t n_max;                // k-mer instances per round the buffers hold
    SlGeom sg;
    bool unsupported;             // this graph cannot use the engine (hash counts / filter sizes): the direct engine serves it
    // probes: 4-byte slice-local indices, one answer byte per probe, 6 remembered positions per k-mer (or distinct key)
    uint32_t* probe_data; uint8_t* ans; unsigned int* probe_cursor; uint32_t* probe_roff; int probe_B;
    uint32_t* pos;
    // insert: keys by range, hash table, dense distinct keys, raises
    unsigned long long* key_data; unsigned int* key_cursor; uint32_t* key_roff; int key_B; int key_shift;
    unsigned long long* tab_keys; unsigned int* tab_counts; int64_t T; int tab_shift;
    unsigned long long* dkey; unsigned int* dmult; unsigned int* n_distinct;
    uint32_t* raise_data; unsigned int* raise_cursor; uint32_t* raise_roff;
    int* chunk_prefix;
    int* overflow;
};
static void sliced_engine_free(rb_graph* g) {
    SlicedEngine* e = g->se;
    if (!e) return;
    cudaStreamSynchronize(g->ctx->stream);
    cudaFree(e->probe_data); cudaFree(e->ans); cudaFree(e->probe_cursor); cudaFree(e->probe_roff); cudaFree(e->pos);
    cudaFree(e->key_data); cudaFree(e->key_cursor); cudaFree(e->key_roff);
    cudaFree(e->tab_keys); cudaFree(e->tab_counts); cudaFree(e->dkey); cudaFree(e->dmult); cudaFree(e->n_distinct);
    cudaFree(e->raise_data); cudaFree(e->raise_cursor); cudaFree(e->raise_roff);
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
static int sl_chunk() { return env_int("RB_SLICED_CHUNK", 2048, 256, 1 << 16); }
static int sl_consumer_occ() { return env_int("RB_SLICED_CONSUMER_OCC", 2, 1, 8); }
static int64_t sl_pow2_at_least(int64_t v) { int64_t p = 1024; while (p < v) p <<= 1; return p; }
// capacity of a region that expects `expected` records from uniform hashes: 4 % + 8 sigma + a constant
static int64_t sl_capacity(double expected) { return (int64_t)(expected * 1.04 + 8.0 * std::sqrt(expected + 1.0)) + 2048; }
static int64_t sliced_round_kmers(const rb_ctx* ctx) {
    if (ctx->subbatch_user_set) return std::min<int64_t>(ctx->subbatch_kmers, 1LL << 29);
    return 1LL << env_int("RB_SLICED_ROUND_LOG2", 28, 10, 29);
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
    sg.raise_log2 = std::min(sg.cbf_log2, env_int("RB_SLICE_RAISE_LOG2", 25, 2, 25));
    const int64_t n_raise = div_up(g->cbf->size, 1LL << sg.raise_log2);
    sg.n_raise = (int)std::min<int64_t>(n_raise, 1 << 20);
    if (g->hd > kSlMaxH || g->hc > kSlMaxH || sg.n_dbg + sg.n_cbf > kSlMaxRegions || n_raise > kSlMaxRegions) { e->unsupported = true; return RB_OK; }
    const int64_t n_max = sl_pow2_at_least(n_round);
    e->n_max = n_max;
    e->T = sl_pow2_at_least(2 * n_max);
    int lgT = 0; while ((1LL << lgT) < e->T) ++lgT;
    e->tab_shift = 64 - lgT;
    const int lgRange = env_int("RB_SLICE_TABLE_LOG2", 21, 4, 30);   // table slots per key range (12 B each)
    const int lgR = std::max(0, std::min(lgT - lgRange, 11));
    e->key_B = 1 << lgR;
    e->key_shift = 64 - lgR;
    e->probe_B = sg.n_dbg + sg.n_cbf;
    // ---- capacities ----
    const double dbg_slices = std::max(1.0, (double)g->dbg->size / (double)(1LL << sg.dbg_log2));
    const double cbf_slices = std::max(1.0, (double)g->cbf->size / (double)(1LL << sg.cbf_log2));
    const double raise_slices = std::max(1.0, (double)g->cbf->size / (double)(1LL << sg.raise_log2));
    std::vector<int64_t> caps;
    for (int b = 0; b < sg.n_dbg; ++b) caps.push_back(sl_capacity((double)n_max * g->hd / dbg_slices));
    for (int b = 0; b < sg.n_cbf; ++b) caps.push_back(sl_capacity((double)n_max * g->hc / cbf_slices));
    int64_t probe_slots = 0, key_slots = 0, raise_slots = 0;
    int32_t rc = sl_make_roff(ctx, caps, &e->probe_roff, &probe_slots);
    if (rc) { sliced_engine_free(g); return rc; }
    caps.assign((size_t)e->key_B, sl_capacity((double)n_max / e->key_B));
    rc = sl_make_roff(ctx, caps, &e->key_roff, &key_slots);
    if (rc) { sliced_engine_free(g); return rc; }
    caps.assign((size_t)sg.n_raise, sl_capacity((double)n_max * g->hc / raise_slices));
    rc = sl_make_roff(ctx, caps, &e->raise_roff, &raise_slots);
    if (rc) { sliced_engine_free(g); return rc; }
    const int maxB = std::max(std::max(e->probe_B, e->key_B), sg.n_raise);
    cudaError_t er = cudaMalloc(&e->probe_data, (size_t)probe_slots * 4 + 64);
    if (er == cudaSuccess) er = cudaMalloc(&e->ans, (size_t)probe_slots + 64);
    if (er == cudaSuccess) er = cudaMalloc(&e->probe_cursor, (size_t)e->probe_B * kSlPad * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->pos, ((size_t)n_max + 8) * kSlNJ * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->key_data, (size_t)key_slots * 8 + 64);
    if (er == cudaSuccess) er = cudaMalloc(&e->key_cursor, (size_t)e->key_B * kSlPad * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->tab_keys, (size_t)(e->T + 1) * 8);
    if (er == cudaSuccess) er = cudaMalloc(&e->tab_counts, (size_t)(e->T + 1) * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->dkey, ((size_t)n_max + 8) * 8);
    if (er == cudaSuccess) er = cudaMalloc(&e->dmult, ((size_t)n_max + 8) * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->n_distinct, 64);
    if (er == cudaSuccess) er = cudaMalloc(&e->raise_data, (size_t)raise_slots * 4 + 64);
    if (er == cudaSuccess) er = cudaMalloc(&e->raise_cursor, (size_t)sg.n_raise * kSlPad * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->chunk_prefix, (size_t)(maxB + 1) * 4);
    if (er == cudaSuccess) er = cudaMalloc(&e->overflow, 64);
    if (er == cudaSuccess) er = cudaMemsetAsync(e->overflow, 0, 4, ctx->stream);
    if (er != cudaSuccess) { sliced_engine_free(g); return fail(ctx, RB_ENOMEM, std::string("sliced engine buffers: ") + cudaGetErrorString(er)); }
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
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kSlThreads, smem));
    if (occ < 1) return fail(ctx, RB_ECUDA, "sliced engine: kernel does not fit on an SM");
    *grid = ctx->sm_count * std::min(occ, sl_consumer_occ());
    return RB_OK;
}
template <typename K>
static int32_t sl_allow_smem(rb_ctx* ctx, K kernel, size_t smem) {
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
static SlArena sl_probe_arena(SlicedEngine* e) { SlArena a; a.data = e->probe_data; a.cursor = e->probe_cursor; a.roff = e->probe_roff; a.B = e->probe_B; a.chunk = sl_chunk(); return a; }

// S1..S3
static int32_t sliced_count_round(rb_graph* g, const Ingest& ing, int mode, float* counts, int64_t* fh, int64_t* rh, bool* fell_back) {
    rb_ctx* ctx = g->ctx;
    SlicedEngine* e = nullptr;
    int32_t rc = sliced_engine_get(g, ing.n_pos, &e);
    if (rc) return rc;
    if (e->unsupported) { *fell_back = true; return RB_OK; }
    const HashMults hm = make_hm(g->k);
    const SlArena probes = sl_probe_arena(e);
    CK(cudaMemsetAsync(probes.cursor, 0, (size_t)probes.B * kSlPad * 4, ctx->stream));
    const int grid_pos = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kChunk);
    const size_t sm_sort = TileSort<uint32_t, kSlRoundKmers * kSlNJ>::smem_bytes(probes.B);
    if (mode == RB_MODE_FWD) {
        rc = sl_allow_smem(ctx, ks_route_lookup<0>, sm_sort); if (rc) return rc;
        PROF("ks_route_lookup<0>");
        RB_LAUNCH(grid_pos, kSlThreads, sm_sort, ctx->stream, ks_route_lookup<0>)(ing, g->k, hm, e->sg, probes, e->pos, fh, rh, e->overflow);
    } else {
        rc = sl_allow_smem(ctx, ks_route_lookup<2>, sm_sort); if (rc) return rc;
        PROF("ks_route_lookup<2>");
        RB_LAUNCH(grid_pos, kSlThreads, sm_sort, ctx->stream, ks_route_lookup<2>)(ing, g->k, hm, e->sg, probes, e->pos, fh, rh, e->overflow);
    }
    LAUNCH_CHECK();
    int flag = 0;
    rc = sl_read_flag(ctx, e->overflow, &flag);
    if (rc) return rc;
    if (flag) { *fell_back = true; return RB_OK; }   // skewed hashes (one k-mer dominating the batch): the direct engine redoes the round
    rc = sl_chunk_prefix(ctx, e, probes);
    if (rc) return rc;
    const size_t sm_pre = (size_t)(probes.B + 1) * 4;
    int grid = 0;
    rc = sl_persistent_grid(ctx, ks_apply_probes<0>, sm_pre, &grid);
    if (rc) return rc;
    PROF("ks_apply_probes<0>");
    RB_LAUNCH(grid, kSlThreads, sm_pre, ctx->stream, ks_apply_probes<0>)(probes, e->chunk_prefix, e->sg, g->dbg->dev, g->cbf->dev, e->ans);
    LAUNCH_CHECK();
    const int grid_c = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kSlRoundKmers);
    PROF("ks_combine_lookup");
    RB_LAUNCH(grid_c, kSlThreads, 0, ctx->stream, ks_combine_lookup)(e->pos, e->ans, ing.n_pos, g->hd, g->hc, counts, ing.out_base);
    LAUNCH_CHECK();
    return RB_OK;
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
    SlArena keys; keys.data = e->key_data; keys.cursor = e->key_cursor; keys.roff = e->key_roff; keys.B = e->key_B; keys.chunk = sl_chunk();
    CK(cudaMemsetAsync(keys.cursor, 0, (size_t)keys.B * kSlPad * 4, ctx->stream));
    const int grid_pos = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kChunk);
    const size_t sm_keys = TileSort<unsigned long long, kChunk>::smem_bytes(keys.B);
    if (mode == RB_MODE_FWD) {
        rc = sl_allow_smem(ctx, ks_route_keys<0>, sm_keys); if (rc) return rc;
        PROF("ks_route_keys<0>");
        RB_LAUNCH(grid_pos, kSlThreads, sm_keys, ctx->stream, ks_route_keys<0>)(ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
    } else if (mode == RB_MODE_RC) {
        rc = sl_allow_smem(ctx, ks_route_keys<1>, sm_keys); if (rc) return rc;
        PROF("ks_route_keys<1>");
        RB_LAUNCH(grid_pos, kSlThreads, sm_keys, ctx->stream, ks_route_keys<1>)(ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
    } else {
        rc = sl_allow_smem(ctx, ks_route_keys<2>, sm_keys); if (rc) return rc;
        PROF("ks_route_keys<2>");
        RB_LAUNCH(grid_pos, kSlThreads, sm_keys, ctx->stream, ks_route_keys<2>)(ing, g->k, e->key_B, e->key_shift, keys, e->overflow);
    }
    LAUNCH_CHECK();
    int flag = 0;
    rc = sl_read_flag(ctx, e->overflow, &flag);
    if (rc) return rc;
    if (flag) { *fell_back = true; return RB_OK; }   // extreme key skew: nothing modified yet
    // I2 aggregate
    SlTable t; t.keys = e->tab_keys; t.counts = e->tab_counts; t.n_slots = (uint64_t)e->T; t.shift = e->tab_shift;
    PROF("memset(table)");
    CK(cudaMemsetAsync(t.keys, 0, (size_t)(e->T + 1) * 8, ctx->stream));
    CK(cudaMemsetAsync(t.counts, 0, (size_t)(e->T + 1) * 4, ctx->stream));
    if (ctx->prof_pending) prof_end(ctx);
    rc = sl_chunk_prefix(ctx, e, keys);
    if (rc) return rc;
    int grid = 0;
    rc = sl_persistent_grid(ctx, ks_aggregate, (size_t)(keys.B + 1) * 4, &grid);
    if (rc) return rc;
    PROF("ks_aggregate");
    RB_LAUNCH(grid, kSlThreads, (size_t)(keys.B + 1) * 4, ctx->stream, ks_aggregate)(keys, e->chunk_prefix, t);
    LAUNCH_CHECK();
    // I3 dense distinct keys
    CK(cudaMemsetAsync(e->n_distinct, 0, 4, ctx->stream));
    const int grid_t = (int)div_up(e->T + 1, (int64_t)kSlThreads * kSlCompactPer);
    PROF("ks_compact_table");
    RB_LAUNCH(grid_t, kSlThreads, 0, ctx->stream, ks_compact_table)(t, e->dkey, e->dmult, e->n_distinct);
    LAUNCH_CHECK();
    // I4 probes by filter slice
    const SlArena probes = sl_probe_arena(e);
    CK(cudaMemsetAsync(probes.cursor, 0, (size_t)probes.B * kSlPad * 4, ctx->stream));
    const int with_cbf = policy != POLICY_DBG_ONLY;
    const size_t sm_sort = TileSort<uint32_t, kSlRoundKmers * kSlNJ>::smem_bytes(probes.B);
    rc = sl_allow_smem(ctx, ks_emit_probes, sm_sort);
    if (rc) return rc;
    const int grid_d = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kSlRoundKmers);   // distinct keys <= instances
    PROF("ks_emit_probes");
    RB_LAUNCH(grid_d, kSlThreads, sm_sort, ctx->stream, ks_emit_probes)(e->dkey, e->n_distinct, hm, e->sg, with_cbf, probes, e->pos, e->overflow);
    LAUNCH_CHECK();
    rc = sl_read_flag(ctx, e->overflow, &flag);
    if (rc) return rc;
    if (flag) { *fell_back = true; return RB_OK; }   // still nothing modified
    // I5 apply
    rc = sl_chunk_prefix(ctx, e, probes);
    if (rc) return rc;
    const size_t sm_pre = (size_t)(probes.B + 1) * 4;
    if (policy != POLICY_COUNT_IF_PRESENT) {
        rc = sl_persistent_grid(ctx, ks_apply_probes<1>, sm_pre, &grid); if (rc) return rc;
        PROF("ks_apply_probes<1>");
        RB_LAUNCH(grid, kSlThreads, sm_pre, ctx->stream, ks_apply_probes<1>)(probes, e->chunk_prefix, e->sg, g->dbg->dev, g->cbf->dev, e->ans);
    } else {
        rc = sl_persistent_grid(ctx, ks_apply_probes<0>, sm_pre, &grid); if (rc) return rc;
        PROF("ks_apply_probes<0>");
        RB_LAUNCH(grid, kSlThreads, sm_pre, ctx->stream, ks_apply_probes<0>)(probes, e->chunk_prefix, e->sg, g->dbg->dev, g->cbf->dev, e->ans);
    }
    LAUNCH_CHECK();
    if (with_cbf) {
        // I6 + I7
        SlArena raises; raises.data = e->raise_data; raises.cursor = e->raise_cursor; raises.roff = e->raise_roff; raises.B = e->sg.n_raise; raises.chunk = sl_chunk();
        CK(cudaMemsetAsync(raises.cursor, 0, (size_t)raises.B * kSlPad * 4, ctx->stream));
        const uint64_t seed = ctx->rng_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(ctx->launches + 1);
        const size_t sm_r = TileSort<uint32_t, kSlRoundKmers * kSlMaxH>::smem_bytes(raises.B);
        rc = sl_allow_smem(ctx, ks_combine_insert, sm_r);
        if (rc) return rc;
        PROF("ks_combine_insert");
        RB_LAUNCH(grid_d, kSlThreads, sm_r, ctx->stream, ks_combine_insert)(e->dkey, e->dmult, e->n_distinct, e->pos, e->ans, hm, e->sg, policy, seed, raises,
                                                                         e->overflow);
        LAUNCH_CHECK();
        rc = sl_chunk_prefix(ctx, e, raises);
        if (rc) return rc;
        const size_t sm_rp = (size_t)(raises.B + 1) * 4;
        rc = sl_persistent_grid(ctx, ks_apply_raises, sm_rp, &grid);
        if (rc) return rc;
        PROF("ks_apply_raises");
        RB_LAUNCH(grid, kSlThreads, sm_rp, ctx->stream, ks_apply_raises)(raises, e->chunk_prefix, e->sg, g->cbf->dev);
        LAUNCH_CHECK();
        rc = sl_read_flag(ctx, e->overflow, &flag);
        if (rc) return rc;
        if (flag) return fail(ctx, RB_ESTATE, "sliced engine: a raise region overflowed after filters were modified (hash skew beyond the slack)");
    }
    claim_invalidate(ctx);   // bits were set without going through the claim table
    return RB_OK;
}
