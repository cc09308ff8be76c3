// rb_sshard_host.inl -- hash-sharded graph on the sliced engine: one rank's phases between the exchanges.  Included by rnabloom_gpu.cu.
//
// The tile sort of the sliced engine *is* the routing step: a producer sorts its probes by (owner rank, filter slice of the owner),
// so the regions of one destination rank are one contiguous piece of the arena and the exchange is an all-to-all with equal split
// sizes (regions travel with their full capacity; the counts travel beside them).  The owner consumes the regions it received
// slice by slice (all sources of a slice together, so the slice stays L2-resident), writes one answer byte per probe at the probe's
// own position, and the answers go back with the mirror-image all-to-all -- they land exactly where the producer's tile metadata
// points.  Per round and rank:   lookup: probes -> | -> apply -> answers <- | -> combine
//   insert: keys -> | (home rank of the key's hash range) -> split + dedup -> probes -> | -> apply (test-and-set) -> answers <- | ->
//           combine -> raises -> | -> apply raises.          ( | = all_to_all_single over NCCL, done by the caller: rna-bloom_b200/sharded.py)
// The exchange buffers belong to the caller (torch tensors), everything else lives here.

struct rb_sshard {
    rb_ctx* ctx;
    int W, rank, hd, hc, k, stranded;
    int64_t dbg_bits, cbf_bytes;      // global sizes
    SlGeom sg_route, sg_apply;        // producer view (global regions) / consumer view (local regions, region_div = W)
    int R, SR, KR;                    // per rank: probe regions, raise regions, key ranges
    int lg1, sub_bits;
    int64_t n_max, n_dense;           // instances per rank and round; capacity of the dense distinct-key arrays
    uint32_t probe_cap, key_cap, raise_cap, sub_cap;
    rb_filter *dbg, *cbf;             // local shares
    unsigned int *probe_cursor, *key_cursor, *raise_cursor, *cons_cursor, *sub_cursor, *n_distinct;
    uint32_t *cons_rlo, *pos;
    uint2* tile_meta;
    unsigned long long *sub_data, *dkey;
    unsigned int* dmult;
    int* chunk_prefix;
    int* overflow;
    int64_t n_items;                  // instances of the last route_lookup
    bool lookup_fast;                 // which route kernel (and so which combine mapping) the last route_lookup used
};

extern "C" int32_t rb_sshard_destroy(rb_sshard* sh) {
    if (!sh) return RB_EINVAL;
    rb_ctx* ctx = sh->ctx;
    LOCK(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (sh->dbg) filter_free(sh->dbg);
    if (sh->cbf) filter_free(sh->cbf);
    cudaFree(sh->probe_cursor); cudaFree(sh->key_cursor); cudaFree(sh->raise_cursor); cudaFree(sh->cons_cursor); cudaFree(sh->sub_cursor);
    cudaFree(sh->n_distinct); cudaFree(sh->cons_rlo); cudaFree(sh->pos); cudaFree(sh->tile_meta); cudaFree(sh->sub_data); cudaFree(sh->dkey);
    cudaFree(sh->dmult); cudaFree(sh->chunk_prefix); cudaFree(sh->overflow);
    delete sh;
    return RB_OK;
}

extern "C" int32_t rb_sshard_create(rb_ctx* ctx, int32_t n_ranks, int32_t rank, int64_t dbg_bits, int64_t cbf_bytes, int32_t hd, int32_t hc, int32_t k,
                                    int32_t stranded, int64_t max_kmers, rb_sshard** out) {
    if (!ctx || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks || max_kmers < 1) return RB_EINVAL;
    if ((n_ranks & (n_ranks - 1)) != 0) return fail(ctx, RB_EINVAL, "sshard: the number of ranks must be a power of two");
    if (hd < 1 || hd > kSlMaxH || hc < 1 || hc > kSlMaxH || dbg_bits < 1 || cbf_bytes < 1) return fail(ctx, RB_EINVAL, "sshard: needs 1..3 hashes per filter");
    LOCK(ctx);
    rb_sshard* sh = new rb_sshard();
    memset(sh, 0, sizeof *sh);
    sh->ctx = ctx; sh->W = n_ranks; sh->rank = rank; sh->hd = hd; sh->hc = hc; sh->k = k; sh->stranded = stranded ? 1 : 0;
    sh->dbg_bits = dbg_bits; sh->cbf_bytes = cbf_bytes;
    const int W = n_ranks;
    SlGeom sg;
    memset(&sg, 0, sizeof sg);
    sg.dbg_fm = make_fm(dbg_bits); sg.cbf_fm = make_fm(cbf_bytes);
    sg.hd = hd; sg.hc = hc;
    sg.dbg_log2 = env_int("RB_SLICE_BITS_LOG2", 29, 5, 31);
    sg.cbf_log2 = env_int("RB_SLICE_BYTES_LOG2", 26, 2, 31);
    int64_t tot_d, tot_c;
    for (;;) {   // slices of every rank: all regions of a producer must fit the tile sort's bucket range
        tot_d = div_up(dbg_bits, 1LL << sg.dbg_log2); tot_c = div_up(cbf_bytes, 1LL << sg.cbf_log2);
        if ((div_up(tot_d, W) + div_up(tot_c, W)) * W <= kSlMaxRegions) break;
        if (tot_d >= tot_c && sg.dbg_log2 < 31) ++sg.dbg_log2; else if (sg.cbf_log2 < 31) ++sg.cbf_log2; else break;
    }
    sg.shard_d = (int)div_up(tot_d, W); sg.shard_c = (int)div_up(tot_c, W);
    sg.raise_log2 = std::min(sg.cbf_log2, env_int("RB_SLICE_RAISE_LOG2", 25, 2, 25));
    while (((int64_t)sg.shard_c << (sg.cbf_log2 - sg.raise_log2)) * W > kSlMaxRegions && sg.raise_log2 < std::min(sg.cbf_log2, 25)) ++sg.raise_log2;
    sg.shard_r = sg.shard_c << (sg.cbf_log2 - sg.raise_log2);
    sg.region_div = 1;
    sh->R = sg.shard_d + sg.shard_c; sh->SR = sg.shard_r;
    if ((int64_t)sh->R * W > kSlMaxRegions || (int64_t)sh->SR * W > kSlMaxRegions) { delete sh; return fail(ctx, RB_EINVAL, "sshard: too many filter slices for this many ranks"); }
    sg.n_dbg = sg.shard_d * W; sg.n_cbf = sg.shard_c * W; sg.n_raise = sg.shard_r * W;
    sh->sg_route = sg;
    sh->sg_apply = sg;
    sh->sg_apply.n_dbg = sg.shard_d; sh->sg_apply.n_cbf = sg.shard_c; sh->sg_apply.n_raise = sg.shard_r; sh->sg_apply.region_div = W;
    // local shares: whole slices, so that the concatenation of the ranks' shares is the global array
    const int64_t share_d = (int64_t)sg.shard_d << sg.dbg_log2, share_c = (int64_t)sg.shard_c << sg.cbf_log2;
    const int64_t local_d = std::max<int64_t>(0, std::min<int64_t>(share_d, dbg_bits - share_d * rank));
    const int64_t local_c = std::max<int64_t>(0, std::min<int64_t>(share_c, cbf_bytes - share_c * rank));
    // rounds: keys
    // capacities follow the caller's round size (no power-of-two rounding: every record of slack travels through the exchanges); the
    // distinct keys a home rank can see exceed its share of the instances only by the imbalance of the hash ranges (~1e-4 at 10^8 keys)
    const int64_t n_max = div_up(std::max<int64_t>(max_kmers, 1024), 4096) * 4096;
    sh->n_max = n_max; sh->n_dense = n_max + n_max / 32 + 4096;
    const int lgSub = env_int("RB_SLICED_SUBRANGE_LOG2", 10, 4, 11);
    int lgW = 0; while ((1 << lgW) < W) ++lgW;
    int lgS = 0; while (((n_max * W) >> lgSub) > (1LL << lgS)) ++lgS;       // sub-ranges over all ranks
    sh->lg1 = std::max(lgW, std::min((lgS + 1) / 2, 11));
    sh->sub_bits = std::max(0, std::min(lgS - sh->lg1, 11));
    sh->KR = (1 << sh->lg1) / W;
    const double keys_per_sub = (double)n_max * W / (double)(1LL << (sh->lg1 + sh->sub_bits));
    sh->sub_cap = (uint32_t)sl_capacity(keys_per_sub);
    sh->key_cap = (uint32_t)sl_capacity((double)n_max / (double)(1 << sh->lg1));
    const double slices_d = std::max(1.0, (double)dbg_bits / (double)(1LL << sg.dbg_log2)), slices_c = std::max(1.0, (double)cbf_bytes / (double)(1LL << sg.cbf_log2));
    sh->probe_cap = (uint32_t)sl_capacity(std::max((double)sh->n_dense * hd / slices_d, (double)sh->n_dense * hc / slices_c));
    const double slices_r = std::max(1.0, (double)cbf_bytes / (double)(1LL << sg.raise_log2));
    sh->raise_cap = (uint32_t)sl_capacity((double)sh->n_dense * hc / slices_r);
    const int64_t n_sub_regions = (int64_t)sh->KR << sh->sub_bits;
    if (sh->sub_cap >= (uint32_t)kSlDedupSlots || (int64_t)sh->R * W * sh->probe_cap >= (1LL << 32) - (1LL << 20) ||
        n_sub_regions * sh->sub_cap >= (1LL << 32) - (1LL << 20) || (int64_t)sh->SR * W * sh->raise_cap >= (1LL << 32) - (1LL << 20)) {
        delete sh;
        return fail(ctx, RB_EINVAL, "sshard: max_kmers_per_round too large for 32-bit record positions");
    }
    int32_t rc = filter_alloc(ctx, RB_BLOOM, std::max<int64_t>(local_d, 32), hd, k, &sh->dbg);
    if (!rc) rc = filter_alloc(ctx, RB_COUNTING, std::max<int64_t>(local_c, 4), hc, k, &sh->cbf);
    cudaError_t er = cudaSuccess;
    if (!rc) {
        const int maxB = std::max(std::max(sh->R, sh->SR), sh->KR) * W;
        const int64_t n_tiles = sh->n_dense / SlShape<6>::TILE + 8;
        er = cudaMalloc(&sh->probe_cursor, (size_t)sh->R * W * kSlPad * 4);
        if (er == cudaSuccess) er = cudaMalloc(&sh->key_cursor, (size_t)sh->KR * W * kSlPad * 4);
        if (er == cudaSuccess) er = cudaMalloc(&sh->raise_cursor, (size_t)sh->SR * W * kSlPad * 4);
        if (er == cudaSuccess) er = cudaMalloc(&sh->cons_cursor, (size_t)maxB * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&sh->cons_rlo, (size_t)maxB * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&sh->sub_cursor, (size_t)n_sub_regions * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&sh->sub_data, ((size_t)n_sub_regions * sh->sub_cap + kSlSpill) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&sh->n_distinct, 64);
        if (er == cudaSuccess) er = cudaMalloc(&sh->pos, ((size_t)sh->n_dense + 8) * kSlNJ * 4);
        if (er == cudaSuccess) er = cudaMalloc(&sh->tile_meta, (size_t)n_tiles * ((size_t)sh->R * W + 1) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&sh->dkey, ((size_t)sh->n_dense + 8) * 8);
        if (er == cudaSuccess) er = cudaMalloc(&sh->dmult, ((size_t)sh->n_dense + 8) * 4);
        if (er == cudaSuccess) er = cudaMalloc(&sh->chunk_prefix, (size_t)(maxB + 2) * 4 + 64);
        if (er == cudaSuccess) er = cudaMalloc(&sh->overflow, 64);
        if (er == cudaSuccess) er = cudaMemsetAsync(sh->overflow, 0, 4, ctx->stream);
    }
    if (rc || er != cudaSuccess) {
        if (!rc) rc = fail(ctx, RB_ENOMEM, std::string("sshard alloc: ") + cudaGetErrorString(er));
        rb_sshard_destroy(sh);
        return rc;
    }
    sh->dbg->in_graph = sh->cbf->in_graph = true;
    *out = sh;
    return RB_OK;
}
// geom: [0] probe regions per rank  [1] records per probe region  [2] key ranges per rank  [3] records per key range
//       [4] raise regions per rank  [5] records per raise region  [6] dbgbf bits of a full share  [7] cbf bytes of a full share
//       [8] local dbgbf bits  [9] local cbf bytes  [10] spill records the caller must add behind every send / receive buffer
extern "C" int32_t rb_sshard_geometry(rb_sshard* sh, int64_t* geom) {
    if (!sh || !geom) return RB_EINVAL;
    geom[0] = sh->R; geom[1] = sh->probe_cap; geom[2] = sh->KR; geom[3] = sh->key_cap; geom[4] = sh->SR; geom[5] = sh->raise_cap;
    geom[6] = (int64_t)sh->sg_route.shard_d << sh->sg_route.dbg_log2; geom[7] = (int64_t)sh->sg_route.shard_c << sh->sg_route.cbf_log2;
    geom[8] = sh->dbg->size; geom[9] = sh->cbf->size; geom[10] = kSlSpill;
    return RB_OK;
}
extern "C" int32_t rb_sshard_filter(rb_sshard* sh, int32_t which, rb_filter** out) {
    if (!sh || !out) return RB_EINVAL;
    *out = which == RB_DBGBF ? sh->dbg : which == RB_CBF ? sh->cbf : nullptr;
    return RB_OK;
}
extern "C" int32_t rb_sshard_overflow(rb_sshard* sh, int32_t* flag) {
    if (!sh || !flag) return RB_EINVAL;
    rb_ctx* ctx = sh->ctx;
    LOCK(ctx);
    int f = 0;
    const int32_t rc = sl_read_flag(ctx, sh->overflow, &f);
    *flag = f;
    return rc;
}

static SlArena ss_producer(rb_sshard* sh, void* data, unsigned int* cursor, int per_rank, uint32_t cap) {
    SlArena a = sl_arena(data, cursor, nullptr, per_rank * sh->W, sl_chunk());
    a.cap = cap;
    return a;
}
// regions received from every rank, in the consumer's order (local region first, source rank second)
static int32_t ss_consumer(rb_sshard* sh, void* data, const uint32_t* recv_cnt, int per_rank, uint32_t cap, int chunk, SlArena* out) {
    rb_ctx* ctx = sh->ctx;
    const int n = per_rank * sh->W;
    RB_LAUNCH((int)div_up(n, kSlThreads), kSlThreads, 0, ctx->stream, ks_order_counts)(recv_cnt, sh->W, per_rank, cap, sh->cons_cursor, sh->cons_rlo);
    LAUNCH_CHECK();
    SlArena a = sl_arena(data, sh->cons_cursor, nullptr, n, chunk);
    a.cap = cap; a.cursor_stride = 1; a.rlo = sh->cons_rlo;
    *out = a;
    RB_LAUNCH(1, kSlThreads, ((size_t)((a.B + 3) & ~3) + 296) * 4, ctx->stream, ks_chunk_prefix)(a, sh->chunk_prefix);
    LAUNCH_CHECK();
    return RB_OK;
}
static int32_t ss_pack_counts(rb_sshard* sh, const SlArena& a, uint32_t* dense) {
    rb_ctx* ctx = sh->ctx;
    RB_LAUNCH((int)div_up(a.B, kSlThreads), kSlThreads, 0, ctx->stream, ks_pack_counts)(a.cursor, a.cursor_stride, a.cap, a.B, dense);
    LAUNCH_CHECK();
    return RB_OK;
}

struct SsRouteUser { rb_sshard* sh; int mode; bool lookup; void* send; int64_t *fh, *rh; int launches; };
static int32_t ss_route_launch(rb_ctx* ctx, const Ingest& ing_in, void* user) {
    SsRouteUser* u = (SsRouteUser*)user;
    rb_sshard* sh = u->sh;
    if (++u->launches > 1 || ing_in.n_pos > sh->n_max) return fail(ctx, RB_EINVAL, "sshard: reads exceed max_kmers_per_round");
    Ingest ing = ing_in;
    ing.out_base = 0;
    const HashMults hm = make_hm(sh->k);
    const bool fast = u->lookup ? sl_uniform_fast_probes<6>(ing, sh->k) : sl_uniform_fast_keys(ing, sh->k);
    int32_t rc;
    if (u->lookup) {
        const SlArena probes = ss_producer(sh, u->send, sh->probe_cursor, sh->R, sh->probe_cap);
        CK(cudaMemsetAsync(probes.cursor, 0, (size_t)probes.B * kSlPad * 4, ctx->stream));
        const size_t sm_sort = TileSort<uint32_t, kSlRoundKmers * kSlNJ>::smem_bytes(probes.B);
        sh->n_items = ing.n_pos; sh->lookup_fast = fast;
        if (fast) {
            const int grid = (int)div_up(ing.n_pos, (int64_t)kSlTile);
            const size_t sm = std::max(sm_sort, PrefixKmerizer<8, SlShape<6>::TILE>::smem_bytes());
            auto kf = ks_route_lookup_u<0, 6>; auto kc = ks_route_lookup_u<2, 6>;
            if (u->mode == RB_MODE_FWD) SL_LAUNCH("ks_route_lookup_u<0>", kf, grid, sm, ing, sh->k, hm, sh->sg_route, probes, sh->pos, sh->tile_meta, u->fh, u->rh, sh->overflow);
            else SL_LAUNCH("ks_route_lookup_u<2>", kc, grid, sm, ing, sh->k, hm, sh->sg_route, probes, sh->pos, sh->tile_meta, u->fh, u->rh, sh->overflow);
        } else {
            const int grid = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kChunk);
            auto kf = ks_route_lookup<0, 6>; auto kc = ks_route_lookup<2, 6>;
            if (u->mode == RB_MODE_FWD) SL_LAUNCH("ks_route_lookup<0>", kf, grid, sm_sort, ing, sh->k, hm, sh->sg_route, probes, sh->pos, sh->tile_meta, u->fh, u->rh, sh->overflow);
            else SL_LAUNCH("ks_route_lookup<2>", kc, grid, sm_sort, ing, sh->k, hm, sh->sg_route, probes, sh->pos, sh->tile_meta, u->fh, u->rh, sh->overflow);
        }
    } else {
        const SlArena keys = ss_producer(sh, u->send, sh->key_cursor, sh->KR, sh->key_cap);
        CK(cudaMemsetAsync(keys.cursor, 0, (size_t)keys.B * kSlPad * 4, ctx->stream));
        const int n_ranges = 1 << sh->lg1, shift = 64 - sh->lg1;
        if (fast) {
            const int grid = (int)div_up(ing.n_pos, (int64_t)kKeyTile);
            const size_t sm = std::max(TileSort<unsigned long long, kKeyE, true>::smem_bytes(keys.B), KeyKmerizer::smem_bytes());
            if (u->mode == RB_MODE_FWD) SL_LAUNCH("ks_route_keys_u<0>", ks_route_keys_u<0>, grid, sm, ing, sh->k, n_ranges, shift, keys, sh->overflow);
            else if (u->mode == RB_MODE_RC) SL_LAUNCH("ks_route_keys_u<1>", ks_route_keys_u<1>, grid, sm, ing, sh->k, n_ranges, shift, keys, sh->overflow);
            else SL_LAUNCH("ks_route_keys_u<2>", ks_route_keys_u<2>, grid, sm, ing, sh->k, n_ranges, shift, keys, sh->overflow);
        } else {
            const int grid = (int)div_up(ing.n_pos, (int64_t)kSlThreads * kChunk);
            const size_t sm = TileSort<unsigned long long, kChunk, true>::smem_bytes(keys.B);
            if (u->mode == RB_MODE_FWD) SL_LAUNCH("ks_route_keys<0>", ks_route_keys<0>, grid, sm, ing, sh->k, n_ranges, shift, keys, sh->overflow);
            else if (u->mode == RB_MODE_RC) SL_LAUNCH("ks_route_keys<1>", ks_route_keys<1>, grid, sm, ing, sh->k, n_ranges, shift, keys, sh->overflow);
            else SL_LAUNCH("ks_route_keys<2>", ks_route_keys<2>, grid, sm, ing, sh->k, n_ranges, shift, keys, sh->overflow);
        }
    }
    return RB_OK;
}
static int32_t ss_route(rb_sshard* sh, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len, int64_t n_reads,
                        int32_t uniform_len, int64_t uniform_stride, int mode, bool lookup, void* send, uint32_t* send_cnt, int64_t* fh, int64_t* rh,
                        int64_t* n_out) {
    rb_ctx* ctx = sh->ctx;
    ReadsArg ra{packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, true};
    SsRouteUser u{sh, mode, lookup, send, fh, rh, 0};
    const int64_t keep = ctx->subbatch_kmers;
    ctx->subbatch_kmers = INT64_MAX / 4;      // one round = one launch
    int64_t n = 0;
    const int32_t rc = for_each_launch(ctx, ra, sh->k, ss_route_launch, &u, &n);
    ctx->subbatch_kmers = keep;
    if (rc) return rc;
    if (n_out) *n_out = n;
    const SlArena a = lookup ? ss_producer(sh, send, sh->probe_cursor, sh->R, sh->probe_cap) : ss_producer(sh, send, sh->key_cursor, sh->KR, sh->key_cap);
    if (u.launches == 0) {   // no k-mer at all on this rank: the exchange still happens, with empty regions
        CK(cudaMemsetAsync(a.cursor, 0, (size_t)a.B * kSlPad * 4, ctx->stream));
        if (lookup) sh->n_items = 0;
    }
    return ss_pack_counts(sh, a, send_cnt);
}
extern "C" int32_t rb_sshard_route_lookup(rb_sshard* sh, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                          int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, uint32_t* send_probes, uint32_t* send_cnt,
                                          int64_t* fhash, int64_t* rhash, int64_t* n_out) {
    if (!sh || !send_probes || !send_cnt) return RB_EINVAL;
    LOCK(sh->ctx);
    return ss_route(sh, packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, sh->stranded ? RB_MODE_FWD : RB_MODE_CANON, true, send_probes,
                    send_cnt, fhash, sh->stranded ? nullptr : rhash, n_out);
}
extern "C" int32_t rb_sshard_route_keys(rb_sshard* sh, const uint64_t* packed, const uint32_t* mask, const int64_t* read_off, const int32_t* read_len,
                                        int64_t n_reads, int32_t uniform_len, int64_t uniform_stride, uint32_t flags, unsigned long long* send_keys,
                                        uint32_t* send_cnt, int64_t* n_out) {
    if (!sh || !send_keys || !send_cnt) return RB_EINVAL;
    LOCK(sh->ctx);
    const int mode = !sh->stranded ? RB_MODE_CANON : ((flags & RB_REVCOMP) ? RB_MODE_RC : RB_MODE_FWD);
    return ss_route(sh, packed, mask, read_off, read_len, n_reads, uniform_len, uniform_stride, mode, false, send_keys, send_cnt, nullptr, nullptr, n_out);
}
// owner side: probes received from every rank -> answers at the same positions of recv_ans
extern "C" int32_t rb_sshard_apply(rb_sshard* sh, const uint32_t* recv_probes, const uint32_t* recv_cnt, uint8_t* recv_ans, int32_t set_bits) {
    if (!sh || !recv_probes || !recv_cnt || !recv_ans) return RB_EINVAL;
    rb_ctx* ctx = sh->ctx;
    LOCK(ctx);
    SlArena a;
    int32_t rc = ss_consumer(sh, (void*)recv_probes, recv_cnt, sh->R, sh->probe_cap, sl_chunk(), &a);
    if (rc) return rc;
    const size_t sm_pre = (size_t)(a.B + 1) * 4;
    int grid = 0;
    if (set_bits) {
        rc = sl_persistent_grid(ctx, ks_apply_probes<1>, sm_pre, &grid); if (rc) return rc;
        SL_LAUNCH("ks_apply_probes<1>", ks_apply_probes<1>, grid, sm_pre, a, sh->chunk_prefix, sh->sg_apply, sh->dbg->dev, sh->cbf->dev, recv_ans);
    } else {
        rc = sl_persistent_grid(ctx, ks_apply_probes<0>, sm_pre, &grid); if (rc) return rc;
        SL_LAUNCH("ks_apply_probes<0>", ks_apply_probes<0>, grid, sm_pre, a, sh->chunk_prefix, sh->sg_apply, sh->dbg->dev, sh->cbf->dev, recv_ans);
    }
    return RB_OK;
}
extern "C" int32_t rb_sshard_combine_lookup(rb_sshard* sh, const uint8_t* home_ans, float* counts) {
    if (!sh || !home_ans || !counts) return RB_EINVAL;
    rb_ctx* ctx = sh->ctx;
    LOCK(ctx);
    if (sh->n_items == 0) return RB_OK;
    int32_t rc;
    const int B = sh->R * sh->W;
    const size_t sm_ans = TileAnswers::smem_bytes(B, kSlTile * kSlNJ);
    auto k1 = ks_combine_lookup<1, 6>; auto k0 = ks_combine_lookup<0, 6>;
    if (sh->lookup_fast) SL_LAUNCH("ks_combine_lookup<1>", k1, (int)div_up(sh->n_items, (int64_t)kSlTile), sm_ans, sh->pos, sh->tile_meta, B, home_ans, sh->n_items, sh->hd, sh->hc, counts, (int64_t)0);
    else SL_LAUNCH("ks_combine_lookup<0>", k0, (int)div_up(sh->n_items, (int64_t)kSlThreads * kChunk), sm_ans, sh->pos, sh->tile_meta, B, home_ans, sh->n_items, sh->hd, sh->hc, counts, (int64_t)0);
    return RB_OK;
}
// home side: keys of this rank's hash ranges from every rank -> distinct keys with multiplicities
extern "C" int32_t rb_sshard_dedup(rb_sshard* sh, const unsigned long long* recv_keys, const uint32_t* recv_cnt) {
    if (!sh || !recv_keys || !recv_cnt) return RB_EINVAL;
    rb_ctx* ctx = sh->ctx;
    LOCK(ctx);
    SlArena keys;
    int32_t rc = ss_consumer(sh, (void*)recv_keys, recv_cnt, sh->KR, sh->key_cap, kSlThreads * kKeyE, &keys);
    if (rc) return rc;
    const int n_sub = 1 << sh->sub_bits;
    const int n_sub_regions = sh->KR << sh->sub_bits;
    SlArena subs = sl_arena(sh->sub_data, sh->sub_cursor, nullptr, n_sub_regions, 0);
    subs.cap = sh->sub_cap; subs.cursor_stride = 1;
    CK(cudaMemsetAsync(subs.cursor, 0, (size_t)n_sub_regions * 4, ctx->stream));
    int grid = 0;
    const size_t sm_split = TileSort<unsigned long long, kKeyE, true>::smem_bytes(n_sub) + (size_t)(keys.B + 1) * 4;
    rc = sl_stream_grid(ctx, ks_split_keys, sm_split, &grid);
    if (rc) return rc;
    SL_LAUNCH("ks_split_keys", ks_split_keys, grid, sm_split, keys, sh->chunk_prefix, sh->sub_bits, 64 - sh->lg1 - sh->sub_bits, sh->W, subs, sh->overflow);
    CK(cudaMemsetAsync(sh->n_distinct, 0, 4, ctx->stream));
    const size_t sm_dedup = (size_t)kSlDedupSlots * 12;
    rc = sl_stream_grid(ctx, ks_dedup, sm_dedup, &grid);
    if (rc) return rc;
    SL_LAUNCH("ks_dedup", ks_dedup, std::min(grid, n_sub_regions), sm_dedup, subs, n_sub_regions, sh->lg1 + sh->sub_bits, sh->dkey, sh->dmult, sh->n_distinct,
              (unsigned int)sh->n_dense, sh->overflow, SpillTable{nullptr, nullptr, 0, 0});
    return RB_OK;
}
extern "C" int32_t rb_sshard_emit_probes(rb_sshard* sh, int32_t with_cbf, uint32_t* send_probes, uint32_t* send_cnt) {
    if (!sh || !send_probes || !send_cnt) return RB_EINVAL;
    rb_ctx* ctx = sh->ctx;
    LOCK(ctx);
    int32_t rc;
    const HashMults hm = make_hm(sh->k);
    const SlArena probes = ss_producer(sh, send_probes, sh->probe_cursor, sh->R, sh->probe_cap);
    CK(cudaMemsetAsync(probes.cursor, 0, (size_t)probes.B * kSlPad * 4, ctx->stream));
    const size_t sm_sort = TileSort<uint32_t, kSlRoundKmers * kSlNJ>::smem_bytes(probes.B);
    const int grid_d = (int)div_up(sh->n_dense, (int64_t)kSlTile);
    SL_LAUNCH("ks_emit_probes", ks_emit_probes<6>, grid_d, sm_sort, sh->dkey, sh->n_distinct, hm, sh->sg_route, with_cbf, probes, sh->pos, sh->tile_meta, sh->overflow);
    return ss_pack_counts(sh, probes, send_cnt);
}
extern "C" int32_t rb_sshard_combine_insert(rb_sshard* sh, const uint8_t* home_ans, int32_t policy, uint32_t* send_raises, uint32_t* send_cnt) {
    if (!sh || !home_ans || !send_raises || !send_cnt) return RB_EINVAL;
    rb_ctx* ctx = sh->ctx;
    LOCK(ctx);
    int32_t rc;
    const HashMults hm = make_hm(sh->k);
    const SlArena raises = ss_producer(sh, send_raises, sh->raise_cursor, sh->SR, sh->raise_cap);
    CK(cudaMemsetAsync(raises.cursor, 0, (size_t)raises.B * kSlPad * 4, ctx->stream));
    const uint64_t seed = ctx->rng_seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(ctx->launches + 1);
    const int B = sh->R * sh->W;
    const size_t sm_r = std::max(TileSort<uint32_t, kSlRoundKmers * kSlMaxH>::smem_bytes(raises.B), TileAnswers::smem_bytes(B, kSlTile * kSlNJ));
    const int grid_d = (int)div_up(sh->n_dense, (int64_t)kSlTile);
    SL_LAUNCH("ks_combine_insert", ks_combine_insert<6>, grid_d, sm_r, sh->dkey, sh->dmult, sh->n_distinct, sh->pos, sh->tile_meta, B, home_ans, hm, sh->sg_route, policy, seed,
              raises, sh->overflow);
    return ss_pack_counts(sh, raises, send_cnt);
}
extern "C" int32_t rb_sshard_apply_raises(rb_sshard* sh, const uint32_t* recv_raises, const uint32_t* recv_cnt) {
    if (!sh || !recv_raises || !recv_cnt) return RB_EINVAL;
    rb_ctx* ctx = sh->ctx;
    LOCK(ctx);
    SlArena a;
    int32_t rc = ss_consumer(sh, (void*)recv_raises, recv_cnt, sh->SR, sh->raise_cap, sl_chunk(), &a);
    if (rc) return rc;
    const size_t sm_rp = (size_t)(a.B + 1) * 4;
    int grid = 0;
    rc = sl_persistent_grid(ctx, ks_apply_raises, sm_rp, &grid);
    if (rc) return rc;
    SL_LAUNCH("ks_apply_raises", ks_apply_raises, grid, sm_rp, a, sh->chunk_prefix, sh->sg_apply, sh->cbf->dev);
    claim_invalidate(ctx);
    return RB_OK;
}
