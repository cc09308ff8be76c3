"""bench.py --config 3 | 4: the other two single-GPU workloads of BASELINE.json, same metric, same JSON line, their own algorithmic bytes.

configs[3]  2x150 bp reads, STRANDED, k=35, read paired-k-mer filter on (the fragment graph path): left mates inserted forward, right mates
            through the reverse-complement iterators (`-revcomp-right`, RNABloom.java:551-634), every read also stores its paired k-mers at
            distance d = max(1, 150 - 35 - 10) = 105 (RNABloom.java:1022) in the rpkbf; look-ups = graph.getKmers of every read.
            A = 32 (h_d + h_c) + 32 h_p (L-k-d+1)/(L-k+1) + input  per inserted k-mer,  32 (h_d + h_c) + input  per looked-up k-mer.
configs[4]  ONT-like long reads (500..3.5 kb, mean 2 kb; 2 % substitutions, 1.5 % insertions, 1.5 % deletions), k=17, canonical, 16 GiB
            counting filter (+ 16 GiB dbgbf), ragged read layout.  A = 32 (h_d + h_c) + input per k-mer and phase.
Reads are synthetic (counter-based generators with a bit-identical twin in the CPU checker); `cpu_baseline` = the oracle port on a bounded
sample, as in the default config."""
import ctypes as C
import json
import os
import time

import numpy as np

import bench as B1

SEED = B1.SEED


def _np_mix64(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def long_read_lengths(seed, first, n):
    """numpy restatement of rb_synth_long_read_len (csrc/rb_kernels.cuh synth_long_len)."""
    with np.errstate(over="ignore"):
        r = np.arange(first, first + n, dtype=np.uint64)
        h = _np_mix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(2) * r)
    return (500 + (h % np.uint64(1000)) + ((h >> np.uint64(20)) % np.uint64(1000)) + ((h >> np.uint64(40)) % np.uint64(1000))).astype(np.int64)


def _finish(line_base, args, ctx, prof, t_ins, t_look, n_ins, n_look, a_ins, a_look, wall, launches, clocks, e2e, cpu, engine):
    hbm, peak_src = B1.peaks()
    t_total = t_ins + t_look
    value = 0.5 * (n_ins + n_look) / (t_total * 1e-3)
    ins_dom = t_ins >= t_look
    kern_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items()}
    top = max(prof.items(), key=lambda kv: kv[1][0]) if prof else ("?", (0.0, 1))
    achieved = (n_ins * a_ins / (t_ins * 1e-3) if ins_dom else n_look * a_look / (t_look * 1e-3)) / 1e9
    roofline = {"bound": "hbm", "kernel": ("insert" if ins_dom else "lookup") + " rounds (kernel chain, see kernels_ms_per_step)", "achieved": achieved,
                "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_kmer": a_ins if ins_dom else a_look, "insert_gkmers_s": n_ins / t_ins / 1e6, "lookup_gkmers_s": n_look / t_look / 1e6,
                "step_frac": (n_ins * a_ins + n_look * a_look) / (t_total * 1e-3) / 1e9 / hbm, "engine": engine, "kernels_ms_per_step": kern_ms,
                "top_kernel": top[0], "top_kernel_share_of_step": top[1][0] / t_total}
    line = dict(line_base, value=value, ms_per_step=t_total / args.steps, clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roofline,
                cpu_baseline=cpu, wall_s_timed_region=wall)
    print(json.dumps(line), flush=True)


def _base_line(args, workload, extra):
    cfg = {"workload": workload, "l2": "filters and every step's fresh read batch exceed the 126 MB L2; no reuse between timed iterations",
           "sharding": "none (1 GPU)"}
    cfg.update(extra)
    return {"metric": "k-mers/s (insert+lookup), BASELINE.json metric on " + workload.split(":")[0], "unit": "k-mers/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": cfg}


# ---- configs[3] -------------------------------------------------------------------------------------------------------------------------------
def run_config3(args):
    import rnabloom_b200 as rb
    K, L, STRIDE, H, D = 35, 150, 160, 3, 105
    DBG, CBF, PK = 1 << 36, 1 << 33, 1 << 35
    kpr, ppr = L - K + 1, L - K - D + 1
    os.environ["RB_ENGINE"] = args.engine
    args.warmup = max(args.warmup, 3)
    ctx = rb.Context(0)
    g = rb.BloomFilterDeBruijnGraph(ctx, DBG, CBF, PK, H, H, H, K, True, True)
    g.setPairedKmerDistances(D, -1)
    n_reads = args.reads_per_step - args.reads_per_step % 2
    half = n_reads // 2
    nk = n_reads * kpr
    total_steps = args.warmup + args.steps
    words = n_reads * STRIDE // 32
    n_batches = min(total_steps, max(1, 100_000_000 // n_reads))
    batches = [ctx.dev_alloc(words * 8 + 64) for _ in range(n_batches)]
    for s, p in enumerate(batches):
        ctx.synth_reads_dev(SEED + 3, B1.GENOME, s * n_reads, n_reads, L, B1.ERR_PPM, STRIDE, p)
    counts_dev = ctx.dev_alloc(nk * 4)
    ctx.sync()
    right_off = half * (STRIDE // 4)

    def step(p):
        ctx.timer_start()
        g.addReadsDev(p, half, L, STRIDE, flags=rb.STORE_READ_PAIRS)                         # left mates: forward strand
        g.addReadsDev(p + right_off, half, L, STRIDE, flags=rb.REVCOMP | rb.STORE_READ_PAIRS)   # right mates: -revcomp-right
        t_i = ctx.timer_stop()
        ctx.timer_start()
        g.getKmersDev(p, n_reads, L, STRIDE, counts_dev)
        t_l = ctx.timer_stop()
        return t_i, t_l

    for s in range(args.warmup):
        step(batches[s % n_batches])
    ctx.sync()
    ctx.profile_enable(True)
    ctx.profile_read()
    sampler = B1.ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    l0 = ctx.kernel_launches()
    t_ins = t_look = 0.0
    wall0 = time.perf_counter()
    for s in range(args.warmup, total_steps):
        a, b = step(batches[s % n_batches])
        t_ins += a
        t_look += b
    ctx.sync()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches() - l0
    clocks = sampler.stop()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle.binding import F_REVCOMP, F_STORE_READ_PAIRS, Oracle, OracleGraph
        orc = Oracle()
        threads = os.cpu_count() or 1
        og = OracleGraph(orc, DBG, CBF, PK, H, H, H, K, True, True)
        og.set_distances(D, -1)
        for f in (orc.lib.orc_graph_dbgbf(og.g), orc.lib.orc_graph_rpkbf(og.g)):   # touch every page before timing
            orc.lib.orc_bf_empty(f)
        orc.lib.orc_cbf_empty(orc.lib.orc_graph_cbf(og.g))
        sample = 200_000
        reads = orc.synth_reads(SEED + 3, B1.GENOME, 0, sample, L, B1.ERR_PPM)
        t0 = time.perf_counter()
        km, _ = og.run_mt(reads[: sample // 2], F_STORE_READ_PAIRS, False, threads)
        km2, _ = og.run_mt(reads[sample // 2:], F_STORE_READ_PAIRS | F_REVCOMP, False, threads)
        t1 = time.perf_counter()
        og.run_mt(reads, 0, True, threads)
        t2 = time.perf_counter()
        og.close()
        cpu = {"value": (km + km2) / (t2 - t0), "unit": "k-mers/s", "cores": threads, "kind": "port",
               "sample": "%d reads insert (+ read pairs) then lookup, %d threads, full-size filters; C restatement of the Java path, not JVM" % (sample, threads),
               "insert_kmers_s": (km + km2) / (t1 - t0), "lookup_kmers_s": (km + km2) / (t2 - t1)}
    a_in = 32.0 * 2 * H + 32.0 * H * ppr / kpr + L / (8.0 * kpr)
    a_lk = 32.0 * 2 * H + L / (8.0 * kpr)
    base = _base_line(args, "BASELINE.json configs[3]: 2x150 bp reads, stranded, k=35, read paired-k-mer filter (d=105), 8 GiB Bloom + 8 GiB counting + 4 GiB rpkbf, 3 hashes",
                      {"reads_per_step_per_gpu": n_reads, "read_len": L, "k": K, "stranded": True, "read_pair_distance": D, "dbgbf_bits": DBG, "cbf_bytes": CBF,
                       "rpkbf_bits": PK, "num_hash": H, "pairs_per_read": ppr, "engine": args.engine,
                       "note": "uniform-genome reads (the Zipf transcriptome generator of SURVEY 8d is not built); pair inserts run on the direct kernels"})
    _finish(base, args, ctx, prof, t_ins, t_look, nk * args.steps, nk * args.steps, a_in, a_lk, wall, launches, clocks, None, cpu, args.engine)
    g.destroy()
    ctx.close()


# ---- configs[4] -------------------------------------------------------------------------------------------------------------------------------
def run_config4(args):
    import rnabloom_b200 as rb
    K, H = 17, 3
    DBG, CBF = 1 << 37, 1 << 34
    SUB, INS, DEL = 20000, 15000, 15000
    GENOME = 500_000_000
    os.environ["RB_ENGINE"] = args.engine
    args.warmup = max(args.warmup, 3)
    ctx = rb.Context(0)
    L_ = ctx.L
    g = rb.BloomFilterDeBruijnGraph(ctx, DBG, CBF, 64, H, H, 1, K, False, False)
    n_reads = max(1000, args.reads_per_step // 16)          # ~2 kb reads: 250 k reads ~ 500 M k-mers per step
    total_steps = args.warmup + args.steps
    n_batches = min(total_steps, 40)
    batches = []
    for s in range(n_batches):
        lens = long_read_lengths(SEED + 4, s * n_reads, n_reads)
        words = (lens + 31) // 32
        off = np.zeros(n_reads, dtype=np.int64)
        off[1:] = np.cumsum(words[:-1]) * 32
        n_words = int(words.sum())
        d_off, d_len, d_packed = ctx.dev_alloc(n_reads * 8 + 64), ctx.dev_alloc(n_reads * 4 + 64), ctx.dev_alloc(n_words * 8 + 64)
        ctx.h2d(d_off, off)
        ctx.h2d(d_len, lens.astype(np.int32))
        ctx.check(L_.rb_synth_long_reads_dev(ctx.h, SEED + 4, GENOME, s * n_reads, n_reads, SUB, INS, DEL, C.c_void_p(d_off), C.c_void_p(d_packed)))
        batches.append((d_packed, d_off, d_len, int(np.maximum(lens - K + 1, 0).sum())))
    max_nk = max(b[3] for b in batches)
    counts_dev = ctx.dev_alloc(max_nk * 4 + 64)
    ctx.sync()

    def step(b):
        n = C.c_int64()
        reads = (C.c_void_p(b[0]), None, C.c_void_p(b[1]), C.c_void_p(b[2]), n_reads, 0, 0)
        ctx.timer_start()
        ctx.check(L_.rb_graph_add_reads_dev(g.h, *reads, 0, C.byref(n)))
        t_i = ctx.timer_stop()
        assert n.value == b[3], (n.value, b[3])
        ctx.timer_start()
        ctx.check(L_.rb_graph_count_reads_dev(g.h, *reads, C.c_void_p(counts_dev), None, None, C.byref(n)))
        t_l = ctx.timer_stop()
        return t_i, t_l

    for s in range(args.warmup):
        step(batches[s % n_batches])
    ctx.sync()
    ctx.profile_enable(True)
    ctx.profile_read()
    sampler = B1.ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    l0 = ctx.kernel_launches()
    t_ins = t_look = 0.0
    kmers = 0
    wall0 = time.perf_counter()
    for s in range(args.warmup, total_steps):
        b = batches[s % n_batches]
        a, c = step(b)
        t_ins += a
        t_look += c
        kmers += b[3]
    ctx.sync()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches() - l0
    clocks = sampler.stop()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle.binding import Oracle, OracleGraph
        orc = Oracle()
        threads = os.cpu_count() or 1
        og = OracleGraph(orc, DBG, CBF, 64, H, H, 1, K, False, False)
        orc.lib.orc_bf_empty(orc.lib.orc_graph_dbgbf(og.g))   # touch every page before timing
        orc.lib.orc_cbf_empty(orc.lib.orc_graph_cbf(og.g))
        bases, off = orc.synth_long_reads(SEED + 4, GENOME, 0, 15000, SUB, INS, DEL)
        t0 = time.perf_counter()
        km, _ = og.run_mt_ragged(bases, off, 0, False, threads)
        t1 = time.perf_counter()
        og.run_mt_ragged(bases, off, 0, True, threads)
        t2 = time.perf_counter()
        og.close()
        cpu = {"value": km / (t2 - t0), "unit": "k-mers/s", "cores": threads, "kind": "port",
               "sample": "15000 long reads (%.1f M k-mers) insert then lookup, %d threads, full-size filters; C restatement of the Java path, not JVM" % (km / 1e6, threads),
               "insert_kmers_s": km / (t1 - t0), "lookup_kmers_s": km / (t2 - t1)}
    mean_len = float(np.mean([b[3] for b in batches])) / n_reads + K - 1
    a_k = 32.0 * 2 * H + mean_len / (8.0 * (mean_len - K + 1))
    base = _base_line(args, "BASELINE.json configs[4]: ONT-like long reads (mean 2 kb, 2 % sub + 1.5 % ins + 1.5 % del), k=17, 16 GiB Bloom + 16 GiB counting filter, 3 hashes, canonical",
                      {"reads_per_step_per_gpu": n_reads, "mean_read_len": mean_len, "k": K, "dbgbf_bits": DBG, "cbf_bytes": CBF, "num_hash": H,
                       "genome_len": GENOME, "engine": args.engine, "layout": "ragged (read_off / read_len), rolling-walker k-merizer"})
    _finish(base, args, ctx, prof, t_ins, t_look, kmers, kmers, a_k, a_k, wall, launches, clocks, None, cpu, args.engine)
    g.destroy()
    ctx.close()
