#!/usr/bin/env python
"""bench.py -- k-mers/s (insert + lookup) of the Bloom-filter de Bruijn graph hot path at k=25 on synthetic 2x150 bp reads.

One step = one batch of synthetic reads through graph.add (insert) and graph.getKmers (lookup):
    value = k-mers per step * steps / device time          (inputs resident in HBM, CUDA events on the library's stream)
    e2e   = the same through the host-pointer C-ABI calls  (pinned host buffers; H2D of the packed reads and D2H of the
            per-k-mer counts inside the timed region)
Workload = BASELINE.json configs[1]: 2x150 bp reads, k=25, 8 GiB Bloom (2^36 bits) + 8 GiB counting filter (2^33 bytes),
3 hashes each, canonical (unstranded) k-mers; reads from a virtual 3 Gb genome, 0.5 % substitutions.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 25
READ_LEN = 150
STRIDE = 160                      # bases per read slot in the uniform ingest layout (multiple of 32)
HD = HC = 3
GENOME = 3_000_000_000
ERR_PPM = 5000
SEED = 20261017
DBG_BITS = 1 << 36                # 8 GiB
CBF_BYTES = 1 << 33               # 8 GiB
KMERS_PER_READ = READ_LEN - K + 1
# SURVEY.md section 8d / BASELINE.md section 3: algorithmic bytes per k-mer (32 B sector per probe, reads only)
A_INSERT = 32.0 * (HD + HC) + READ_LEN / (4.0 * KMERS_PER_READ) / 2
A_LOOKUP = 32.0 * (HD + HC) + READ_LEN / (4.0 * KMERS_PER_READ) / 2
A_STEP = A_INSERT + A_LOOKUP      # 384.3 B per (insert + lookup) k-mer


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_sample(n_reads, threads, lookup_too=True, dbg_bits=DBG_BITS, cbf_bytes=CBF_BYTES, first_read=0, graph=None, orc=None):
    """The oracle port (reference-faithful threaded C restatement of the Java path) on a bounded sample of the workload."""
    from oracle.binding import Oracle, OracleGraph
    orc = orc or Oracle()
    own = graph is None
    if own:
        graph = OracleGraph(orc, dbg_bits, cbf_bytes, 64, HD, HC, 1, K, False, False)
        orc.lib.orc_bf_empty(orc.lib.orc_graph_dbgbf(graph.g))   # touch every page before timing
        orc.lib.orc_cbf_empty(orc.lib.orc_graph_cbf(graph.g))
    reads = orc.synth_reads(SEED, GENOME, first_read, n_reads, READ_LEN, ERR_PPM)
    t0 = time.perf_counter()
    km, _ = graph.run_mt(reads, 0, False, threads)
    t1 = time.perf_counter()
    if lookup_too:
        graph.run_mt(reads, 0, True, threads)
    t2 = time.perf_counter()
    if own:
        graph.close()
    return km, t1 - t0, t2 - t1


def host_filter_sizes():
    """Full-size filters need ~16 GiB of host RAM for the CPU arm; scale down (and say so) on small hosts."""
    try:
        avail = int(next(line.split()[1] for line in open("/proc/meminfo") if line.startswith("MemAvailable"))) * 1024
    except Exception:
        avail = 8 << 30
    shift = 0
    while (DBG_BITS >> (3 + shift)) + (CBF_BYTES >> shift) > 0.6 * avail:
        shift += 1
    return DBG_BITS >> shift, CBF_BYTES >> shift, shift


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle port; no JVM exists in this image) with every host thread."""
    if rank != 0:
        return
    from oracle.binding import Oracle, OracleGraph
    threads = os.cpu_count() or 1
    dbg_bits, cbf_bytes, shift = host_filter_sizes()
    orc = Oracle()
    g = OracleGraph(orc, dbg_bits, cbf_bytes, 64, HD, HC, 1, K, False, False)
    orc.lib.orc_bf_empty(orc.lib.orc_graph_dbgbf(g.g))
    orc.lib.orc_cbf_empty(orc.lib.orc_graph_cbf(g.g))
    km, ti, tl = cpu_sample(20_000, threads, graph=g, orc=orc)              # calibration
    rate = km / (ti + tl)
    n_reads = int(max(20_000, min(2_000_000, rate * 4.0 / KMERS_PER_READ)))  # ~4 s per step
    first = 20_000
    for _ in range(args.warmup):
        cpu_sample(n_reads, threads, graph=g, orc=orc, first_read=first)
        first += n_reads
    t_total, kmers = 0.0, 0
    for _ in range(args.steps):
        km, ti, tl = cpu_sample(n_reads, threads, graph=g, orc=orc, first_read=first)
        first += n_reads
        t_total += ti + tl
        kmers += km
    g.close()
    value = kmers / t_total
    sample = "%d reads/step insert+lookup, filters %s" % (n_reads, "full size (8 GiB + 8 GiB)" if shift == 0 else "scaled 1/%d to fit host RAM" % (1 << shift))
    line = {"impl": "reference", "metric": "k-mers/s (insert+lookup) at k=25, 2x150 bp reads", "value": value, "unit": "k-mers/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": config_dict(n_reads, 1),
            "cpu_baseline": {"value": value, "unit": "k-mers/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def config_dict(reads_per_step, n_gpus):
    return {"workload": "BASELINE.json configs[1]: 50M synthetic 2x150 bp read pairs, k=25, 8 GiB Bloom + 8 GiB counting filter, 3 hashes, canonical k-mers",
            "reads_per_step_per_gpu": reads_per_step, "read_len": READ_LEN, "k": K, "dbgbf_bits": DBG_BITS, "cbf_bytes": CBF_BYTES,
            "num_hash": HD, "genome_len": GENOME, "substitution_ppm": ERR_PPM,
            "l2": "filters (16 GiB) and every step's fresh read batch (%d MB) exceed the 126 MB L2; no reuse between timed iterations"
                  % (reads_per_step * STRIDE // 4 // 1000000),
            "sharding": "none (1 GPU)" if n_gpus == 1 else "filters sharded by index range over %d GPUs" % n_gpus}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=22)   # 3 + 22 steps of 4 M reads = the 100 M reads (50 M pairs) of configs[1]
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=4_000_000)
    ap.add_argument("--subbatch-kmers", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the oracle comparison of the sharded path after the timed region")
    ap.add_argument("--engine", default=os.environ.get("RB_ENGINE", "sliced"), choices=["direct", "sliced", "auto"])
    ap.add_argument("--sharded", action="store_true", help="experiments only: run the sharded pipeline even on one GPU")
    ap.add_argument("--genome", type=int, default=GENOME, help="experiments only: virtual genome length (coverage knob)")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3, 4],
                    help="BASELINE.json configs[i]: 1 = the headline workload (default), 3 = stranded k=35 + read paired k-mers, 4 = ONT-like long reads k=17 "
                         "(single GPU; bench_configs.py)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.config != 1:
        if world > 1:
            raise SystemExit("--config 3 / 4 are single-GPU workloads")
        import bench_configs
        (bench_configs.run_config3 if args.config == 3 else bench_configs.run_config4)(args)
        return
    if world > 1 or args.sharded:
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            os.environ.setdefault("RANK", "0")
            os.environ.setdefault("WORLD_SIZE", "1")
        from bench_multi import run_sharded   # hash-sharded filters + all-to-all (rna-bloom_b200/sharded.py)
        run_sharded(args, rank, world, local_rank)
        return

    import rnabloom_b200 as rb
    os.environ["RB_ENGINE"] = args.engine
    args.warmup = max(args.warmup, 3)
    ctx = rb.Context(local_rank)
    if args.subbatch_kmers:
        ctx.set_subbatch_kmers(args.subbatch_kmers)
    g = rb.BloomFilterDeBruijnGraph(ctx, DBG_BITS, CBF_BYTES, 64, HD, HC, 1, K, False, False)
    n_reads = args.reads_per_step
    nk = n_reads * KMERS_PER_READ
    total_steps = args.warmup + args.steps
    words = n_reads * STRIDE // 32
    n_batches = min(total_steps, max(1, 100_000_000 // n_reads))   # the data set has 100 M reads; longer runs wrap around
    batches = [ctx.dev_alloc(words * 8 + 64) for _ in range(n_batches)]
    for s, p in enumerate(batches):
        ctx.synth_reads_dev(SEED, args.genome, s * n_reads, n_reads, READ_LEN, ERR_PPM, STRIDE, p)
    counts_dev = ctx.dev_alloc(nk * 4)
    ctx.sync()

    def step(p):
        ctx.timer_start()
        g.addReadsDev(p, n_reads, READ_LEN, STRIDE)
        t_i = ctx.timer_stop()
        ctx.timer_start()
        g.getKmersDev(p, n_reads, READ_LEN, STRIDE, counts_dev)
        t_l = ctx.timer_stop()
        return t_i, t_l

    for s in range(args.warmup):
        step(batches[s % n_batches])
    ctx.sync()
    ctx.profile_enable(True)   # CUDA events around every kernel launch of the timed region (a few microseconds per launch)
    ctx.profile_read()
    sampler = ClockSampler(local_rank)
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
    t_total_ms = t_ins + t_look
    value = nk * args.steps / (t_total_ms * 1e-3)
    hbm, peak_src = peaks()

    # ---- end to end: host-pointer C-ABI calls, pinned host buffers, H2D + D2H inside the timed region ----------------
    e2e = None
    if not args.no_e2e:
        e_steps = max(2, min(args.steps, 8))
        h_packed = [ctx.host_alloc(words * 8, np.uint64) for _ in range(e_steps + 4)]   # 1 + 2 for the blocking calls, 1 + e_steps pipelined
        tmp = ctx.dev_alloc(words * 8 + 64)
        for i, hp in enumerate(h_packed):   # fresh reads (ids after the timed ones), generated on the device, parked in pinned host memory
            ctx.synth_reads_dev(SEED, args.genome, (total_steps + i) * n_reads, n_reads, READ_LEN, ERR_PPM, STRIDE, tmp)
            ctx.sync()
            ctx.d2h(hp, tmp)
        ctx.dev_free(tmp)
        h_counts = [ctx.host_alloc(nk * 4, np.float32) for _ in range(2)]
        reads = [rb.PackedReads(hp, None, None, None, n_reads, READ_LEN, STRIDE) for hp in h_packed]
        from rnabloom_b200.filters import _ptr
        import ctypes as C

        # (a) every call waits for its results: rb_graph_add_reads + rb_graph_count_reads
        def e2e_step(pr):
            n = C.c_int64()
            ctx.check(ctx.L.rb_graph_add_reads(g.h, *pr.args(), 0, C.byref(n)))
            ctx.check(ctx.L.rb_graph_count_reads(g.h, *pr.args(), _ptr(h_counts[0]), None, None, C.byref(n)))
            return float(h_counts[0][:1024].sum())
        e2e_step(reads[0])
        t0 = time.perf_counter()
        for i in range(1, 3):
            e2e_step(reads[i])
        te_sync = (time.perf_counter() - t0) / 2

        # (b) the headline: the same work with rb_graph_count_reads_async -- the D2H of step i's counts (4 B per k-mer: 2 GB) runs behind
        # the insert of step i + 1; two pinned result arrays used alternately; every step's result is read on the host after its ticket
        def e2e_pipelined(idx):
            n = C.c_int64()
            prev, acc = None, 0.0
            for i in idx:
                ctx.check(ctx.L.rb_graph_add_reads(g.h, *reads[i].args(), 0, C.byref(n)))
                t = g.getKmersAsync(reads[i], h_counts[i & 1])
                if prev is not None:
                    ctx.wait(prev[0])
                    acc += float(prev[1][:1024].sum())
                prev = (t, h_counts[i & 1])
            ctx.wait(prev[0])
            return acc + float(prev[1][:1024].sum())
        e2e_pipelined([3])
        t0 = time.perf_counter()
        chk = e2e_pipelined(list(range(4, e_steps + 4)))
        te = time.perf_counter() - t0
        assert chk >= 1024.0 * e_steps   # every k-mer of a step was inserted before it was looked up: count >= 1
        piped = {"value": nk * e_steps / te, "ms_per_step": 1e3 * te / e_steps, "steps": e_steps,
                 "calls": "rb_graph_add_reads + rb_graph_count_reads_async / rb_ctx_wait (pinned host buffers, results double-buffered)"}
        block = {"value": nk / te_sync, "ms_per_step": 1e3 * te_sync, "steps": 2, "calls": "rb_graph_add_reads + rb_graph_count_reads"}
        # both are the library's public calls on the same work with the same copies; the line's e2e is the better way to call it on this box
        # (the pipelined calls lose when the driver shares one copy engine between the two directions), the other one stays beside it
        best = piped if piped["value"] >= block["value"] else block
        e2e = {"value": best["value"], "unit": "k-mers/s", "h2d_bytes_per_step": 2 * words * 8, "d2h_bytes_per_step": nk * 4,
               "steps": best["steps"], "ms_per_step": best["ms_per_step"], "calls": best["calls"], "pipelined_calls": piped, "blocking_calls": block}

        # the same through the ASCII entry points -- what the Java insert workers and graph.getKmers(String) call (RNABloom.java:551-634,
        # graph :1224-1234): ASCII bases in host memory in, segmentation + 2-bit packing on the GPU, counts back to host memory
        n_a = min(n_reads, 1_000_000)
        w_a = n_a * STRIDE // 32
        pk = np.zeros(w_a, dtype=np.uint64)
        tmp = ctx.dev_alloc(w_a * 8 + 64)
        ctx.synth_reads_dev(SEED, args.genome, (total_steps + e_steps + 5) * n_reads, n_a, READ_LEN, ERR_PPM, STRIDE, tmp)
        ctx.sync()
        ctx.d2h(pk, tmp)
        ctx.dev_free(tmp)
        codes = ((pk[:, None] >> (2 * np.arange(32, dtype=np.uint64))[None, :]) & np.uint64(3)).astype(np.uint8).reshape(n_a, STRIDE)[:, :READ_LEN]
        ascii_bases = ctx.host_alloc(n_a * READ_LEN, np.uint8)   # the chunk buffer a worker fills: pinned (rb_host_alloc; Java: a direct ByteBuffer over it)
        ascii_bases[:] = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].reshape(-1)
        ascii_off = np.arange(n_a + 1, dtype=np.int64) * READ_LEN
        a_counts = ctx.host_alloc(n_a * KMERS_PER_READ * 4, np.float32)
        na = C.c_int64()

        def ascii_step():
            ctx.check(ctx.L.rb_graph_add_reads_ascii(g.h, _ptr(ascii_bases), None, _ptr(ascii_off), n_a, 0, 0, C.byref(na)))
            ctx.check(ctx.L.rb_graph_count_reads_ascii(g.h, _ptr(ascii_bases), _ptr(ascii_off), n_a, _ptr(a_counts), None, None, C.byref(na)))
        ascii_step()
        t0 = time.perf_counter()
        for _ in range(2):
            ascii_step()
        ta = time.perf_counter() - t0
        assert na.value == n_a * KMERS_PER_READ and float(a_counts[:1024].min()) >= 1.0
        e2e["ascii"] = {"value": n_a * KMERS_PER_READ * 2 / ta, "unit": "k-mers/s", "reads_per_call": n_a, "h2d_bytes_per_step": 2 * (n_a * READ_LEN + 8 * (n_a + 1)),
                        "d2h_bytes_per_step": n_a * KMERS_PER_READ * 4,
                        "note": "rb_graph_add_reads_ascii + rb_graph_count_reads_ascii: ASCII records of equal length in pinned host memory, segmented and packed on the GPU"}

    # ---- roofline (live CUDA-event times of the timed region; algorithmic bytes = SURVEY 8d sector model, 192.15 B per k-mer and phase)
    kmers_total = nk * args.steps
    ins_dominant = t_ins >= t_look
    a_k = A_INSERT if ins_dominant else A_LOOKUP
    kern_ms = {k: round(v[0] / args.steps, 4) for k, v in prof.items()}       # per step
    top = max(prof.items(), key=lambda kv: kv[1][0]) if prof else ("?", (0.0, 1))
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    traffic_tab = json.load(open(tp)) if os.path.exists(tp) else {}
    if args.engine == "direct" or not any(k_.startswith("ks_") for k_ in prof):
        name = "k_graph_insert" if ins_dominant else "k_graph_count"
        ms_sum, calls = prof.get(name, (t_ins if ins_dominant else t_look, args.steps))
        kmers_per_launch = kmers_total / calls
        ms_per_launch = ms_sum / calls
        traffic = traffic_tab.get(name)
    else:
        # a phase of the sliced engine is a chain of kernels over one round; the sector model prices the phase (the k-mer operation),
        # so the roofline line is the whole chain: everything between the phase's first and last kernel, memsets and syncs included
        name = ("insert" if ins_dominant else "lookup") + " round of the sliced engine (kernel chain, see kernels_ms_per_step)"
        first = "ks_route_keys" if ins_dominant else "ks_route_lookup"
        rounds = max([v[1] for k_, v in prof.items() if k_.startswith(first)] + [args.steps])
        kmers_per_launch = kmers_total / rounds
        ms_per_launch = (t_ins if ins_dominant else t_look) / rounds
        traffic = traffic_tab.get("sliced_insert_round" if ins_dominant else "sliced_lookup_round")
    achieved = kmers_per_launch * a_k / (ms_per_launch * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": hbm,
                "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_kmer": a_k, "kmers_per_launch": kmers_per_launch, "ms_per_launch": ms_per_launch,
                "insert_gkmers_s": nk * args.steps / t_ins / 1e6, "lookup_gkmers_s": nk * args.steps / t_look / 1e6,
                "step_frac": value * A_STEP / 1e9 / hbm, "engine": args.engine, "kernels_ms_per_step": kern_ms,
                "top_kernel": top[0], "top_kernel_share_of_step": top[1][0] / t_total_ms}

    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        dbg_bits, cbf_bytes, shift = host_filter_sizes()
        km, ti, tl = cpu_sample(300_000, threads, dbg_bits=dbg_bits, cbf_bytes=cbf_bytes)
        cpu = {"value": km / (ti + tl), "unit": "k-mers/s", "cores": threads, "kind": "port",
               "sample": "300000 reads (37.8 M k-mers) insert then lookup, %d threads, filters %s; C restatement of the Java path, not JVM"
                         % (threads, "full size" if shift == 0 else "scaled 1/%d to fit host RAM" % (1 << shift)),
               "insert_kmers_s": km / ti, "lookup_kmers_s": km / tl}

    line = {"metric": "k-mers/s (insert+lookup) at k=25, 2x150 bp reads", "value": value, "unit": "k-mers/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": dict(config_dict(n_reads, 1), engine=args.engine),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": wall}
    print(json.dumps(line), flush=True)
    g.destroy()
    ctx.close()


if __name__ == "__main__":
    main()
