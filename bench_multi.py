"""bench.py's N>1 arm: the hash-sharded pipeline (rna-bloom_b200/sharded.py) under torchrun, one rank per GPU over NCCL.

Weak scaling: every rank feeds `reads_per_step` fresh synthetic reads per step; the logical filters grow with the GPU count so that each
rank always owns 8 GiB of dbgbf + 8 GiB of cbf (N=8: 64 GiB + 64 GiB, BASELINE.json configs[2]); the virtual genome grows alike (3 Gb per GPU).
value = k-mers inserted and looked up by all ranks / max-over-ranks device time."""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

import bench as single


class _DevReads:
    """ASCII reads -> packed -> device tensors; .args is the raw-pointer tuple the sharded calls take."""

    def __init__(self, rb, seqs, device):
        import ctypes as C
        pr = rb.pack_reads(seqs)
        self.keep = []

        def up(a):
            if a is None:
                return None
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy()).to(device)
            pad = torch.zeros(64, dtype=torch.uint8, device=device)
            t = torch.cat([t, pad])
            self.keep.append(t)
            return C.c_void_p(t.data_ptr())
        self.args = (up(pr.packed), up(pr.mask), up(pr.read_off), up(pr.read_len), pr.n_reads, pr.uniform_len, pr.uniform_stride)


def sharded_parity_check(make_graph, rank, world, device, full_dbg_bits=None, full_cbf_bytes=None, full_graph=None,
                         small=((1 << 30) + 77, (1 << 28) + 13, 2000, 500), n_full=20000):
    """Correctness of the N-rank data path against the CPU oracle, on the very box and process group the numbers come from
    (graph/BloomFilterDeBruijnGraph.java:405-412 add, :562-570 getCount).  The oracle is the checker here, never the thing timed.

    A  small odd-sized filters (no power of two, shares of unequal length): every rank inserts its own reads in several rounds, then
       re-inserts some (multiplicities inside a round, k-mers present before the round), then addCountIfPresent / addDbgOnly; all
       shares are gathered and compared with the sequential oracle: dbgbf byte-equal, cbf equal except on counters that two distinct
       k-mers share, getKmers counts of every rank's query equal to the oracle's.
    B  the production shard geometry (`full_graph`, the graph the bench just timed, emptied first): a fixture sparse enough that no
       two k-mers share a bit or counter, so the expected state follows from the k-mer multiplicities alone: popcount of all dbgbf
       shares == distinct probe indices, non-zero counters == distinct counter indices of repeated k-mers, count of every queried
       k-mer == its multiplicity.
    Returns the dict that goes into the bench line; raises AssertionError on any difference."""
    import rnabloom_b200 as rb
    sys_path_tests = os.path.join(single.ROOT, "tests")
    import sys
    if sys_path_tests not in sys.path:
        sys.path.insert(0, sys_path_tests)
    from oracle.binding import MODE_CANON, Oracle, OracleGraph
    from parity_util import all_bases, assert_cbf_close, np_slots
    orc = Oracle()
    k, hd, hc = single.K, single.HD, single.HC
    out = {"n_ranks": world}
    alive = []   # device buffers of every round stay allocated until the check is over (the calls only take raw pointers)

    def sync():
        if device.type == "cuda":   # torch allocates / fills on its own stream, the library works on the context's stream
            torch.cuda.synchronize()

    def dev(seqs):
        alive.append(_DevReads(rb, seqs, device))
        sync()
        return alive[-1]
    # ---- A ---------------------------------------------------------------------------------------------------------------------
    dbg_bits, cbf_bytes, per_rank, per_round = small
    q0, q1 = per_rank // 20, per_rank // 20 + per_rank // 10           # the queried / re-inserted reads
    # ~1.3x coverage: counters stay in the exact MiniFloat range at any world size
    reads = [bytes(r).decode() for r in orc.synth_reads(71, 120 * per_rank * world, 0, per_rank * world, 150, 6000)]
    reads[3] = reads[3][:70] + "N" + reads[3][71:]
    mine = reads[rank::world]
    sg = make_graph(dbg_bits, cbf_bytes, per_round * 126)
    for r in range(0, per_rank, per_round):
        sg.add_round(dev(mine[r:r + per_round]).args, 0)
    again = mine[:q1] + mine[:q1 // 2]
    for r in range(0, len(again), per_round):
        sg.add_round(dev(again[r:r + per_round]).args, 0)
    dq = dev(mine[q0:q1])
    sg.add_round(dq.args, rb.ADD_COUNT_IF_PRESENT)
    sg.add_round(dq.args, rb.DBG_ONLY)
    sg.check_overflow()
    n_inst = sum(max(0, len(s) - k + 1) for s in mine[q0:q1])
    counts = torch.zeros(n_inst, dtype=torch.float32, device=device)
    fh = torch.zeros(n_inst, dtype=torch.int64, device=device)
    sync()
    assert sg.count_round(dq.args, counts, fh) == n_inst
    sg.check_overflow()
    sync()
    dbg = sg.gather_filter(rb.RB_DBGBF, (dbg_bits + 7) // 8)
    cbf = sg.gather_filter(rb.RB_CBF, cbf_bytes)
    og = OracleGraph(orc, dbg_bits, cbf_bytes, 64, hd, hc, 1, k, False, False)
    for s in reads:
        og.add_read(s)
    for r in range(world):
        m = reads[r::world]
        for s in m[:q1] + m[:q1 // 2]:
            og.add_read(s)
    for r in range(world):
        for s in reads[r::world][q0:q1]:
            og.add_read(s, flags=2)
    assert og.cbf().max() <= 16, "parity fixture reached the probabilistic MiniFloat range"
    assert (dbg == og.dbgbf()).all(), "rank %d: gathered dbgbf differs from the oracle" % rank
    assert_cbf_close(cbf, og.cbf(), all_bases(orc, reads, k, [MODE_CANON]), k, hc, cbf_bytes)
    want = np.concatenate([og.count_seq(s)[0] for s in mine[q0:q1]])
    wantf = np.concatenate([og.count_seq(s)[1] for s in mine[q0:q1]])
    got = counts.cpu().numpy()
    assert (fh.cpu().numpy() == wantf).all(), "rank %d: forward hashes differ" % rank
    frac = float((got == want).mean())
    assert frac > 0.999, "rank %d: getKmers counts differ from the oracle (%.5f equal)" % (rank, frac)
    out.update({"dbgbf": "equal", "cbf": "equal up to shared counters (%d of %d bytes differ)" % (int((cbf != og.cbf()).sum()), cbf_bytes),
                "counts_equal_frac": frac, "small": {"dbgbf_bits": dbg_bits, "cbf_bytes": cbf_bytes, "reads_per_rank": per_rank + len(again) + 2 * (q1 - q0)}})
    og.close()
    sg.close()
    # ---- B ---------------------------------------------------------------------------------------------------------------------
    if full_graph is not None:
        fg = full_graph
        fg.clear()
        n_b = n_full
        n_r, n_qr = n_b // 4, (3 * n_b) // 10
        reads = [bytes(r).decode() for r in orc.synth_reads(73, 50 * n_b * world, 0, n_b * world, 150, 4000)]
        mine = reads[rank::world]
        dr = dev(mine)
        fg.add_round(dr.args, 0)
        dr2 = dev(mine[:n_r])
        fg.add_round(dr2.args, 0)
        fg.check_overflow()
        n_q = sum(max(0, len(s) - k + 1) for s in mine[:n_qr])
        counts = torch.zeros(n_q, dtype=torch.float32, device=device)
        dq = dev(mine[:n_qr])
        sync()
        assert fg.count_round(dq.args, counts) == n_q
        fg.check_overflow()
        pops = torch.tensor([fg.popcount(rb.RB_DBGBF), fg.popcount(rb.RB_CBF)], dtype=torch.int64, device=device)
        dist.all_reduce(pops)
        bases = np.concatenate([orc.kmer_hashes(s, k, MODE_CANON)[2] for s in reads] +
                               [orc.kmer_hashes(s, k, MODE_CANON)[2] for r in range(world) for s in reads[r::world][:n_r]])
        keys, mult = np.unique(bases, return_counts=True)
        assert mult.max() <= 17
        sd = np_slots(keys, k, hd, full_dbg_bits)
        sc = np_slots(keys, k, hc, full_cbf_bytes)
        # sparse fixture: k-mers that share a bit or a counter with another k-mer are vanishingly few; they only loosen the bounds
        n_bits = len(np.unique(sd.reshape(-1)))
        rep = sc[mult >= 2].reshape(-1)
        n_cnt = len(np.unique(rep))
        shared_c = sc.size - len(np.unique(sc.reshape(-1)))
        assert int(pops[0]) == n_bits, "dbgbf popcount over all shares %d != distinct probe indices %d" % (int(pops[0]), n_bits)
        assert abs(int(pops[1]) - n_cnt) <= shared_c, "cbf non-zero counters %d != %d" % (int(pops[1]), n_cnt)
        mq = np.concatenate([orc.kmer_hashes(s, k, MODE_CANON)[2] for s in mine[:n_qr]])
        want = mult[np.searchsorted(keys, mq)].astype(np.float32)
        got = counts.cpu().numpy()
        frac_b = float((got == want).mean())
        assert frac_b > 0.9999, "full geometry: counts differ from the k-mer multiplicities (%.6f equal)" % frac_b
        out["full_geometry"] = {"dbgbf_bits": full_dbg_bits, "cbf_bytes": full_cbf_bytes, "reads": (n_b + n_r) * world,
                                "dbgbf_popcount": int(pops[0]), "expected": n_bits, "cbf_nonzero": int(pops[1]), "expected_cbf": n_cnt,
                                "counts_equal_frac": frac_b}
    return out


def run_sharded(args, rank, world, local_rank):
    import rnabloom_b200 as rb
    from rnabloom_b200.sharded import ShardedGraph, broadcast_nccl_id
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # its banner goes to stdout, where the bench contract wants one JSON line
        os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    args.warmup = max(args.warmup, 3)
    K, L, STRIDE = single.K, single.READ_LEN, single.STRIDE
    kpr = single.KMERS_PER_READ
    dbg_bits, cbf_bytes = single.DBG_BITS * world, single.CBF_BYTES * world
    genome = args.genome * world
    # rounds of up to 504 M k-mers per rank, as on one GPU (one sweep of the local 16 GiB per phase is amortised over the round)
    reads_per_round = min(args.reads_per_step, int(os.environ.get("RB_BENCH_READS_PER_ROUND", "4000000")))
    rounds = max(1, args.reads_per_step // reads_per_round)
    n_reads = rounds * reads_per_round
    ctx = rb.Context(local_rank)
    dev = torch.device("cuda", local_rank)
    # the library owns the exchange: its own NCCL communicator on its own stream; torch.distributed only carries the 128-byte id,
    # the barriers and the max-over-ranks of the timings
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    sg = ShardedGraph(ctx, world, rank, dbg_bits, cbf_bytes, single.HD, single.HC, K, False, reads_per_round * kpr, nccl_id=broadcast_nccl_id(dev))
    total_steps = args.warmup + args.steps
    n_batches = min(total_steps, max(1, 100_000_000 // n_reads))
    words = n_reads * STRIDE // 32
    batches = []
    for s in range(n_batches):
        t = torch.empty(words + 8, dtype=torch.int64, device="cuda")
        first = (s * world + rank) * n_reads                   # disjoint read ids per rank and step
        ctx.synth_reads_dev(single.SEED, genome, first, n_reads, L, single.ERR_PPM, STRIDE, t.data_ptr())
        batches.append(t)
    counts = torch.empty(n_reads * kpr, dtype=torch.float32, device="cuda")   # one step's counts; round r writes its own slice
    import ctypes as C

    def reads_of(t, r):
        base = t.data_ptr() + r * reads_per_round * (STRIDE // 4)
        return (C.c_void_p(base), None, None, None, reads_per_round, L, STRIDE)

    def step(t):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream)
        for r in range(rounds):
            sg.add_round(reads_of(t, r), 0)
        e[1].record(stream)
        for r in range(rounds):
            sg.count_round(reads_of(t, r), counts[r * reads_per_round * kpr:])
        e[2].record(stream)
        return e

    for s in range(args.warmup):
        step(batches[s % n_batches])
    sg.check_overflow()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = single.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.kernel_launches()
    x0 = sg.exchanged_bytes
    evs = []
    ctx.profile_enable(True)   # CUDA-event spans around every kernel and every exchange of the timed region (library side)
    ctx.profile_read()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(args.warmup, total_steps):
        evs.append(step(batches[s % n_batches]))
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    t_ins = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_look = sum(e[1].elapsed_time(e[2]) for e in evs)
    tt = torch.tensor([t_ins + t_look, t_ins, t_look, wall * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_ins, t_look, wall_ms = [float(x) for x in tt.cpu()]
    nk_rank = n_reads * kpr
    value = nk_rank * world * args.steps / (t_total * 1e-3)
    launches = ctx.kernel_launches() - l0
    xbytes = (sg.exchanged_bytes - x0) / args.steps
    # per-phase device time of THIS rank (rank 0 reports its own; an exchange span includes waiting for the slowest peer)
    kern_ms = {k_: round(v[0] / args.steps, 4) for k_, v in prof.items() if k_ != "exchange"}
    exch_ms = round(prof.get("exchange", (0.0, 0))[0] / args.steps, 4)

    # end to end at N GPUs: packed reads start in pinned host memory, counts end there.  The D2H of a step's counts runs on a copy
    # stream out of one of two count buffers, so it overlaps the next step's insert rounds (as the single-GPU host-pointer calls do)
    e2e = None
    if not args.no_e2e:
        e_steps = 5
        h_in = torch.empty(words + 8, dtype=torch.int64).pin_memory()
        h_in.copy_(batches[0].cpu())
        h_out = torch.empty(n_reads * kpr, dtype=torch.float32).pin_memory()
        d_in = torch.empty_like(batches[0])
        cbuf = [counts, torch.empty_like(counts)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [None, None]
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(e_steps):
            c = cbuf[i % 2]
            with torch.cuda.stream(stream):
                d_in.copy_(h_in, non_blocking=True)
            for r in range(rounds):
                sg.add_round(reads_of(d_in, r), 0)
            if copied[i % 2] is not None:
                stream.wait_event(copied[i % 2])      # the D2H that last read this count buffer is done
            for r in range(rounds):
                sg.count_round(reads_of(d_in, r), c[r * reads_per_round * kpr:])
            done = torch.cuda.Event()
            done.record(stream)
            copy_stream.wait_event(done)
            with torch.cuda.stream(copy_stream):
                piece = 8 << 20   # floats: 32 MiB copies, so that the next step's reads (H2D) are not held behind one 2 GB D2H on a shared copy engine
                for o in range(0, h_out.numel(), piece):
                    h_out[o:o + piece].copy_(c[o:o + piece], non_blocking=True)
                copied[i % 2] = torch.cuda.Event()
                copied[i % 2].record(copy_stream)
        torch.cuda.synchronize()
        dist.barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": nk_rank * world * e_steps / float(te.item()), "unit": "k-mers/s", "h2d_bytes_per_step": words * 8 * world,
               "d2h_bytes_per_step": nk_rank * 4 * world, "steps": e_steps}
        del cbuf, h_out, h_in

    # ---- correctness of the path that was just timed: oracle comparison over NCCL (all ranks; a difference fails the run) -----------------
    parity = None
    if not getattr(args, "no_parity", False):
        def make_graph(db, cb, max_kmers):
            return ShardedGraph(ctx, world, rank, db, cb, single.HD, single.HC, K, False, max_kmers, nccl_id=broadcast_nccl_id(dev))
        try:
            parity = sharded_parity_check(make_graph, rank, world, torch.device("cuda", local_rank), dbg_bits, cbf_bytes, sg)
            ok = 1
        except AssertionError as e:
            parity = {"n_ranks": world, "error": str(e)}
            ok = 0
        okt = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if not int(okt.item()) and ok:
            parity = {"n_ranks": world, "error": "another rank reported a difference"}
        parity_ok = bool(int(okt.item()))
    else:
        parity_ok = True

    if rank == 0:
        clocks = sampler.stop()
        hbm, peak_src = single.peaks()
        cfg = single.config_dict(n_reads, world)
        cfg.update({"dbgbf_bits": dbg_bits, "cbf_bytes": cbf_bytes, "genome_len": genome,
                    "workload": "BASELINE.json configs[2] shape, weak-scaled: %d x %d reads/step, k=25, dbgbf %d GiB + cbf %d GiB sharded by index range over %d GPUs"
                                % (world, n_reads, dbg_bits >> 33, cbf_bytes >> 30, world),
                    "exchange": ("peer-to-peer: consumer kernels read the producers' arenas over NVLink (CUDA IPC), NCCL only for barriers"
                                 if sg.peer_to_peer else "library-owned NCCL all-to-all (rb_mgraph_*), %.0f MB per rank per step" % (xbytes / 1e6)),
                    "paired_probe_records": bool(sg.paired), "kmers_per_round_per_gpu": reads_per_round * kpr})
        line = {"metric": "k-mers/s (insert+lookup) at k=25, 2x150 bp reads", "value": value, "unit": "k-mers/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches) * world,
                "roofline": {"bound": "hbm", "kernel": "sharded pipeline (whole step, per GPU)", "achieved": value / world * single.A_STEP / 1e9,
                             "peak": hbm, "unit": "GB/s", "frac": value / world * single.A_STEP / 1e9 / hbm, "traffic": None,
                             "peak_source": peak_src, "insert_gkmers_s": nk_rank * world * args.steps / t_ins / 1e6,
                             "lookup_gkmers_s": nk_rank * world * args.steps / t_look / 1e6,
                             "kernels_ms_per_step": kern_ms, "exchange_ms_per_step": exch_ms,
                             "exchange_gb_per_rank_per_step": xbytes / 1e9},
                "cpu_baseline": None, "wall_s_timed_region": wall_ms / 1e3, "parity_check": parity}
        print(json.dumps(line), flush=True)
    elif parity is not None and "error" in parity:
        import sys
        print("rank %d parity: %s" % (rank, parity["error"]), file=sys.stderr, flush=True)
    sg.close()
    ctx.close()
    dist.destroy_process_group()
    if not parity_ok:
        raise SystemExit(3)
