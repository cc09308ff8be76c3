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


def run_sharded(args, rank, world, local_rank):
    import rnabloom_b200 as rb
    from rnabloom_b200.sharded import GpuBackend, ShardedGraph, SlicedBackend, SlicedShardedGraph
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # its banner goes to stdout, where the bench contract wants one JSON line
        os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    args.warmup = max(args.warmup, 3)
    K, L, STRIDE = single.K, single.READ_LEN, single.STRIDE
    kpr = single.KMERS_PER_READ
    dbg_bits, cbf_bytes = single.DBG_BITS * world, single.CBF_BYTES * world
    genome = args.genome * world
    sliced = getattr(args, "sharded_engine", "sliced") == "sliced"
    # sliced: rounds of 252 M k-mers per rank (the tile sort of the sliced engine routes; equal-split all-to-all of whole regions);
    # legacy: the first-generation pipeline (per-record cursor scatter), 63 M k-mers per round
    reads_per_round = min(args.reads_per_step, 2_000_000 if sliced else 500_000)
    rounds = max(1, args.reads_per_step // reads_per_round)
    n_reads = rounds * reads_per_round
    ctx = rb.Context(local_rank)
    if sliced:
        be = SlicedBackend(ctx, world, rank, dbg_bits, cbf_bytes, single.HD, single.HC, K, False, reads_per_round * kpr)
        sg = SlicedShardedGraph(be, rank, world)
    else:
        be = GpuBackend(ctx, world, rank, dbg_bits, cbf_bytes, single.HD, single.HC, K, False, reads_per_round * kpr)
        sg = ShardedGraph(be, rank, world)
    total_steps = args.warmup + args.steps
    n_batches = min(total_steps, max(1, 100_000_000 // n_reads))
    words = n_reads * STRIDE // 32
    batches = []
    for s in range(n_batches):
        t = torch.empty(words + 8, dtype=torch.int64, device="cuda")
        first = (s * world + rank) * n_reads                   # disjoint read ids per rank and step
        ctx.synth_reads_dev(single.SEED, genome, first, n_reads, L, single.ERR_PPM, STRIDE, t.data_ptr())
        batches.append(t)
    counts = torch.empty(reads_per_round * kpr, dtype=torch.float32, device="cuda")
    import ctypes as C

    def reads_of(t, r):
        base = t.data_ptr() + r * reads_per_round * (STRIDE // 4)
        return (C.c_void_p(base), None, None, None, reads_per_round, L, STRIDE)

    stream = be.stream

    def step(t):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream)
        for r in range(rounds):
            sg.add_round(reads_of(t, r), 0)
        e[1].record(stream)
        for r in range(rounds):
            sg.count_round(reads_of(t, r), counts)
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
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(args.warmup, total_steps):
        evs.append(step(batches[s % n_batches]))
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    sg.check_overflow()
    t_ins = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_look = sum(e[1].elapsed_time(e[2]) for e in evs)
    tt = torch.tensor([t_ins + t_look, t_ins, t_look, wall * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total, t_ins, t_look, wall_ms = [float(x) for x in tt.cpu()]
    nk_rank = n_reads * kpr
    value = nk_rank * world * args.steps / (t_total * 1e-3)
    launches = ctx.kernel_launches() - l0
    xbytes = (sg.exchanged_bytes - x0) / args.steps

    # end to end at N GPUs: packed reads start in pinned host memory, counts end there
    e2e = None
    if not args.no_e2e:
        e_steps = 2
        h_in = torch.empty(words + 8, dtype=torch.int64).pin_memory()
        h_in.copy_(batches[0].cpu())
        h_out = torch.empty(reads_per_round * kpr, dtype=torch.float32).pin_memory()
        d_in = torch.empty_like(batches[0])
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            with torch.cuda.stream(stream):
                d_in.copy_(h_in, non_blocking=True)
            for r in range(rounds):
                sg.add_round(reads_of(d_in, r), 0)
            for r in range(rounds):
                sg.count_round(reads_of(d_in, r), counts)
                with torch.cuda.stream(stream):
                    h_out.copy_(counts, non_blocking=True)
            torch.cuda.synchronize()
        dist.barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": nk_rank * world * e_steps / float(te.item()), "unit": "k-mers/s", "h2d_bytes_per_step": words * 8 * world,
               "d2h_bytes_per_step": nk_rank * 4 * world, "steps": e_steps}

    if rank == 0:
        clocks = sampler.stop()
        hbm, peak_src = single.peaks()
        cfg = single.config_dict(n_reads, world)
        cfg.update({"dbgbf_bits": dbg_bits, "cbf_bytes": cbf_bytes, "genome_len": genome,
                    "workload": "BASELINE.json configs[2] shape, weak-scaled: %d x %d reads/step, k=25, dbgbf %d GiB + cbf %d GiB sharded by index range over %d GPUs"
                                % (world, n_reads, dbg_bits >> 33, cbf_bytes >> 30, world),
                    "exchange": "NCCL all-to-all (torch.distributed), %.0f MB per rank per step" % (xbytes / 1e6),
                    "sharded_engine": "sliced" if sliced else "legacy", "kmers_per_round_per_gpu": reads_per_round * kpr})
        line = {"metric": "k-mers/s (insert+lookup) at k=25, 2x150 bp reads", "value": value, "unit": "k-mers/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches) * world,
                "roofline": {"bound": "hbm", "kernel": "sharded pipeline (whole step, per GPU)", "achieved": value / world * single.A_STEP / 1e9,
                             "peak": hbm, "unit": "GB/s", "frac": value / world * single.A_STEP / 1e9 / hbm, "traffic": None,
                             "peak_source": peak_src, "insert_gkmers_s": nk_rank * world * args.steps / t_ins / 1e6,
                             "lookup_gkmers_s": nk_rank * world * args.steps / t_look / 1e6},
                "cpu_baseline": None, "wall_s_timed_region": wall_ms / 1e3}
        print(json.dumps(line), flush=True)
    be.close()
    ctx.close()
    dist.destroy_process_group()
