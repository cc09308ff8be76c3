"""The hash-sharded graph (rb_mgraph_*, csrc/rb_mgraph_host.inl; host mirror rna-bloom_b200/sharded.py ShardedGraph) at world sizes 2, 4
and 8 on CPU: the library's own round orchestrator with a gloo transport plugged in, between processes that each run the *real* kernel
sources through the host emulation of tests/emu (test infrastructure, see tests/test_emu_parity.py).  Checks that probes reach the owner of their filter slice, answers come
back to the right k-mer, duplicates of a k-mer that live on different ranks are aggregated at the key's home rank, and the
concatenated shares equal the sequential oracle's single arrays."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

K, HD, HC = 25, 3, 2
DBG_BITS, CBF_BYTES = 3_000_017, 1_000_003
SLICE_ENV = {"RB_SLICE_BITS_LOG2": "17", "RB_SLICE_BYTES_LOG2": "15", "RB_SLICED_CELLS": "1", "RB_SLICED_STAGE": "1", "RB_SLICED_SUBRANGE_LOG2": "6",
             "RB_SLICED_CHUNK": "8192"}


def _reads(seed=13):
    from oracle.binding import Oracle
    orc = Oracle()
    reads = [bytes(r).decode() for r in orc.synth_reads(seed, 9000, 0, 240, 100, 8000)]   # ~2.7x coverage: counters stay exact
    reads[5] = reads[5][:40] + "N" + reads[5][41:]
    reads[17] = "ACGT"          # shorter than k
    return reads


def _worker(rank, world, port, stranded, out, cfg=None):
    cfg = cfg or {}
    K, HD, HC = cfg.get("k", globals()["K"]), cfg.get("hd", globals()["HD"]), cfg.get("hc", globals()["HC"])
    DBG_BITS, CBF_BYTES = cfg.get("dbg_bits", globals()["DBG_BITS"]), cfg.get("cbf_bytes", globals()["CBF_BYTES"])
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **dict(SLICE_ENV, **cfg.get("env", {})))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rnabloom_b200 as rb
        from rnabloom_b200 import binding as B
        from rnabloom_b200.sharded import GlooTransport, ShardedGraph
        from test_emu_parity import EMU_SO
        B._lib = B.bind(EMU_SO, allow_missing=True)   # the emulated kernels: "device memory" is host memory
        reads = _reads(cfg.get("seed", 13))
        mine = reads[rank::world]
        ctx = rb.Context(0)
        sg = ShardedGraph(ctx, world, rank, DBG_BITS, CBF_BYTES, HD, HC, K, stranded, 8000, transport=GlooTransport(), device=torch.device("cpu"))
        per_round = 40
        n_rounds = -(-max(len(reads[r::world]) for r in range(world)) // per_round)
        total = 0
        keep = []
        for r in range(n_rounds):
            pr = rb.pack_reads(mine[r * per_round:(r + 1) * per_round])
            keep.append(pr)
            total += sg.add_round(pr.args(), 0)
        sg.check_overflow()
        for chunk, flags in ((mine[:30], 2), (mine[30:60], 4)):   # addCountIfPresent, addDbgOnly
            pr = rb.pack_reads(chunk)
            keep.append(pr)
            sg.add_round(pr.args(), flags)
        q = rb.pack_reads(mine[:per_round])
        n_inst = sum(max(0, len(s) - K + 1) for s in mine[:per_round])
        counts = torch.zeros(n_inst + 8, dtype=torch.float32)
        fh = torch.zeros(n_inst + 8, dtype=torch.int64)
        assert sg.count_round(q.args(), counts, fh) == n_inst
        sg.check_overflow()
        dbg = sg.gather_filter(0, (DBG_BITS + 7) // 8)
        cbf = sg.gather_filter(1, CBF_BYTES)
        if rank == 0:
            np.save(os.path.join(out, "dbg.npy"), dbg), np.save(os.path.join(out, "cbf.npy"), cbf)
        np.save(os.path.join(out, "counts%d.npy" % rank), counts.numpy()[:n_inst])
        np.save(os.path.join(out, "fh%d.npy" % rank), fh.numpy()[:n_inst])
        assert sg.exchanged_bytes > 0
        sg.close(), ctx.close()
    finally:
        dist.destroy_process_group()


# paired: cbf_bytes = 2^c dividing dbg_bits and h_d >= h_c -> one probe record per hash, a rank owns paired slices (its bits are a
# range inside every chunk of cbf_bytes bits: gather_filter reassembles them)
PAIRED = {"dbg_bits": 5 << 22, "cbf_bytes": 1 << 22, "hd": 3, "hc": 3, "env": {"RB_SLICE_PAIR_LOG2": "13", "RB_SLICE_REGION_TARGET": "64"}}


@pytest.mark.parametrize("world,stranded,paired", [(2, False, False), (2, True, False), (4, False, False), (8, False, False), (2, False, True),
                                                   (4, True, True)])
def test_sharded_sliced_graph_matches_oracle(tmp_path, orc, world, stranded, paired):
    from oracle.binding import MODE_CANON, MODE_FWD, OracleGraph
    from parity_util import all_bases, assert_cbf_close
    from test_emu_parity import build_emu
    build_emu()
    port = 31500 + os.getpid() % 2000 + world * 3 + (1 if stranded else 0) + (40 if paired else 0)
    cfg = PAIRED if paired else None
    mp.spawn(_worker, args=(world, port, stranded, str(tmp_path), cfg), nprocs=world, join=True)
    reads = _reads()
    DBG_BITS, CBF_BYTES, HD, HC = ((cfg["dbg_bits"], cfg["cbf_bytes"], cfg["hd"], cfg["hc"]) if paired else
                                   (globals()["DBG_BITS"], globals()["CBF_BYTES"], globals()["HD"], globals()["HC"]))
    og = OracleGraph(orc, DBG_BITS, CBF_BYTES, 64, HD, HC, 1, K, stranded, False)
    for s in reads:
        og.add_read(s)
    for r in range(world):
        mine = reads[r::world]
        for s in mine[:30]:
            og.add_read(s, flags=2)
        for s in mine[30:60]:
            og.add_read(s, flags=4)
    assert og.cbf().max() <= 16, "fixture reached the probabilistic MiniFloat range"
    assert (np.load(tmp_path / "dbg.npy") == og.dbgbf()).all()
    bases = all_bases(orc, reads, K, [MODE_FWD if stranded else MODE_CANON])
    assert_cbf_close(np.load(tmp_path / "cbf.npy"), og.cbf(), bases, K, HC, CBF_BYTES, max_frac=0.05)
    for r in range(world):
        mine = reads[r::world][:40]
        want = np.concatenate([og.count_seq(s)[0] for s in mine if len(s) >= K])
        wantf = np.concatenate([og.count_seq(s)[1] for s in mine if len(s) >= K])
        got = np.load(tmp_path / ("counts%d.npy" % r))
        assert (np.load(tmp_path / ("fh%d.npy" % r)) == wantf).all()
        # counts are read from the gathered state, which may differ from the oracle only on shared counters
        assert len(got) == len(want) and (got == want).mean() > 0.99


@pytest.mark.parametrize("seed", range(int(os.environ.get("RB_FUZZ_SEEDS", "2"))))
def test_sharded_sliced_graph_random_configurations(tmp_path, orc, seed):
    """Seeded random world size, k, hash counts, filter sizes and slice geometry through the same protocol."""
    from oracle.binding import MODE_CANON, MODE_FWD, OracleGraph
    from parity_util import all_bases, assert_cbf_close
    from test_emu_parity import build_emu
    build_emu()
    rng = np.random.default_rng(9000 + seed)
    world = int(rng.choice([2, 4, 8]))
    stranded = bool(rng.integers(0, 2))
    cfg = {"k": int(rng.integers(15, 61)), "hd": int(rng.integers(1, 4)), "hc": int(rng.integers(1, 4)), "seed": 100 + seed,
           "dbg_bits": int(rng.integers(1 << 22, 1 << 25)) | 1, "cbf_bytes": int(rng.integers(1 << 19, 1 << 22)) | 1,
           "env": {"RB_SLICE_BITS_LOG2": str(int(rng.integers(15, 21))), "RB_SLICE_BYTES_LOG2": str(int(rng.integers(13, 19))),
                   "RB_SLICED_CELLS": str(int(rng.integers(0, 2))), "RB_SLICED_SUBRANGE_LOG2": str(int(rng.integers(4, 8)))}}
    port = 33500 + os.getpid() % 2000 + seed
    mp.spawn(_worker, args=(world, port, stranded, str(tmp_path), cfg), nprocs=world, join=True)
    k, hd, hc, dbg_bits, cbf_bytes = cfg["k"], cfg["hd"], cfg["hc"], cfg["dbg_bits"], cfg["cbf_bytes"]
    reads = _reads(cfg["seed"])
    og = OracleGraph(orc, dbg_bits, cbf_bytes, 64, hd, hc, 1, k, stranded, False)
    for s in reads:
        og.add_read(s)
    for r in range(world):
        mine = reads[r::world]
        for s in mine[:30]:
            og.add_read(s, flags=2)
        for s in mine[30:60]:
            og.add_read(s, flags=4)
    assert (np.load(tmp_path / "dbg.npy") == og.dbgbf()).all()
    if og.cbf().max() <= 15:
        # allowed differences = the order dependence the reference has itself: k-mers that share a counter, and k-mers that share a
        # dbgbf bit with another k-mer (whether such a k-mer counts as present at its first sighting depends on which came first)
        from parity_util import counters_that_may_differ, np_slots
        bases = all_bases(orc, reads, k, [MODE_FWD if stranded else MODE_CANON])
        allowed, _ = counters_that_may_differ(bases, k, hc, cbf_bytes)
        bits = np_slots(bases, k, hd, dbg_bits)
        uniq, cnt = np.unique(bits.reshape(-1), return_counts=True)
        fp_prone = np.isin(bits, uniq[cnt > 1]).any(axis=1)
        allowed |= set(int(x) for x in np_slots(bases[fp_prone], k, hc, cbf_bytes).reshape(-1))
        diff = np.nonzero(np.load(tmp_path / "cbf.npy") != og.cbf())[0]
        assert set(diff.tolist()) <= allowed, "cbf differs on counters of k-mers that share neither a counter nor a dbgbf bit"
    for r in range(world):
        mine = reads[r::world][:40]
        wantf = np.concatenate([og.count_seq(s)[1] for s in mine if len(s) >= k])
        assert (np.load(tmp_path / ("fh%d.npy" % r)) == wantf).all()


def _parity_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **SLICE_ENV)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rnabloom_b200 as rb
        from rnabloom_b200 import binding as B
        from rnabloom_b200.sharded import GlooTransport, ShardedGraph
        from test_emu_parity import EMU_SO
        import bench_multi
        B._lib = B.bind(EMU_SO, allow_missing=True)
        ctx = rb.Context(0)
        cpu = torch.device("cpu")

        def make_graph(db, cb, max_kmers):
            return ShardedGraph(ctx, world, rank, db, cb, 3, 3, 25, False, max_kmers, transport=GlooTransport(), device=cpu)
        full_d, full_c = 1 << 27, 1 << 24
        full = make_graph(full_d, full_c, 40000)
        res = bench_multi.sharded_parity_check(make_graph, rank, world, cpu, full_d, full_c, full, small=(3_000_017, 1_000_003, 120, 40), n_full=120)
        if rank == 0:
            import json
            with open(os.path.join(out, "parity.json"), "w") as fh:
                json.dump(res, fh)
        full.close(), ctx.close()
    finally:
        dist.destroy_process_group()


def test_bench_parity_check_runs_at_world_2(tmp_path):
    """bench_multi.sharded_parity_check -- the oracle comparison `bench.py --gpus N` runs after its timed region -- on 2 gloo ranks over
    the emulated kernels: the checker itself is exercised without a GPU (fixtures scaled down, same code path)."""
    import json
    from test_emu_parity import build_emu
    build_emu()
    port = 35500 + os.getpid() % 2000
    mp.spawn(_parity_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    res = json.load(open(tmp_path / "parity.json"))
    assert res["dbgbf"] == "equal" and res["counts_equal_frac"] > 0.999 and res["full_geometry"]["counts_equal_frac"] > 0.9999
