"""The N>1 host path on CPU: world_size-2 gloo run of the sharded orchestrator (rna-bloom_b200/sharded.py) with the CPU stand-in
backend.  Checks that probes reach the owner of their index, replies come back to the right k-mer, rounds stay in lock-step, and the
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


def _reads():
    from oracle.binding import Oracle
    orc = Oracle()
    reads = [bytes(r).decode() for r in orc.synth_reads(13, 6000, 0, 240, 100, 8000)]
    reads[5] = reads[5][:40] + "N" + reads[5][41:]
    reads[17] = "ACGT"          # shorter than k
    return reads


def _worker(rank, world, port, stranded, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rnabloom_b200  # noqa: F401  (package import must work without a GPU)
        from rnabloom_b200.sharded import ShardedGraph
        from oracle.binding import Oracle
        from shard_cpu_backend import CpuShardBackend
        orc = Oracle()
        reads = _reads()
        mine = reads[rank::world]
        be = CpuShardBackend(orc, world, rank, DBG_BITS, CBF_BYTES, HD, HC, K, stranded, max_kmers=4000)
        sg = ShardedGraph(be, rank, world)
        per_round = 40
        n_rounds = -(-max(len(reads[r::world]) for r in range(world)) // per_round)
        total = 0
        for r in range(n_rounds):
            total += sg.add_round(mine[r * per_round:(r + 1) * per_round], 0)
            sg.check_overflow()
        # second pass over a subset with the other policies
        sg.add_round(mine[:30], 2)          # addCountIfPresent
        sg.add_round(mine[30:60], 4)        # addDbgOnly
        counts = torch.zeros(sum(max(0, len(s) - K + 1) for s in mine[:per_round]), dtype=torch.float32)
        sg.count_round(mine[:per_round], counts)
        dbg = sg.gather_filter(0, (DBG_BITS + 7) // 8)
        cbf = sg.gather_filter(1, CBF_BYTES)
        if rank == 0:
            np.save(os.path.join(out, "dbg.npy"), dbg), np.save(os.path.join(out, "cbf.npy"), cbf)
        np.save(os.path.join(out, "counts%d.npy" % rank), counts.numpy())
        assert sg.exchanged_bytes > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("stranded", [False, True])
def test_two_rank_sharded_protocol_matches_oracle(tmp_path, orc, stranded):
    from oracle.binding import MODE_CANON, MODE_FWD, OracleGraph
    from parity_util import all_bases, assert_cbf_close
    world, port = 2, 29500 + os.getpid() % 2000 + (1 if stranded else 0)
    mp.spawn(_worker, args=(world, port, stranded, str(tmp_path)), nprocs=world, join=True)
    reads = _reads()
    og = OracleGraph(orc, DBG_BITS, CBF_BYTES, 64, HD, HC, 1, K, stranded, False)
    for s in reads:
        og.add_read(s)
    for r in range(world):
        mine = reads[r::world]
        for s in mine[:30]:
            og.add_read(s, flags=2)
        for s in mine[30:60]:
            og.add_read(s, flags=4)
    assert (np.load(tmp_path / "dbg.npy") == og.dbgbf()).all()
    bases = all_bases(orc, reads, K, [MODE_FWD if stranded else MODE_CANON])
    assert_cbf_close(np.load(tmp_path / "cbf.npy"), og.cbf(), bases, K, HC, CBF_BYTES, max_frac=0.05)
    for r in range(world):
        mine = reads[r::world][:40]
        want = np.concatenate([og.count_seq(s)[0] for s in mine if len(s) >= K])
        got = np.load(tmp_path / ("counts%d.npy" % r))
        # counts are read from the gathered state, which may differ from the oracle only on shared counters
        assert len(got) == len(want) and (got == want).mean() > 0.99
