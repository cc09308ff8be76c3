"""Parity of the sharded pipeline over NCCL (run under torchrun on >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_nccl.py
Every rank inserts its own reads; the concatenated shares must equal the sequential oracle's arrays."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rnabloom_b200 as rb  # noqa: E402
from oracle.binding import MODE_CANON, Oracle, OracleGraph  # noqa: E402
from parity_util import all_bases, assert_cbf_close  # noqa: E402
from rnabloom_b200.sharded import GpuBackend, ShardedGraph  # noqa: E402
from test_gpu_parity import DevReads  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
orc = Oracle()
k, hd, hc, dbg_bits, cbf_bytes = 25, 3, 3, (1 << 30) + 77, (1 << 28) + 13
reads = [bytes(r).decode() for r in orc.synth_reads(71, 150000, 0, 2000, 150, 6000)]
mine = reads[rank::world]
ctx = rb.Context(local)
be = GpuBackend(ctx, world, rank, dbg_bits, cbf_bytes, hd, hc, k, False, 80000)
sg = ShardedGraph(be, rank, world)
for r in range(0, len(mine), 500):
    dr = DevReads(ctx, rb.pack_reads(mine[r:r + 500]))
    sg.add_round(dr.args, 0)
    sg.check_overflow()
    torch.cuda.synchronize()
    dr.free()
dr = DevReads(ctx, rb.pack_reads(mine[:300]))
sg.add_round(dr.args, 0)
n_inst = sum(max(0, len(s) - k + 1) for s in mine[:300])
counts = torch.zeros(n_inst, dtype=torch.float32, device="cuda")
sg.count_round(dr.args, counts)
torch.cuda.synchronize()
dbg = sg.gather_filter(rb.RB_DBGBF, (dbg_bits + 7) // 8)
cbf = sg.gather_filter(rb.RB_CBF, cbf_bytes)
og = OracleGraph(orc, dbg_bits, cbf_bytes, 64, hd, hc, 1, k, False, False)
for s in reads:
    og.add_read(s)
for r in range(world):
    for s in reads[r::world][:300]:
        og.add_read(s)
assert (dbg == og.dbgbf()).all(), "dbgbf differs"
assert_cbf_close(cbf, og.cbf(), all_bases(orc, reads, k, [MODE_CANON]), k, hc, cbf_bytes)
want = np.concatenate([og.count_seq(s)[0] for s in mine[:300]])
assert (counts.cpu().numpy() == want).mean() > 0.999
print("rank %d/%d: sharded NCCL parity ok (%d MB exchanged)" % (rank, world, sg.exchanged_bytes >> 20), flush=True)
be.close(); ctx.close(); dist.destroy_process_group()
