"""Parity of the hash-sharded graph (rb_mgraph_*: library-owned NCCL all-to-all) on >= 2 real GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_nccl.py
Runs bench_multi.sharded_parity_check -- the same oracle comparison `bench.py --gpus N` performs after its timed region -- without the bench:
every rank inserts its own reads; the gathered shares must equal the sequential oracle's arrays, counts must equal the oracle's."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as single  # noqa: E402
import bench_multi  # noqa: E402
import rnabloom_b200 as rb  # noqa: E402
from rnabloom_b200.sharded import ShardedGraph, broadcast_nccl_id  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = rb.Context(local)
dev = torch.device("cuda", local)


def make_graph(dbg_bits, cbf_bytes, max_kmers):
    return ShardedGraph(ctx, world, rank, dbg_bits, cbf_bytes, single.HD, single.HC, single.K, False, max_kmers, nccl_id=broadcast_nccl_id(dev))


full_d, full_c = single.DBG_BITS * world, single.CBF_BYTES * world
full = make_graph(full_d, full_c, 30000 * 126)
res = bench_multi.sharded_parity_check(make_graph, rank, world, dev, full_d, full_c, full)
if rank == 0:
    print("sharded NCCL parity ok: " + json.dumps(res), flush=True)
full.close()
ctx.close()
dist.destroy_process_group()
