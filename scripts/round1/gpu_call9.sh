#!/bin/bash
# round-1 GPU call 9 (4 GPUs): the sharded graph on the sliced engine at world size 4
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_r01_sliced_n4.json 2> gpurun_out/c9_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench_r01_sliced_n4.json; tail -3 gpurun_out/c9_bench.err
