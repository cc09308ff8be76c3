#!/bin/bash
# round-1 GPU call 11 (2 GPUs): sharded graph with tight capacities and rounds of 252 M k-mers per GPU
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e > gpurun_out/bench_r01_sliced_n2_v2.json 2> gpurun_out/c11_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_r01_sliced_n2_v2.json')); print(d['value']/1e9, d['roofline']['insert_gkmers_s'], d['roofline']['lookup_gkmers_s'], d['ms_per_step'], d['config']['exchange'])"
tail -3 gpurun_out/c11_bench.err
