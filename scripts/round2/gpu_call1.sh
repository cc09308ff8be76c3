#!/bin/bash
# round 2, call 1 (2 GPUs): the N>1 data path against the oracle on real GPUs, inside the command the driver runs.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2c1_gpu.txt
free -g >> gpurun_out/r2c1_gpu.txt; nproc >> gpurun_out/r2c1_gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_nccl.py > gpurun_out/r2c1_check.log 2>&1
echo "check rc=$?" >> gpurun_out/r2c1_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2c1_bench_n2.json 2> gpurun_out/r2c1_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2c1_check.log
tail -3 gpurun_out/r2c1_check.log; cat gpurun_out/r2c1_bench_n2.json | cut -c1-1500
