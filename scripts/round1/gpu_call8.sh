#!/bin/bash
# round-1 GPU call 8 (2 GPUs): the sharded graph on the sliced engine -- single-rank parity on the GPU, then the 2-rank bench over NCCL
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest sharded sliced (1 GPU)" ; date +%s
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sharded_sliced_graph_single_rank and direct" > gpurun_out/c8_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c8_pytest.log
tail -3 gpurun_out/c8_pytest.log
echo "== bench --gpus 2 (sliced sharded)" ; date +%s
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_r01_sliced_n2.json 2> gpurun_out/c8_bench.err; echo "bench exit $?"
tail -c 1800 gpurun_out/bench_r01_sliced_n2.json; tail -5 gpurun_out/c8_bench.err
date +%s
