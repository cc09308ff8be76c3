#!/bin/bash
# round-1 GPU call 10: sanity of the final tree on one GPU (the single-GPU engine after the sharded-graph changes)
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(sliced and (getkmers or uniform or duplicates)) or (direct and full_size)" > gpurun_out/c10_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c10_pytest.log
tail -3 gpurun_out/c10_pytest.log
timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/c10_bench.json')); print(d['value']/1e9, d['roofline']['insert_gkmers_s'], d['roofline']['lookup_gkmers_s'])"
