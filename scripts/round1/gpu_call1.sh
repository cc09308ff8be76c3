#!/bin/bash
# round-1 GPU call: sliced engine parity + bench + launch list + geometry sweep
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_gpu.txt 2>&1
echo "== pytest sliced" ; date +%s
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sliced" > gpurun_out/c1_pytest_sliced.log 2>&1; echo "pytest exit $?" >> gpurun_out/c1_pytest_sliced.log
tail -5 gpurun_out/c1_pytest_sliced.log
echo "== bench sliced" ; date +%s
timeout 300 python bench.py --engine sliced --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_sliced.json 2> gpurun_out/c1_bench_sliced.err; echo "exit $?"
tail -c 1500 gpurun_out/c1_bench_sliced.json
echo "== sweep" ; date +%s
for cfg in "28 25 28" "30 27 28" "29 26 29" "29 26 27"; do
  set -- $cfg
  RB_SLICE_BITS_LOG2=$1 RB_SLICE_BYTES_LOG2=$2 RB_SLICED_ROUND_LOG2=$3 timeout 200 python bench.py --engine sliced --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c1_sweep_$1_$2_$3.json 2> gpurun_out/c1_sweep_$1_$2_$3.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/c1_sweep_$1_$2_$3.json"))
    print("$cfg", round(d["value"]/1e9, 3), d["roofline"]["insert_gkmers_s"], d["roofline"]["lookup_gkmers_s"])
except Exception as e:
    print("$cfg failed", e)
PY
done
echo "== ncu launch list" ; date +%s
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r01_sliced_launches.csv python bench.py --engine sliced --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --reads-per-step 2000000 > gpurun_out/c1_ncu_bench.log 2>&1; echo "ncu exit $?"
echo "== bench direct (comparison)" ; date +%s
timeout 200 python bench.py --engine direct --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c1_bench_direct.json 2> gpurun_out/c1_bench_direct.err; echo "exit $?"
tail -c 600 gpurun_out/c1_bench_direct.json
date +%s
