#!/bin/bash
# round-1 GPU call 2: consumer-window sweep of the sliced engine + ncu --set full of one round
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest sliced subset" ; date +%s
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sliced and (duplicates or collision_free or getkmers or policies or full_size)" > gpurun_out/c2_pytest_sliced.log 2>&1; echo "pytest exit $?" >> gpurun_out/c2_pytest_sliced.log
tail -3 gpurun_out/c2_pytest_sliced.log
echo "== sweep (bits bytes round chunk occ)" ; date +%s
for cfg in "29 26 28 2048 2" "28 25 28 2048 2" "27 24 28 2048 2" "28 25 28 1024 2" "28 25 28 2048 1" "28 25 28 4096 4" "28 25 29 2048 2" "29 26 29 1024 2"; do
  set -- $cfg
  f=gpurun_out/c2_sweep_$1_$2_$3_$4_$5
  RB_SLICE_BITS_LOG2=$1 RB_SLICE_BYTES_LOG2=$2 RB_SLICED_ROUND_LOG2=$3 RB_SLICED_CHUNK=$4 RB_SLICED_CONSUMER_OCC=$5 timeout 200 python bench.py --engine sliced --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $f.json 2> $f.err
  python - <<PY
import json
try:
    d = json.load(open("$f.json"))
    r = d["roofline"]
    print("$cfg", "value %.3f ins %.3f look %.3f" % (d["value"]/1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"]), {k: round(v, 1) for k, v in r["kernels_ms_per_step"].items()})
except Exception as e:
    print("$cfg failed", e)
PY
done
echo "== ncu full, one round" ; date +%s
timeout 500 ncu --set full --clock-control none --import-source on -k regex:ks_ -s 45 -c 15 -o gpurun_out/r01_sliced_full python bench.py --engine sliced --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --reads-per-step 2000000 > gpurun_out/c2_ncu_full.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
date +%s
