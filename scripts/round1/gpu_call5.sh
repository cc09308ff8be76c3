#!/bin/bash
# round-1 GPU call 5: in-order chunk dispensing, leaner tile sort, staged answers: parity subset, window sweep, DRAM bytes per kernel
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest sliced subset" ; date +%s
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sliced and (uniform or duplicates or collision_free or getkmers or policies or full_size or subbatching)" > gpurun_out/c5_pytest_sliced.log 2>&1; echo "pytest exit $?" >> gpurun_out/c5_pytest_sliced.log
tail -3 gpurun_out/c5_pytest_sliced.log
echo "== sweep (bits bytes round chunk occ)" ; date +%s
for cfg in "28 25 28 4096 4" "28 25 28 2048 8" "28 25 28 1024 8" "28 25 28 2048 4" "28 25 29 2048 8" "29 26 29 1024 8"; do
  set -- $cfg
  f=gpurun_out/c5_sweep_$1_$2_$3_$4_$5
  RB_SLICE_BITS_LOG2=$1 RB_SLICE_BYTES_LOG2=$2 RB_SLICED_ROUND_LOG2=$3 RB_SLICED_CHUNK=$4 RB_SLICED_CONSUMER_OCC=$5 timeout 200 python bench.py --engine sliced --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $f.json 2> $f.err
  python - <<PY
import json
try:
    d = json.load(open("$f.json"))
    r = d["roofline"]
    print("$cfg", "value %.3f ins %.3f look %.3f" % (d["value"]/1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"]), {k: round(v, 1) for k, v in r["kernels_ms_per_step"].items()})
except Exception as e:
    print("$cfg failed", e); print(open("$f.err").read()[-600:])
PY
done
echo "== ncu launch list with DRAM bytes" ; date +%s
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ks_ -s 42 -c 14 --csv --log-file gpurun_out/r01_sliced_v2_launches.csv python bench.py --engine sliced --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --reads-per-step 2000000 > gpurun_out/c5_ncu.log 2>&1; echo "ncu exit $?"
du -sh gpurun_out; date +%s
