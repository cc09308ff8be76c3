#!/bin/bash
# round-1 GPU call 6: strided item mapping, parallel run staging, match.any ranking experiment
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest sliced mini" ; date +%s
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sliced and (uniform or getkmers or duplicates)" > gpurun_out/c6_pytest_sliced.log 2>&1; echo "pytest exit $?" >> gpurun_out/c6_pytest_sliced.log
tail -3 gpurun_out/c6_pytest_sliced.log
echo "== sweep (bits bytes round chunk occ rank)" ; date +%s
for cfg in "29 26 29 1024 8 atoms" "29 26 29 1024 8 match" "29 26 29 2048 8 atoms" "29 26 28 1024 8 atoms" "30 27 29 1024 8 atoms" "28 25 29 1024 8 atoms"; do
  set -- $cfg
  f=gpurun_out/c6_sweep_$1_$2_$3_$4_$5_$6
  RB_SLICE_BITS_LOG2=$1 RB_SLICE_BYTES_LOG2=$2 RB_SLICED_ROUND_LOG2=$3 RB_SLICED_CHUNK=$4 RB_SLICED_CONSUMER_OCC=$5 RB_SLICED_RANK=$6 timeout 200 python bench.py --engine sliced --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $f.json 2> $f.err
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
du -sh gpurun_out; date +%s
