#!/bin/bash
# round 2, call 5 (1 GPU): register-trimmed tile sorts (3 CTAs/SM), spill test with a roomy filter
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spill or skewed or duplicates or paired" > gpurun_out/r2c5_spill.log 2>&1; echo "spill rc=$?" >> gpurun_out/r2c5_spill.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
tail -n 4 gpurun_out/r2c5_spill.log
python - <<'PY'
import json
for n in ("bench",):
    try:
        d = json.loads(open("gpurun_out/r2c5_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f e2e %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"], d["e2e"] and d["e2e"]["value"] / 1e9))
        print("  ", r["kernels_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
