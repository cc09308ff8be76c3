#!/bin/bash
# round 2, call 11 (1 GPU): GPU suite after the fixture fix, config 4 with the walker at 2 CTAs/SM
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2c11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c11_pytest.log
timeout 900 python bench.py --config 4 --steps 6 --warmup 3 > gpurun_out/r2c11_bench_cfg4.json 2> gpurun_out/r2c11_bench_cfg4.err
timeout 900 python bench.py --config 3 --steps 6 --warmup 3 > gpurun_out/r2c11_bench_cfg3.json 2> gpurun_out/r2c11_bench_cfg3.err
tail -n 12 gpurun_out/r2c11_pytest.log
python - <<'PY'
import json
for n in ("bench_cfg3", "bench_cfg4"):
    try:
        d = json.loads(open("gpurun_out/r2c11_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f cpu %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"], d["cpu_baseline"] and d["cpu_baseline"]["value"] / 1e6))
        print("  ", r["kernels_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
