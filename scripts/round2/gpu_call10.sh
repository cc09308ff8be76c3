#!/bin/bash
# round 2, call 10 (1 GPU): full GPU suite (new: IUPAC, cascade, sync_to_host, full-size oracle comparison, configs 3/4 settings), bench configs 1, 3, 4
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c10_pytest.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err
timeout 900 python bench.py --config 3 --steps 6 --warmup 3 > gpurun_out/r2c10_bench_cfg3.json 2> gpurun_out/r2c10_bench_cfg3.err
timeout 900 python bench.py --config 4 --steps 6 --warmup 3 > gpurun_out/r2c10_bench_cfg4.json 2> gpurun_out/r2c10_bench_cfg4.err
tail -n 14 gpurun_out/r2c10_pytest.log
tail -n 3 gpurun_out/r2c10_bench_cfg3.err gpurun_out/r2c10_bench_cfg4.err
python - <<'PY'
import json
for n in ("bench", "bench_cfg3", "bench_cfg4"):
    try:
        d = json.loads(open("gpurun_out/r2c10_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f e2e %s cpu %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"], d["e2e"] and d["e2e"]["value"] / 1e9, d["cpu_baseline"] and d["cpu_baseline"]["value"] / 1e6))
        print("  ", r["kernels_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
