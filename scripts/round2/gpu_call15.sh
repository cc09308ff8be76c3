#!/bin/bash
# round 2, call 15 (1 GPU): co-located cells (one access per paired record) + 32 MiB pair slices + f4 screening-filter ops:
# the whole GPU suite, the default bench, A/B without cells, an ncu metrics pass, where the ASCII entry points spend their time
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c15_pytest.log
tail -n 10 gpurun_out/r2c15_pytest.log
timeout 900 python bench.py > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; echo "bench rc=$?"
RB_SLICED_CELLS=0 timeout 600 python bench.py --steps 8 --no-e2e --no-cpu-baseline > gpurun_out/r2c15_bench_nocells.json 2> gpurun_out/r2c15_bench_nocells.err; echo "bench nocells rc=$?"
RB_SLICE_PAIR_LOG2=25 timeout 600 python bench.py --steps 8 --no-e2e --no-cpu-baseline > gpurun_out/r2c15_bench_p25.json 2> gpurun_out/r2c15_bench_p25.err; echo "bench p25 rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ks_ -s 42 -c 14 --csv --log-file gpurun_out/r2c15_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c15_ncu_list.log 2>&1
timeout 300 python scripts/time_ascii.py > gpurun_out/r2c15_time_ascii.log 2>&1; tail -8 gpurun_out/r2c15_time_ascii.log
python - <<'PY'
import json
for n in ("bench", "bench_nocells", "bench_p25"):
    try:
        d = json.loads(open("gpurun_out/r2c15_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        e = d.get("e2e") or {}
        print(n, "value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"]))
        print("   e2e", e.get("value"), e.get("ms_per_step"), "blocking", e.get("blocking_calls"), "ascii", (e.get("ascii") or {}).get("value"))
        print("  ", r["kernels_ms_per_step"])
    except Exception as ex:
        print(n, "failed", ex)
PY
tail -3 gpurun_out/r2c15_bench.err
