#!/bin/bash
# round 2, call 2 (1 GPU): paired probe records (SlShape<3>) -- GPU parity suite, spill path on a GPU for the first time, A/B bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
RB_TEST_SPILL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spill or skewed" > gpurun_out/r2c2_spill.log 2>&1; echo "spill rc=$?" >> gpurun_out/r2c2_spill.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2c2_bench_paired.json 2> gpurun_out/r2c2_bench_paired.err
RB_SLICED_PAIRED=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c2_bench_unpaired.json 2> gpurun_out/r2c2_bench_unpaired.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c2_ncu_bench.log 2>&1
tail -3 gpurun_out/r2c2_pytest.log gpurun_out/r2c2_spill.log
python - <<'PY'
import json
for n in ("paired", "unpaired"):
    try:
        d = json.loads(open("gpurun_out/r2c2_bench_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f e2e %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"], d["e2e"] and d["e2e"]["value"] / 1e9))
        print("  ", r["kernels_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
