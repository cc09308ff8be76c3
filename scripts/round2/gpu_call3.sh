#!/bin/bash
# round 2, call 3 (1 GPU): spill path on the GPU, 4096-key tiles in the key sorts, smem-staged slice microbenchmark
set -x
mkdir -p gpurun_out
RB_TEST_SPILL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spill or skewed or duplicates" > gpurun_out/r2c3_spill.log 2>&1; echo "spill rc=$?" >> gpurun_out/r2c3_spill.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
RB_SLICED_SUBRANGE_LOG2=10 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c3_bench_sub10.json 2> gpurun_out/r2c3_bench_sub10.err
timeout 300 ./scripts/microbench_smem_slice > gpurun_out/r2c3_microbench_smem.txt 2>&1
tail -n 4 gpurun_out/r2c3_spill.log; cat gpurun_out/r2c3_microbench_smem.txt
python - <<'PY'
import json
for n in ("bench", "bench_sub10"):
    try:
        d = json.loads(open("gpurun_out/r2c3_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f e2e %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"], d["e2e"] and d["e2e"]["value"] / 1e9))
        print("  ", r["kernels_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
