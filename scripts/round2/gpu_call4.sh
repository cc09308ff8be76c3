#!/bin/bash
# round 2, call 4 (1 GPU): L2 evict_last hints in the apply kernels, pair-slice size sweep, spill path, smem-slice microbenchmark,
# ncu --set full + source of one round (2 M reads)
set -x
mkdir -p gpurun_out
RB_TEST_SPILL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "spill or skewed or duplicates or paired" > gpurun_out/r2c4_spill.log 2>&1; echo "spill rc=$?" >> gpurun_out/r2c4_spill.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
RB_SLICE_PAIR_LOG2=24 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c4_bench_p24.json 2> gpurun_out/r2c4_bench_p24.err
RB_SLICE_PAIR_LOG2=24 RB_SLICE_RAISE_LOG2=24 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c4_bench_p24r24.json 2> gpurun_out/r2c4_bench_p24r24.err
RB_SLICED_CHUNK=1024 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c4_bench_c1024.json 2> gpurun_out/r2c4_bench_c1024.err
timeout 300 ./scripts/microbench_smem_slice > gpurun_out/r2c4_microbench_smem.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ks_ -s 42 -c 14 --csv --log-file gpurun_out/r2c4_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c4_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ks_ -s 42 -c 14 -f -o /tmp/r2c4_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --reads-per-step 2000000 > gpurun_out/r2c4_ncu_full.log 2>&1
ncu -i /tmp/r2c4_full.ncu-rep --page details --csv > gpurun_out/r2c4_details.csv 2> gpurun_out/r2c4_ncu_export.err
ncu -i /tmp/r2c4_full.ncu-rep --page raw --csv > gpurun_out/r2c4_raw.csv 2>> gpurun_out/r2c4_ncu_export.err
ncu -i /tmp/r2c4_full.ncu-rep --page source --csv --print-source sass > gpurun_out/r2c4_source_sass.csv 2>> gpurun_out/r2c4_ncu_export.err
ncu -i /tmp/r2c4_full.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2c4_source_cuda.csv 2>> gpurun_out/r2c4_ncu_export.err
ls -la /tmp/r2c4_full.ncu-rep gpurun_out/ | tail -20
tail -n 4 gpurun_out/r2c4_spill.log; cat gpurun_out/r2c4_microbench_smem.txt
python - <<'PY'
import json
for n in ("bench", "bench_p24", "bench_p24r24", "bench_c1024"):
    try:
        d = json.loads(open("gpurun_out/r2c4_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f e2e %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"], d["e2e"] and d["e2e"]["value"] / 1e9))
        print("  ", r["kernels_ms_per_step"])
    except Exception as e:
        print(n, "failed", e)
PY
