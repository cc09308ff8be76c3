#!/bin/bash
# round 2, call 13 (1 GPU): the whole GPU suite, smoke, the default bench (with e2e, e2e.ascii, cpu baseline), reference arm, ncu launch list + full capture
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c13_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c13_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c13_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/r2c13_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/r2c13_ref.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ks_ -s 42 -c 14 --csv --log-file gpurun_out/r02_sliced_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c13_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ks_ -s 42 -c 14 -f -o /tmp/r02_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --reads-per-step 2000000 > gpurun_out/r2c13_ncu_full.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page details --csv > gpurun_out/r02_sliced_ncu_details.csv 2> gpurun_out/r2c13_ncu_export.err
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_sliced_ncu_raw.csv 2>> gpurun_out/r2c13_ncu_export.err
tail -n 12 gpurun_out/r2c13_pytest.log; tail -n 2 gpurun_out/r2c13_smoke.log
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_n1.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("bench value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f e2e %s ascii %s cpu %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"], d["e2e"]["value"] / 1e9, d["e2e"].get("ascii", {}).get("value", 0) / 1e9, d["cpu_baseline"]["value"] / 1e6))
print("  ", r["kernels_ms_per_step"])
print(open("gpurun_out/bench_r02_reference.json").read()[:600])
PY
