#!/bin/bash
# round 2, call 23 (1 GPU, final): the whole GPU suite, smoke, the default bench (e2e pipelined + blocking + ascii, cpu baseline), the reference arm,
# ncu launch list + full capture of one round, then configs[3] / configs[4]
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2c23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c23_pytest.log
tail -n 4 gpurun_out/r2c23_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c23_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c23_smoke.log; tail -n 2 gpurun_out/r2c23_smoke.log
timeout 600 python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/r2c23_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_final_reference.json 2> gpurun_out/r2c23_ref.err; echo "ref rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ks_ -s 42 -c 14 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c23_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ks_ -s 42 -c 14 -f -o /tmp/r02_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --reads-per-step 2000000 > gpurun_out/r2c23_ncu_full.log 2>&1
ncu -i /tmp/r02_final.ncu-rep --page details --csv > gpurun_out/r02_final_ncu_details.csv 2> gpurun_out/r2c23_ncu_export.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r02_final_n1.json").read().strip().splitlines()[-1])
r = d["roofline"]; e = d["e2e"]
print("bench value %.3f G  insert %.2f lookup %.2f  frac %.3f step_frac %.3f e2e %.3f blocking %.3f ascii %.3f cpu %.1f M" % (
    d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], r["frac"], r["step_frac"], e["value"] / 1e9, e["blocking_calls"]["value"] / 1e9,
    e["ascii"]["value"] / 1e9, d["cpu_baseline"]["value"] / 1e6))
print("  ", r["kernels_ms_per_step"])
print(open("gpurun_out/bench_r02_final_reference.json").read()[:300])
PY
timeout 200 python bench.py --config 3 --steps 6 > gpurun_out/bench_r02_final_cfg3.json 2> gpurun_out/r2c23_cfg3.err; echo "cfg3 rc=$?"
timeout 200 python bench.py --config 4 --steps 6 > gpurun_out/bench_r02_final_cfg4.json 2> gpurun_out/r2c23_cfg4.err; echo "cfg4 rc=$?"
tail -c 600 gpurun_out/bench_r02_final_cfg3.json; tail -c 600 gpurun_out/bench_r02_final_cfg4.json
