#!/bin/bash
# round-1 GPU call 7: full GPU suite, smoke, default bench (sliced) with e2e + cpu baseline, ncu details of one round, launch list at bench size
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu (all)" ; date +%s
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/c7_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/c7_pytest_gpu.log
tail -4 gpurun_out/c7_pytest_gpu.log
echo "== smoke" ; date +%s
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.log 2>&1; tail -2 gpurun_out/c7_smoke.log
echo "== bench default" ; date +%s
timeout 400 python bench.py > gpurun_out/bench_r01_sliced_n1.json 2> gpurun_out/c7_bench.err; echo "bench exit $?"
tail -c 2500 gpurun_out/bench_r01_sliced_n1.json
echo "== launch list at bench size" ; date +%s
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:ks_ -s 42 -c 14 --csv --log-file gpurun_out/r01_sliced_v3_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c7_ncu_list.log 2>&1; echo "ncu exit $?"
echo "== ncu full details, one round of 252M k-mers" ; date +%s
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ks_ -s 42 -c 14 -o /tmp/r01_sliced_v3_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --reads-per-step 2000000 > gpurun_out/c7_ncu_full.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/r01_sliced_v3_full.ncu-rep --page details --csv > gpurun_out/r01_sliced_v3_details.csv 2> gpurun_out/c7_ncu_export.err
ncu -i /tmp/r01_sliced_v3_full.ncu-rep --page raw --csv > gpurun_out/r01_sliced_v3_raw.csv 2>> gpurun_out/c7_ncu_export.err
du -sh gpurun_out; date +%s
