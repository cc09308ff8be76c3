#!/bin/bash
# round 2, call 6 (2 GPUs): library-owned NCCL exchange (rb_mgraph_*): oracle parity over NCCL, then the N=2 bench with its parity check
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_nccl.py > gpurun_out/r2c6_check.log 2>&1
echo "check rc=$?" >> gpurun_out/r2c6_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2c6_bench_n2.json 2> gpurun_out/r2c6_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2c6_check.log
RB_BENCH_READS_PER_ROUND=2000000 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --no-parity --no-e2e > gpurun_out/r2c6_bench_n2_r2m.json 2> gpurun_out/r2c6_bench_n2_r2m.err
echo "bench2 rc=$?" >> gpurun_out/r2c6_check.log
grep -E "parity ok|rc=|Error|error" gpurun_out/r2c6_check.log | tail -8
python - <<'PY'
import json
for n in ("bench_n2", "bench_n2_r2m"):
    try:
        d = json.loads(open("gpurun_out/r2c6_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f ms/step %.1f e2e %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"] / 1e9))
        print("  ", d["config"]["exchange"], d.get("parity_check"))
    except Exception as e:
        print(n, "failed", e)
PY
