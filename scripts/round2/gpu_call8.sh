#!/bin/bash
# round 2, call 8 (2 GPUs): peer-to-peer mode (CUDA IPC: consumer kernels read the producers' arenas over NVLink): parity + bench, A/B with the staged exchange
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_nccl.py > gpurun_out/r2c8_check.log 2>&1
echo "check rc=$?" >> gpurun_out/r2c8_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2c8_bench_n2.json 2> gpurun_out/r2c8_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2c8_check.log
RB_MGRAPH_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 4 --warmup 3 --no-parity --no-e2e > gpurun_out/r2c8_bench_n2_staged.json 2> gpurun_out/r2c8_bench_n2_staged.err
grep -E "parity ok|rc=|Error|error" gpurun_out/r2c8_check.log | tail -8
python - <<'PY'
import json
for n in ("bench_n2", "bench_n2_staged"):
    try:
        d = json.loads(open("gpurun_out/r2c8_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f ms/step %.1f e2e %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"] / 1e9))
        print("  exchange ms", r["exchange_ms_per_step"], "GB", r["exchange_gb_per_rank_per_step"], r["kernels_ms_per_step"])
        print("  ", d["config"].get("exchange"), "parity", (d.get("parity_check") or {}))
    except Exception as e:
        print(n, "failed", e)
PY
