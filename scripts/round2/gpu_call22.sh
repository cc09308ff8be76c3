#!/bin/bash
# round 2, call 22 (8 GPUs): sharded graph at N=8 with in-place raises, cells and cp.async-staged pull over NVLink (oracle parity check included)
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2c22_bench_n8.json 2> gpurun_out/r2c22_bench_n8.err
echo "bench rc=$?"
python - <<'PY'
import json
for n in ("bench_n8",):
    try:
        d = json.loads(open("gpurun_out/r2c22_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f ms/step %.1f e2e %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"] / 1e9))
        print("  exchange ms", r["exchange_ms_per_step"], r["kernels_ms_per_step"])
        print("  ", d["config"].get("exchange"), "parity", (d.get("parity_check") or {}))
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/r2c22_bench_n8.err
