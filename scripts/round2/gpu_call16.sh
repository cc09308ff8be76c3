#!/bin/bash
# round 2, call 16 (2 GPUs): in-place raises + cells in the sharded graph: peer-to-peer bench with the oracle parity check, then the staged
# (NCCL all-to-all) exchange with the parity check as well (the raise bytes now travel like the probes did)
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2c16_bench_n2.json 2> gpurun_out/r2c16_bench_n2.err
echo "bench rc=$?"
RB_MGRAPH_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2c16_bench_n2_staged.json 2> gpurun_out/r2c16_bench_n2_staged.err
echo "bench staged rc=$?"
python - <<'PY'
import json
for n in ("bench_n2", "bench_n2_staged"):
    try:
        d = json.loads(open("gpurun_out/r2c16_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(n, "value %.3f G  insert %.2f lookup %.2f ms/step %.1f e2e %s" % (d["value"] / 1e9, r["insert_gkmers_s"], r["lookup_gkmers_s"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"] / 1e9))
        print("  exchange ms", r["exchange_ms_per_step"], "GB", r["exchange_gb_per_rank_per_step"], r["kernels_ms_per_step"])
        print("  ", d["config"].get("exchange"), "parity", (d.get("parity_check") or {}))
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/r2c16_bench_n2.err; tail -5 gpurun_out/r2c16_bench_n2_staged.err
