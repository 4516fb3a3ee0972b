#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): the default bench line at N = 2 under torchrun (replicas + partition block + batch64).
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "bench n2 rc=$?"; grep -v "Warning\|warn" gpurun_out/bench_${TAG}_n2.err | tail -15 | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_n2.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("ms_per_step","value","n_gpus","clocks") if k in d})
    print("partition", json.dumps(d.get("partition"), indent=0)[:3500])
    print("batch64", d.get("batch64"))
except Exception as e: print("parse failed", e)
PY
