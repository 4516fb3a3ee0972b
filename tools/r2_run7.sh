#!/bin/bash
# Round 2, GPU call 7 (8 GPUs): the default bench line at N = 8 under torchrun (replicas + partition block incl. the 16 M mesh + batch64).
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 \
    > gpurun_out/bench_${TAG}_n8.json 2> gpurun_out/bench_${TAG}_n8.err; echo "bench n8 rc=$?"; grep -v "Warning\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_${TAG}_n8.err | tail -15 | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_n8.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("ms_per_step","value","n_gpus","clocks") if k in d})
    p=d.get("partition",{})
    for name,b in p.items():
        if not isinstance(b,dict): print(name,b); continue
        print(name, {k:v for k,v in b.items() if k not in ("collectives","clocks")})
    print("batch64", d.get("batch64"))
except Exception as e: print("parse failed", e)
PY
