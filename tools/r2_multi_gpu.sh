#!/bin/bash
# The default bench line at N GPUs under torchrun, as the driver launches it (replicas + partition block + batch64):
#   gpurun --gpus N -- 'bash tools/r2_multi_gpu.sh N'      -> gpurun_out/bench_r2_nN.json
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "bench n$N rc=$?"; grep -v "Warning\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_r2_n$N.err | tail -10 | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","value","n_gpus","clocks") if k in d})
for name,b in d.get("partition",{}).items(): print(name, {k:v for k,v in b.items() if k not in ("collectives",)})
print("batch64", d.get("batch64"))
PY
