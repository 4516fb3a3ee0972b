#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): dist + kernel tests, 1-GPU bench (two-group converter; Morton-ordered mesh experiment), then the default
# bench line at N = 2 under torchrun (replicas + partition block + batch64).
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_kernels.py tests/test_gpu_losses.py -q -m gpu --maxfail=6 -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|world .* slope|FAILED|Error" gpurun_out/pytest_gpu_$TAG.log | tail -16
for ORD in given morton; do
timeout 300 python bench.py --no-extras --no-cpu-baseline --mesh-order $ORD > gpurun_out/bench_${TAG}_$ORD.json 2> gpurun_out/bench_${TAG}_$ORD.err; echo "bench $ORD rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}_$ORD.json").read().strip().splitlines()[-1])
print("$ORD", {k:d[k] for k in ("ms_per_step","clocks") if k in d}, "layer", d.get("gcnconv_layer",{}).get("hbm_frac"))
for k,v in d.get("kernel_families",{}).items(): print("  ",k, {a:round(b,2) for a,b in v.items()})
for k,v in d.get("kernel_shapes_top",{}).items(): print("  ",k, {a:round(b,2) for a,b in v.items()})
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_${TAG}_n2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_n2.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("ms_per_step","value","n_gpus","clocks") if k in d})
    print("partition", json.dumps(d.get("partition"), indent=0)[:3000])
    print("batch64", d.get("batch64"))
except Exception as e: print("parse failed", e)
PY
