#!/bin/bash
# Round 2, GPU call 5 (1 GPU): full-size parity tests on the two-group converter GEMM, then the 1-GPU bench on the generator's and on a
# Morton-ordered mesh (short timeouts: a hang must not eat the budget).
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_zz_fullsize.py tests/test_gpu_kernels.py -q -m gpu --maxfail=4 -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|full-size|FAILED|Error" gpurun_out/pytest_gpu_$TAG.log | tail -12
for ORD in given morton; do
timeout 150 python bench.py --no-extras --no-cpu-baseline --mesh-order $ORD > gpurun_out/bench_${TAG}_$ORD.json 2> gpurun_out/bench_${TAG}_$ORD.err; echo "bench $ORD rc=$?"; tail -2 gpurun_out/bench_${TAG}_$ORD.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$ORD.json").read().strip().splitlines()[-1])
    print("$ORD", {k:d[k] for k in ("ms_per_step","clocks") if k in d}, "layer", d.get("gcnconv_layer",{}).get("hbm_frac"))
    for k,v in d.get("kernel_families",{}).items(): print("  ",k, {a:round(b,2) for a,b in v.items()})
    for k,v in d.get("kernel_shapes_top",{}).items(): print("  ",k, {a:round(b,2) for a,b in v.items()})
except Exception as e: print("parse failed", e)
PY
done
