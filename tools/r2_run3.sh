#!/bin/bash
# Round 2, GPU call 3 (1 GPU): parity tests (bnf loss, mask dilation, overlapped halo + normal loss across cuts), smoke, the default
# bench line with its extra blocks (batch64, small meshes), the reference arm.
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --maxfail=12 -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|world .* slope|FAILED|Error" gpurun_out/pytest_gpu_$TAG.log | tail -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","value","clocks","gcnconv_layer","kernel_families_sum_ms","step_hbm_frac") if k in d})
print("e2e", d.get("e2e"))
for k,v in d.get("kernel_families",{}).items(): print(k, {a:round(b,2) for a,b in v.items()})
for k,v in d.get("kernel_shapes_top",{}).items(): print(k, {a:round(b,2) for a,b in v.items()})
print("batch64", d.get("batch64"))
print("small_meshes", json.dumps(d.get("small_meshes"), indent=0)[:1500])
print("cpu_baseline", d.get("cpu_baseline"))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_$TAG.json
