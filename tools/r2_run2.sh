#!/bin/bash
# Round 2, GPU call 2: parity tests, smoke, bench line, full ncu capture of the reworked k_gemm_f16 (8 epilogue warps).
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --maxfail=12 -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|full-size|world .* slope|FAILED|Error" gpurun_out/pytest_gpu_$TAG.log | tail -30
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/smoke_$TAG.log
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","value","clocks","gcnconv_layer") if k in d})
print("e2e", d.get("e2e"))
for k,v in d.get("kernel_families",{}).items(): print(k, {a:round(b,2) for a,b in v.items()})
for k,v in d.get("kernel_shapes_top",{}).items(): print(k, {a:round(b,2) for a,b in v.items()})
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^k_gemm_f16 -s 4 -c 2 -f -o gpurun_out/prof_${TAG}_gemm_f16 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/prof_${TAG}_gemm_f16.stdout 2>&1
ls -la gpurun_out | tail -4
