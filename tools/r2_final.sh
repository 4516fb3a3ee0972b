#!/bin/bash
# Round-2 verification (1 GPU): full parity suite, smoke, both bench arms, ncu launch list of the bench command, ncu --set full of the
# dominant kernel (SpMM at C = 256) and of the two tcgen05 kernels on CTA pairs (forward transform, weight gradient).  Outputs under gpurun_out/ (scratch) -> profiles/.
TAG=${1:-r2final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --maxfail=12 -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_$TAG.log | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke_$TAG.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("ms_per_step","value","clocks","kernel_families_sum_ms","step_hbm_frac","gpu_launches") if k in d})
print("e2e", d.get("e2e")); print("roofline", d.get("roofline")); print("layer", d.get("gcnconv_layer"))
for k,v in d.get("kernel_families",{}).items(): print(k, {a:round(b,2) for a,b in v.items()})
for k,v in d.get("kernel_shapes_top",{}).items(): print(k, {a:round(b,2) for a,b in v.items()})
PY
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1200 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile --no-extras > gpurun_out/launches_$TAG.stdout 2>&1
for spec in "k_spmm:5:spmm" "k_gemm_f16:4:gemm_f16" "k_gemm_tn_f16_pair:2:gemm_tn_pair"; do
    IFS=: read kern skip name <<< "$spec"
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:^$kern -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile --no-extras > gpurun_out/prof_${TAG}_$name.stdout 2>&1
done
ls -la gpurun_out | tail -6
