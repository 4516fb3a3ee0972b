#!/bin/bash
# One GPU call (1 GPU): parity tests, both bench arms, ncu launch list, ncu --set full of the hot kernels.
# Outputs under gpurun_out/ (scratch); the summaries that are judged are copied into profiles/ afterwards.
TAG=${1:-r1}
mkdir -p gpurun_out
# 0. parity tests through the C-ABI
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
# 1. bench lines (reference arm first, as the driver does)
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 600 python tools/step_breakdown.py > gpurun_out/step_breakdown_$TAG.txt 2>&1
# 2. launch list: every launch of OUR kernels (all named k_*) with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 900 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/launches_$TAG.stdout 2>&1
# 3. full captures of the widest launches of the four dominant families (one process each; -s skips the narrow early layers)
for spec in "k_spmm:5:spmm" "k_gemm_f16:4:gemm_f16" "k_gemm_tn_f16:3:gemm_tn_f16" "k_bn_act_bwd_apply:5:bn_act_bwd"; do
    IFS=: read kern skip name <<< "$spec"
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$kern -s $skip -c 1 -f -o gpurun_out/prof_${TAG}_$name \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/prof_${TAG}_$name.stdout 2>&1
done
ls -la gpurun_out/
