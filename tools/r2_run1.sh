#!/bin/bash
# Round 2, GPU call 1: parity tests (incl. the new full-size oracle layer tests), smoke, bench line, launch list, and full ncu
# captures of the two tcgen05 kernels at their widest shapes.  Outputs under gpurun_out/ (scratch).
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; free -g | head -2; nproc
timeout 900 python -m pytest tests -q -m gpu --maxfail=12 --durations=8 -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
grep -E "passed|failed|full-size|world .* slope|FAILED|Error" gpurun_out/pytest_gpu_$TAG.log | tail -40
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke_$TAG.log
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json | cut -c1-3000
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 900 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/launches_$TAG.stdout 2>&1
for spec in "k_gemm_f16:4:gemm_f16" "k_gemm_tn_f16:3:gemm_tn_f16"; do
    IFS=: read kern skip name <<< "$spec"
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:^$kern -s $skip -c 2 -f -o gpurun_out/prof_${TAG}_$name \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/prof_${TAG}_$name.stdout 2>&1
done
ls -la gpurun_out | tail -8
