#!/bin/bash
# ncu passes of B200_PROFILING.md on the bench workload (1 GPU).  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
# 1. launch list: every launch of OUR kernels (all named k_*) with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 700 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/launches_r1.stdout 2>&1
# 2. full captures: the 256->512 forward transform (5th tensor-core GEMM launch), a 256-wide SpMM, the 256x256 weight gradient
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 4 -c 1 -f -o gpurun_out/prof_r1_gemm_tc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmm -s 4 -c 1 -f -o gpurun_out/prof_r1_spmm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tn_tc -s 4 -c 1 -f -o gpurun_out/prof_r1_gemm_tn_tc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > /dev/null 2>&1
ls -la gpurun_out/
