#!/bin/bash
# SpMM schedule experiment: interleaved (0) vs span schedule with 32 / 64 / 128-channel slices (SGB_SPMM_SPAN = 4 / 8 / 16).
mkdir -p gpurun_out
for v in 0 4 8 16; do
    echo "##### SGB_SPMM_SPAN=$v"
    SGB_SPMM_SPAN=$v timeout 300 python tools/bench_kernels.py spmm 2>&1 | grep -v "^vertices"
done > gpurun_out/exp_span.txt 2>&1
# correctness of the span kernels on the small test graphs (forced on)
for v in 4 8 16; do
    SGB_SPMM_SPAN=$v SGB_SPMM_SPAN_MIN_ROWS=0 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_modules.py -x -q -m gpu 2>&1 | tail -3
done > gpurun_out/exp_span_tests.txt 2>&1
cat gpurun_out/exp_span_tests.txt
# one full capture of the 32-channel-slice candidate at C=256 (5th plain launch at that width in the sweep)
SGB_SPMM_SPAN=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_spmm -s 4 -c 1 -f -o gpurun_out/prof_span4_c256 \
    python tools/bench_kernels.py spmm1 316 256 > gpurun_out/prof_span4.stdout 2>&1
cat gpurun_out/exp_span.txt
