#!/bin/bash
# 1-GPU call: partition-mode N=1 line (strong-scaling base), parity tests, bench, ncu capture of the restructured SpMM.
TAG=${1:-r1b}
mkdir -p gpurun_out
timeout 600 python bench.py --mode partition --steps 5 --warmup 3 > gpurun_out/bench_partition_n1.json 2> gpurun_out/bench_partition_n1.err
cat gpurun_out/bench_partition_n1.json
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$TAG.log 2>&1
tail -2 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_spmm -s 5 -c 1 -f -o gpurun_out/prof_${TAG}_spmm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/prof_${TAG}_spmm.stdout 2>&1
