#!/bin/bash
# SpMM register-cache experiment (rows of v-1 / v kept in registers): A/B builds of spmm.cu via SGB_LIB_PATH.
mkdir -p gpurun_out
V=semigcn_b200/csrc/variants
for v in norc default rc16 rc8 rc32b3; do
    echo "##### variant $v"
    if [ $v = default ]; then unset SGB_LIB_PATH; else export SGB_LIB_PATH=$PWD/$V/lib_$v.so; fi
    timeout 300 python tools/bench_kernels.py spmm 2>&1 | grep -v "^vertices" | grep "mode=0"
    timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "spmm or graph" 2>&1 | tail -1
done > gpurun_out/exp_rc.txt 2>&1
cat gpurun_out/exp_rc.txt
unset SGB_LIB_PATH
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rc.json 2>gpurun_out/bench_rc.err
python -c "
import json; d=json.load(open('gpurun_out/bench_rc.json')); print(d['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in d['kernel_families'].items()})"
