#!/bin/bash
# A/B of two builds of the library on the SAME box, alternating:  tools/exp_ab.sh <what> <variantA> <variantB>
mkdir -p gpurun_out
V=$PWD/semigcn_b200/csrc/variants
what=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    echo "##### variant $v (rep $rep)"
    if [ $v = default ]; then unset SGB_LIB_PATH; else export SGB_LIB_PATH=$V/lib_$v.so; fi
    timeout 300 python tools/bench_kernels.py $what 2>&1 | grep -v "^vertices" | grep -v "mode=1"
  done
done > gpurun_out/exp_ab.txt 2>&1
cat gpurun_out/exp_ab.txt
