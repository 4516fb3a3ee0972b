#!/bin/bash
# A/B builds of one translation unit with extra -D flags, linked against the other objects of the regular build:
#   tools/build_variant.sh <name> <file.cu> [-DFOO=1 ...]  ->  semigcn_b200/csrc/variants/lib_<name>.so  (use with SGB_LIB_PATH)
set -e
cd "$(dirname "$0")/../semigcn_b200/csrc"
name=$1; src=$2; shift 2
mkdir -p variants
base=$(basename "$src" .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr "$@" -c "$src" -o variants/${base}_$name.o
others=$(ls build/*.o | grep -v "build/${base}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so variants/${base}_$name.o $others -lcudart
