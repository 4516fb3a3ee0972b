#!/bin/bash
# Which role bounds k_gemm_f16?  Builds ablated variants (the work of one role removed, its barrier protocol kept) and times the
# bench's transform shapes with each, on the same box.  Build here (no GPU needed), run on the GPU box:
#   tools/ablate_gemm.sh build      -> semigcn_b200/csrc/variants/lib_abl*.so
#   tools/ablate_gemm.sh run        -> gpurun_out/ablate_gemm.txt
set -e
cd "$(dirname "$0")/.."
names=(abl0 abl_nobcopy abl_nostore abl_noepi abl_noconv abl_notma abl_nomma abl_nobcopy_noepi abl_noconv_notma abl_onlymma)
flags=(0 1 2 4 8 16 32 5 24 31)
if [ "$1" = build ]; then
  for i in "${!names[@]}"; do tools/build_variant.sh ${names[$i]} gemm_f16.cu -DSGB_ABL=${flags[$i]} & done; wait
  ls -la semigcn_b200/csrc/variants/
else
  mkdir -p gpurun_out
  V=$PWD/semigcn_b200/csrc/variants
  for rep in 1 2; do
    for n in "${names[@]}"; do SGB_LIB_PATH=$V/lib_$n.so timeout 120 python tools/bench_gemm_shapes.py $n 2>&1 | tail -1; done
  done | tee gpurun_out/ablate_gemm.txt
fi
