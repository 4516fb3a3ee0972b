for st in 2 3; do echo "== stages $st raw3"; SGB_F16_STAGES=$st timeout 60 python tools/debug_f16.py 2>&1 | grep -E "^m=" | tail -3; done
for st in 2 3; do echo "== stages $st raw2"; SGB_LIB_PATH=$PWD/semigcn_b200/csrc/variants/lib_raw2.so SGB_F16_STAGES=$st timeout 60 python tools/debug_f16.py 2>&1 | grep -E "^m=" | tail -3; done
