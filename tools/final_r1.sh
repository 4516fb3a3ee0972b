#!/bin/bash
# Round-end verification inside a small GPU budget: parity tests, the bench line, one full ncu capture of the dominant
# kernel (DRAM traffic for the roofline block), the launch list.  Most important first; every step has its own timeout.
mkdir -p gpurun_out
timeout 110 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -2 gpurun_out/pytest_gpu_final.log
timeout 80 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 45 ncu --set full --clock-control none --import-source on -k regex:^k_spmm -s 5 -c 1 -f -o gpurun_out/prof_final_spmm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/prof_final_spmm.stdout 2>&1
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 900 --csv \
    --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-profile > gpurun_out/launches_final.stdout 2>&1
ls -la gpurun_out | tail -6
