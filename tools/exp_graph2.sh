#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graphed.py -x -q -m gpu 2>&1 | tail -15
for f in 32 100 316; do
  for g in "" "--cuda-graph"; do
    timeout 600 python bench.py --freq $f --steps 20 --warmup 5 --no-cpu-baseline --no-profile $g 2>gpurun_out/bench_graph.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('freq $f graph=[$g]', 'ms/step', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])"
  done
done
