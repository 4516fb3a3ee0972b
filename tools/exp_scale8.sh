#!/bin/bash
# 8-GPU box: strong scaling of the vertex-partitioned mode (4M-vertex mesh at N=4,8; 16M-vertex mesh at N=8,4).
mkdir -p gpurun_out
run() {  # n freq tag
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1)) \
      bench.py --gpus $1 --mode partition --freq $2 --steps 5 --warmup 3 > gpurun_out/bench_partition_$3.json 2> gpurun_out/bench_partition_$3.err
  echo "== $3 rc=$?"; cut -c1-400 gpurun_out/bench_partition_$3.json; tail -2 gpurun_out/bench_partition_$3.err | cut -c1-300
}
run 8 632 n8_4m
run 4 632 n4_4m
run 8 1265 n8_16m
run 4 1265 n4_16m
