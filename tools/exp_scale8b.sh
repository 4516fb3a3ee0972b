#!/bin/bash
# 8-GPU box, after the fast halo SpMM instantiation: 16M- and 4M-vertex meshes at N=8.
mkdir -p gpurun_out
run() {  # n freq tag
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + $1)) \
      bench.py --gpus $1 --mode partition --freq $2 --steps 5 --warmup 3 2> gpurun_out/bench_partition_$3.err | grep "^{" > gpurun_out/bench_partition_$3.json
  echo "== $3 rc=$?"; cut -c1-300 gpurun_out/bench_partition_$3.json
}
run 8 1265 n8_16m_v2
run 8 632 n8_4m_v2
