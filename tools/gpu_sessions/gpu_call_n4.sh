#!/bin/bash
# 4-GPU call: the 4-rank slab test and a short N=4 bench line (weak + parity + strong blocks).
set -u
mkdir -p gpurun_out
{
  echo "== slab tests"; timeout 900 python -m pytest tests/test_slab_gpu.py -q -m gpu -x -k "4" 2>&1 | tail -30
  echo "== bench N=4"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 20 --warmup 3 --no_cpu_baseline > gpurun_out/bench_n4.out 2>&1; tail -5 gpurun_out/bench_n4.out | cut -c1-400; tail -1 gpurun_out/bench_n4.out > gpurun_out/bench_n4.json
} > gpurun_out/call_n4.log 2>&1
tail -30 gpurun_out/call_n4.log
