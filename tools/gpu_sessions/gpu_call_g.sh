#!/bin/bash
# 2-GPU call: slab tests (decomposed = undecomposed, peer-memory = NCCL, replay = eager) with the TMA-fed transposed
# interpolation, then the N=2 bench line (weak, with its parity and strong blocks).
set -u
mkdir -p gpurun_out
{
  echo "== slab tests"; timeout 900 python -m pytest tests/test_slab_gpu.py -q -m gpu -x 2>&1 | tail -60
  echo "== bench N=2"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 3 --no_cpu_baseline > gpurun_out/bench_n2_g.out 2>&1; tail -25 gpurun_out/bench_n2_g.out | cut -c1-600; tail -1 gpurun_out/bench_n2_g.out > gpurun_out/bench_n2_g.json
} > gpurun_out/call_g.log 2>&1
tail -30 gpurun_out/call_g.log
