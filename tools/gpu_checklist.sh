#!/bin/bash
# One gpurun call that answers the open questions of the last round; writes everything to
# gpurun_out/.  Usage:  gpurun --timeout 900 -- 'bash tools/gpu_checklist.sh'
set -u
mkdir -p gpurun_out
{
  echo "== parity (all GPU tests, incl. the configs[2] wave case added without a GPU)"
  timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -5
  echo "== configs[2] wave case and k_tile3d"
  timeout 200 python -m pytest tests/test_zz_wave2_gpu.py -q -m gpu 2>&1 | tail -8
  echo "== smoke"
  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== secondary configurations (configs[1] eager / graph, wave (t,x), Newton-CG)"
  timeout 200 python tools/bench_configs.py 1 2 4 2>&1 | grep -v Running
  echo "== configs[2] at full size: k_tile3d (default), then the per-cell kernel"
  timeout 200 python tools/bench_configs.py 3 2>&1 | grep -v Running
  ODIL_B200_TILE3D=0 timeout 200 python tools/bench_configs.py 3 2>&1 | grep -v Running
  timeout 60 python tools/time_tile3d.py 256 512 512
  echo "== headline"
  timeout 240 python bench.py --steps 30 --warmup 3 2>&1 | tail -1
} > gpurun_out/checklist.log 2>&1
tail -40 gpurun_out/checklist.log
