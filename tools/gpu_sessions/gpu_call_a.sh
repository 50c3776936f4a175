#!/bin/bash
# Validation call: all GPU tests on the current head, smoke, tile3d timing (new / old kernel), headline bench, configs[2].
set -u
mkdir -p gpurun_out
{
  echo "== gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
  echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
  echo "== tile3d8"; timeout 60 python tools/time_tile3d.py 256 512 512
  echo "== tile3d old"; ODIL_B200_TILE3D_OLD=1 timeout 60 python tools/time_tile3d.py 256 512 512
  echo "== headline"; timeout 300 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1.json
  echo "== config 2"; timeout 300 python bench.py --config 2 2>&1 | tail -1 | tee gpurun_out/bench_config2.json
} > gpurun_out/call_a.log 2>&1
tail -30 gpurun_out/call_a.log
