#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== tile3d timing"; timeout 120 python tools/time_tile3d.py 256 512 512
  timeout 120 python tools/time_tile3d.py 128 256 256
  echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
  echo "== config 2"; timeout 300 python bench.py --config 2 2>&1 | tail -1 | tee gpurun_out/bench_config2_e.json | cut -c1-1800
  echo "== launches config 2"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_config2.csv python bench.py --config 2 --steps 3 --warmup 3 --profile > /dev/null 2>&1
  python tools/launch_summary.py gpurun_out/r02_launches_config2.csv --tail 200 2>&1 | tail -25
} > gpurun_out/call_e.log 2>&1
tail -50 gpurun_out/call_e.log
