#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== tile3t tests"
  timeout 600 python -m pytest tests/test_zz_wave2_gpu.py -q -m gpu -x 2>&1 | tail -8
  echo "== tile3d timing"; timeout 120 python tools/time_tile3d.py 256 512 512
  timeout 120 python tools/time_tile3d.py 128 256 256
  echo "== ncu"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile3t -s 1 -c 1 -f -o gpurun_out/r02_tile3t python tools/time_tile3d.py 256 512 512 2>&1 | tail -3
} > gpurun_out/call_d.log 2>&1
tail -50 gpurun_out/call_d.log
