#!/bin/bash
# Call B: the two TMA-fed kernels (k_interp_adjoint3t, k_tile3t): parity, timing against the kernels they replace,
# full GPU suite, headline + configs[2] bench, ncu --set full of one launch each.
set -u
mkdir -p gpurun_out
{
  echo "== new tests"
  timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_zz_wave2_gpu.py -q -m gpu -x -k "adjoint_tma or tile3t or marching" 2>&1 | tail -15
  echo "== adjoint timing"; timeout 120 python tools/time_adjoint.py 256
  echo "== tile3d timing"; timeout 120 python tools/time_tile3d.py 256 512 512
  timeout 120 python tools/time_tile3d.py 128 256 256
  echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
  echo "== headline"; timeout 300 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_n1_b.json | cut -c1-1500
  echo "== config 2"; timeout 300 python bench.py --config 2 2>&1 | tail -1 | tee gpurun_out/bench_config2_b.json | cut -c1-1500
  echo "== ncu"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_interp_adjoint3t -s 3 -c 1 -f -o gpurun_out/r02_adjoint3t python tools/time_adjoint.py 256 tma 2>&1 | tail -3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile3t -s 1 -c 1 -f -o gpurun_out/r02_tile3t python tools/time_tile3d.py 256 512 512 2>&1 | tail -3
} > gpurun_out/call_b.log 2>&1
tail -50 gpurun_out/call_b.log
