#!/bin/bash
# ncu --set full evidence for the kernels added in this sitting: k_tile2w (final version) and the stored-diagonal Newton products
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 ncu --set full --clock-control none -k regex:k_tile2 -c 8 -o gpurun_out/c9_tile2w -f python tools/prof_tile2w.py > gpurun_out/c9_ncu_tile2w.log 2>&1
ncu -i gpurun_out/c9_tile2w.ncu-rep --page raw --csv > gpurun_out/c9_tile2w_raw.csv 2>> gpurun_out/c9_ncu_tile2w.log
rm -f gpurun_out/c9_tile2w.ncu-rep
timeout 300 ncu --set full --clock-control none -k regex:k_g0_ --launch-skip 6 -c 6 -o gpurun_out/c9_newton -f python bench.py --profile --config 4 --steps 1 --warmup 1 > gpurun_out/c9_ncu_newton.log 2>&1
ncu -i gpurun_out/c9_newton.ncu-rep --page raw --csv > gpurun_out/c9_newton_raw.csv 2>> gpurun_out/c9_ncu_newton.log
rm -f gpurun_out/c9_newton.ncu-rep
ls -la gpurun_out/c9_*; tail -n 3 gpurun_out/c9_ncu_tile2w.log gpurun_out/c9_ncu_newton.log
