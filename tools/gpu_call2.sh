#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_tile2 -c 8 -o gpurun_out/c2_tile2w -f python tools/prof_tile2w.py > gpurun_out/c2_ncu.log 2>&1
ncu -i gpurun_out/c2_tile2w.ncu-rep --page raw --csv > gpurun_out/c2_tile2w_raw.csv 2>> gpurun_out/c2_ncu.log
ncu -i gpurun_out/c2_tile2w.ncu-rep --page source --csv -k regex:k_tile2w > gpurun_out/c2_tile2w_source.csv 2>> gpurun_out/c2_ncu.log
ls -la gpurun_out/ | head -30
tail -5 gpurun_out/c2_ncu.log
