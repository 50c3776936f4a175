#!/bin/bash
set -u
mkdir -p gpurun_out
{
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_star8 -s 2 -c 1 -f -o gpurun_out/r02_star8 python tools/profile_fused.py 512 -1 0 2>&1 | tail -3
} > gpurun_out/call_i.log 2>&1
tail -5 gpurun_out/call_i.log
