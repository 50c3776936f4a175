#!/bin/bash
set -u
mkdir -p gpurun_out
{
  for ns in 0 10 20 40 80 160 320; do ODIL_B200_SPIN_NS=$ns timeout 100 python tools/time_star8.py 512; done
  ODIL_B200_SPIN_NS=0 timeout 100 python tools/time_star8.py 512
} > gpurun_out/call_j.log 2>&1
cat gpurun_out/call_j.log
