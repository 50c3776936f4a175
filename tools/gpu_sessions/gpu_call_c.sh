#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== debug lib f32"; ODIL_B200_LIB=$PWD/odil_b200/lib/libodil_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/debug_tile3t.py f32 2>&1 | tail -25
  echo "== debug lib f64"; ODIL_B200_LIB=$PWD/odil_b200/lib/libodil_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/debug_tile3t.py f64 2>&1 | tail -12
  echo "== debug lib f32 one chunk"; ODIL_B200_LIB=$PWD/odil_b200/lib/libodil_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/debug_tile3t.py f32 7 10 12 7 2>&1 | tail -12
  echo "== debug lib f32 40 18 72"; ODIL_B200_LIB=$PWD/odil_b200/lib/libodil_b200_dbg.so CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/debug_tile3t.py f32 40 18 72 0 2>&1 | tail -12
  echo "== sanitizer f32"; timeout 300 compute-sanitizer --tool memcheck python tools/debug_tile3t.py f32 2>&1 | grep -v "^=========     at\|^=========         in\|Host Frame\|^=========     by" | head -40
} > gpurun_out/call_c.log 2>&1
tail -120 gpurun_out/call_c.log
