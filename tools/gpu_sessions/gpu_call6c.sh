#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in memcheck racecheck initcheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/stress_tile3t.py 20 48 128 > gpurun_out/c6c_tile3t_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/c6c_tile3t_$tool.log
done
# the L-BFGS path of configs[2] at a small size under initcheck / memcheck
for tool in memcheck initcheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python bench.py --profile --config 2 --size 64 --steps 4 --warmup 2 > gpurun_out/c6c_lbfgs_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/c6c_lbfgs_$tool.log
done
for f in gpurun_out/c6c_*.log; do echo == $f; grep -c "=========" $f; grep "ERROR SUMMARY\|Invalid\|Uninitialized\|hazard\|Race\|rc=" $f | head -8; tail -n 2 $f; done
