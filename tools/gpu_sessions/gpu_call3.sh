#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "tile2w" > gpurun_out/c3_tests_tile2w.log 2>&1
echo "rc=$?" >> gpurun_out/c3_tests_tile2w.log
timeout 400 python tools/time_tile2w.py > gpurun_out/c3_time_tile2w.log 2>&1
B="timeout 300 python bench.py --no_cpu_baseline --extra_configs= --steps 30 --warmup 3"
ODIL_B200_TILE2W=0 $B --config 1 > gpurun_out/c3_bench_cfg1_tile2d.json 2> gpurun_out/c3_bench_cfg1_tile2d.err
ODIL_B200_TILE2W=1 $B --config 1 > gpurun_out/c3_bench_cfg1_tile2w.json 2> gpurun_out/c3_bench_cfg1_tile2w.err
ODIL_B200_TILE2W=1 ODIL_B200_T2W_ROWS=8 $B --config 1 > gpurun_out/c3_bench_cfg1_tile2w_r8.json 2> gpurun_out/c3_bench_cfg1_tile2w_r8.err
tail -n 3 gpurun_out/c3_tests_tile2w.log
cat gpurun_out/c3_time_tile2w.log
for f in gpurun_out/c3_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('final_loss'))
except Exception as e: print('ERR', e)
"; done
