#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python tools/stress_tile3t.py > gpurun_out/c6b_stress_tile3t.log 2>&1
B="timeout 300 python bench.py --no_cpu_baseline --extra_configs= --e2e_steps 1"
ODIL_B200_TILE3T=0 $B --config 2 --steps 30 --warmup 3 > gpurun_out/c6b_bench_cfg2_swap_tile3d.json 2> gpurun_out/c6b_1.err
ODIL_B200_TILE3T=0 ODIL_B200_LBFGS_COPY=1 $B --config 2 --steps 30 --warmup 3 > gpurun_out/c6b_bench_cfg2_copy_tile3d.json 2> gpurun_out/c6b_2.err
timeout 300 python -m pytest tests/test_api_gpu.py -q -m gpu -k "graph_replay" > gpurun_out/c6b_tests.log 2>&1
cat gpurun_out/c6b_stress_tile3t.log; tail -n 3 gpurun_out/c6b_tests.log
for f in gpurun_out/c6b_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('final_loss'), d.get('gpu_launches'))
except Exception as e: print('ERR', e)
"; done
