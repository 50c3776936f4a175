#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/time_tile2w.py > gpurun_out/c2_time_tile2w.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_tile2 -c 8 -o gpurun_out/c2_tile2w -f python tools/prof_tile2w.py > gpurun_out/c2_ncu.log 2>&1
ncu -i gpurun_out/c2_tile2w.ncu-rep --page raw --csv > gpurun_out/c2_tile2w_raw.csv 2>> gpurun_out/c2_ncu.log
ncu -i gpurun_out/c2_tile2w.ncu-rep --page source --csv -k regex:k_tile2w > gpurun_out/c2_tile2w_source.csv 2>> gpurun_out/c2_ncu.log
rm -f gpurun_out/c2_tile2w.ncu-rep
ODIL_B200_S8_ASYNC=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -x -q -m gpu \
    -k "star or poisson or fullsize or graph_replay or trajectory" > gpurun_out/c2_tests_async.log 2>&1
echo "rc=$?" >> gpurun_out/c2_tests_async.log
timeout 600 python -m pytest tests/test_graph_gpu.py tests/test_newton_gpu.py tests/test_kernels_gpu.py -x -q -m gpu -k "linearize or newton or tile2w" > gpurun_out/c2_tests_dia.log 2>&1
echo "rc=$?" >> gpurun_out/c2_tests_dia.log
B="timeout 300 python bench.py --no_cpu_baseline --extra_configs= --steps 3 --warmup 3 --config 4"
ODIL_B200_NEWTON_DIA=0 $B > gpurun_out/c2_bench_cfg4_nodia.json 2> gpurun_out/c2_bench_cfg4_nodia.err
ODIL_B200_NEWTON_DIA=1 $B > gpurun_out/c2_bench_cfg4_dia.json 2> gpurun_out/c2_bench_cfg4_dia.err
cat gpurun_out/c2_time_tile2w.log
for f in gpurun_out/c2_tests_*.log; do echo == $f; tail -n 3 $f; done
for f in gpurun_out/c2_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('final_loss'), {k:(round(v['ms_per_step'],4), v['calls_per_step']) for k,v in d.get('kernels',{}).items()})
except Exception as e: print('ERR', e)
"; done
