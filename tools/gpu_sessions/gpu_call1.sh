#!/bin/bash
# GPU session 1 of this sitting: parity of the new opt-in paths + A/B timings.  Outputs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "tile2w or tile2d" > gpurun_out/c1_tests_tile2w.log 2>&1
echo "rc=$?" >> gpurun_out/c1_tests_tile2w.log
timeout 300 python -m pytest tests/test_api_gpu.py -x -q -m gpu -k "fused_with_synthesis" > gpurun_out/c1_tests_chain.log 2>&1
echo "rc=$?" >> gpurun_out/c1_tests_chain.log
ODIL_B200_S8_ASYNC=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -x -q -m gpu \
    -k "star or poisson or fullsize or graph_replay or trajectory" > gpurun_out/c1_tests_async.log 2>&1
echo "rc=$?" >> gpurun_out/c1_tests_async.log
timeout 300 python tools/time_tile2w.py > gpurun_out/c1_time_tile2w.log 2>&1
for a in 0 1 0 1; do ODIL_B200_S8_ASYNC=$a timeout 100 python tools/time_star8.py 512; done > gpurun_out/c1_time_star8.log 2>&1
B="timeout 300 python bench.py --no_cpu_baseline --extra_configs= --steps 30 --warmup 3"
$B > gpurun_out/c1_bench_base.json 2> gpurun_out/c1_bench_base.err
ODIL_B200_SYNTH_CHAIN=1 $B > gpurun_out/c1_bench_chain.json 2> gpurun_out/c1_bench_chain.err
ODIL_B200_S8_ASYNC=1 $B > gpurun_out/c1_bench_async.json 2> gpurun_out/c1_bench_async.err
ODIL_B200_S8_ASYNC=1 ODIL_B200_SYNTH_CHAIN=1 $B > gpurun_out/c1_bench_both.json 2> gpurun_out/c1_bench_both.err
ODIL_B200_TILE2W=0 $B --config 1 > gpurun_out/c1_bench_cfg1_tile2d.json 2> gpurun_out/c1_bench_cfg1_tile2d.err
ODIL_B200_TILE2W=1 $B --config 1 > gpurun_out/c1_bench_cfg1_tile2w.json 2> gpurun_out/c1_bench_cfg1_tile2w.err
tail -3 gpurun_out/c1_tests_*.log
cat gpurun_out/c1_time_tile2w.log gpurun_out/c1_time_star8.log
for f in gpurun_out/c1_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], {k:round(v['ms_per_step'],4) for k,v in d.get('kernels',{}).items()})
except Exception as e: print('ERR', e)
"; done
