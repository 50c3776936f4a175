#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py tests/test_zz_wave2_gpu.py tests/test_graph_gpu.py -q -m gpu -k "adam or tile2 or lbfgs or graph or trajectory or config1 or wave or optimize" > gpurun_out/c6_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c6_tests.log
B="timeout 300 python bench.py --no_cpu_baseline --extra_configs= --e2e_steps 1"
for i in 1 2; do
$B --config 2 --steps 30 --warmup 3 > gpurun_out/c6_bench_cfg2_swap$i.json 2> gpurun_out/c6_bench_cfg2_swap$i.err
ODIL_B200_LBFGS_COPY=1 $B --config 2 --steps 30 --warmup 3 > gpurun_out/c6_bench_cfg2_copy$i.json 2> gpurun_out/c6_bench_cfg2_copy$i.err
done
$B --config 1 --steps 200 --warmup 5 > gpurun_out/c6_bench_cfg1.json 2> gpurun_out/c6_bench_cfg1.err
tail -n 4 gpurun_out/c6_tests.log
for f in gpurun_out/c6_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('final_loss'), d.get('gpu_launches'), d.get('graph_replay'))
except Exception as e: print('ERR', e)
"; done
