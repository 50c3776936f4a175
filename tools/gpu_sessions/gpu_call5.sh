#!/bin/bash
# configs[1] launch merging: 2-D adam_synth, table_pick, reduction fused into k_tile2w; L-BFGS changes
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -x -q -m gpu -k "adam or tile2 or lbfgs or graph or trajectory or config1 or wave or optimize" > gpurun_out/c5_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c5_tests.log
B="timeout 300 python bench.py --no_cpu_baseline --extra_configs="
$B --config 1 --steps 200 --warmup 5 > gpurun_out/c5_bench_cfg1.json 2> gpurun_out/c5_bench_cfg1.err
ODIL_B200_FUSE_SYNTH=0 $B --config 1 --steps 200 --warmup 5 > gpurun_out/c5_bench_cfg1_nosynth.json 2> gpurun_out/c5_bench_cfg1_nosynth.err
$B --config 2 --steps 30 --warmup 3 > gpurun_out/c5_bench_cfg2.json 2> gpurun_out/c5_bench_cfg2.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c5_launches_cfg1.csv python bench.py --profile --config 1 --steps 4 --warmup 3 --graph 0 > gpurun_out/c5_ncu_cfg1.log 2>&1
tail -n 4 gpurun_out/c5_tests.log
for f in gpurun_out/c5_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('final_loss'), d.get('gpu_launches'), d.get('graph_replay'))
except Exception as e: print('ERR', e)
"; done
python tools/launch_summary.py gpurun_out/c5_launches_cfg1.csv --tail 40 | tail -14
