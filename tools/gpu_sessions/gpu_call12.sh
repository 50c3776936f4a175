#!/bin/bash
# capture_guard: graph-replay tests + configs[1]
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 python -m pytest tests/test_api_gpu.py -x -q -m gpu -k "graph_replay or fused_with_synthesis or config1" > gpurun_out/c12_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c12_tests.log
timeout 60 python bench.py --no_cpu_baseline --extra_configs= --config 1 --steps 200 --warmup 5 --e2e_steps 2 > gpurun_out/c12_bench_cfg1.json 2> gpurun_out/c12_bench_cfg1.err
tail -n 3 gpurun_out/c12_tests.log
python -c "
import json
d=json.loads(open('gpurun_out/c12_bench_cfg1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('gpu_launches'), d.get('graph_replay'))
"
