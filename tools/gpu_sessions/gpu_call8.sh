#!/bin/bash
# Final validation of the build: GPU test-suite, smoke(), the default bench line (other configs embedded), launch list of configs[1]
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c8_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c8_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c8_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/c8_smoke.log
timeout 600 python bench.py > gpurun_out/c8_bench_default.json 2> gpurun_out/c8_bench_default.err
echo "rc=$?" >> gpurun_out/c8_bench_default.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/c8_launches_cfg1.csv python bench.py --profile --config 1 --steps 4 --warmup 3 --graph 0 > gpurun_out/c8_ncu_cfg1.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/c8_parity_errors.json 2>/dev/null
tail -n 5 gpurun_out/c8_tests.log; tail -n 3 gpurun_out/c8_smoke.log
python -c "
import json
d=json.loads(open('gpurun_out/c8_bench_default.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d.get('final_loss'), d['e2e']['ms_per_step'], d.get('clocks'), {k:(round(v['ms_per_step'],4), v['calls_per_step']) for k,v in d['kernels'].items()})
for k,v in d['other_configs'].items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('gpu_launches'), v.get('final_loss'), v.get('error'))
print(d['cpu_baseline'].get('value'), d['cpu_baseline'].get('cores'))
"
python tools/launch_summary.py gpurun_out/c8_launches_cfg1.csv --tail 28 | tail -10
tail -n 2 gpurun_out/c8_bench_default.err
