#!/bin/bash
# last validation of the committed state: GPU test-suite, smoke(), default bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/c11_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c11_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c11_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/c11_smoke.log
timeout 400 python bench.py > gpurun_out/c11_bench_default.json 2> gpurun_out/c11_bench_default.err
echo "rc=$?" >> gpurun_out/c11_bench_default.err
tail -n 3 gpurun_out/c11_tests.log; tail -n 2 gpurun_out/c11_smoke.log
python -c "
import json
d=json.loads(open('gpurun_out/c11_bench_default.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('clocks'), {k:(round(v['ms_per_step'],4), v['calls_per_step']) for k,v in d['kernels'].items()})
for k,v in d['other_configs'].items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('gpu_launches'), v.get('error'))
print(d['config']['l2'], d['cpu_baseline'].get('value'))
"
