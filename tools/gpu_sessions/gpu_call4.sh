#!/bin/bash
# Full validation of the build: GPU test-suite, smoke(), the default bench line, configs 2 / 4, launch list.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/c4_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c4_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.log 2>&1
echo "rc=$?" >> gpurun_out/c4_smoke.log
timeout 600 python bench.py > gpurun_out/c4_bench_default.json 2> gpurun_out/c4_bench_default.err
echo "rc=$?" >> gpurun_out/c4_bench_default.err
B="timeout 300 python bench.py --no_cpu_baseline --extra_configs="
$B --config 4 --steps 3 --warmup 3 > gpurun_out/c4_bench_cfg4.json 2> gpurun_out/c4_bench_cfg4.err
ODIL_B200_NEWTON_GATHER=0 $B --config 4 --steps 3 --warmup 3 > gpurun_out/c4_bench_cfg4_scatter.json 2> gpurun_out/c4_bench_cfg4_scatter.err
$B --config 2 --steps 30 --warmup 3 > gpurun_out/c4_bench_cfg2.json 2> gpurun_out/c4_bench_cfg2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c4_launches_bench.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/c4_ncu_bench.log 2>&1
cp gpurun_out/parity_errors.json gpurun_out/c4_parity_errors.json 2>/dev/null
tail -n 5 gpurun_out/c4_tests.log; tail -n 3 gpurun_out/c4_smoke.log
for f in gpurun_out/c4_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('final_loss'), d.get('e2e',{}).get('ms_per_step'), d.get('clocks'), {k:(round(v['ms_per_step'],4), v['calls_per_step']) for k,v in d.get('kernels',{}).items()})
    if d.get('other_configs'): print(json.dumps(d['other_configs'])[:600])
    if d.get('cpu_baseline'): print(d['cpu_baseline'].get('value'), d['cpu_baseline'].get('sample','')[:80])
except Exception as e: print('ERR', e)
"; done
tail -n 3 gpurun_out/c4_bench_default.err
