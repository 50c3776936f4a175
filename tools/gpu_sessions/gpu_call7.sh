#!/bin/bash
# 2-GPU sanity of the final build: slab tests + the N=2 bench line (weak + strong block + parity block)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_slab_gpu.py -q -m gpu > gpurun_out/c7_tests_slab.log 2>&1
echo "rc=$?" >> gpurun_out/c7_tests_slab.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c7_bench_n2.json 2> gpurun_out/c7_bench_n2.err
echo "rc=$?" >> gpurun_out/c7_bench_n2.err
tail -n 5 gpurun_out/c7_tests_slab.log; tail -n 3 gpurun_out/c7_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/c7_bench_n2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d.get('parity'), d.get('strong'), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()}, d['e2e'])
"
