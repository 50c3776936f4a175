#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== interp_add timing"; timeout 120 python tools/time_interp_add.py 256
  echo "== tile3d timing small"; timeout 120 python tools/time_tile3d.py 128 256 256 | head -2
  echo "== headline eager"; timeout 300 python bench.py --no_cpu_baseline --extra_configs "" 2>&1 | tail -1 | tee gpurun_out/bench_n1_f_eager.json | cut -c1-200
  echo "== headline graph"; ODIL_B200_GRAPH=1 timeout 300 python bench.py --no_cpu_baseline --extra_configs "" 2>&1 | tail -1 | tee gpurun_out/bench_n1_f_graph.json | cut -c1-200
  echo "== headline PF"; ODIL_B200_ADD_PF=1 timeout 300 python bench.py --no_cpu_baseline --extra_configs "" 2>&1 | tail -1 | tee gpurun_out/bench_n1_f_pf.json | cut -c1-200
  python - <<'PY'
import json
for f in ["eager","graph","pf"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_n1_f_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()}, d.get("graph_replay"), d["clocks"])
    except Exception as e: print(f, "ERR", e)
PY
} > gpurun_out/call_f.log 2>&1
tail -40 gpurun_out/call_f.log
