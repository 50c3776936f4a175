#!/bin/bash
set -u
mkdir -p gpurun_out
{
  echo "== timing occ 3"; timeout 120 python tools/time_adam_synth.py
  echo "== timing occ 2"; ODIL_B200_SYNTH_OCC=2 timeout 120 python tools/time_adam_synth.py
  echo "== new tests"
  timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_api_gpu.py -q -m gpu -x -k "adam_synth or fused_with_synthesis or graph_replay or fused_into" 2>&1 | tail -4
  echo "== headline fused synth"; timeout 300 python bench.py --no_cpu_baseline --extra_configs "" 2>&1 | tail -1 | tee gpurun_out/bench_n1_h_synth.json | cut -c1-200
  echo "== headline unfused"; ODIL_B200_FUSE_SYNTH=0 timeout 300 python bench.py --no_cpu_baseline --extra_configs "" 2>&1 | tail -1 | tee gpurun_out/bench_n1_h_unfused.json | cut -c1-200
  python - <<'PY'
import json
for f in ["synth","unfused"]:
    try:
        d=json.loads(open(f"gpurun_out/bench_n1_h_{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], {k:(round(v["ms_per_step"],4), round(v.get("frac",0),3)) for k,v in d["kernels"].items()}, d.get("graph_replay"), d["clocks"], d["final_loss"], d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
  echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
} > gpurun_out/call_h.log 2>&1
tail -40 gpurun_out/call_h.log
