#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python tools/trace_lbfgs.py gpurun_out/c10_trace_swap.txt 20 > gpurun_out/c10_a.log 2>&1
timeout 200 python tools/trace_lbfgs.py gpurun_out/c10_trace_copy.txt 20 > gpurun_out/c10_b.log 2>&1
timeout 200 python tools/trace_lbfgs.py gpurun_out/c10_trace_copy2.txt 20 > gpurun_out/c10_c.log 2>&1
ODIL_B200_TILE3T=0 timeout 200 python tools/trace_lbfgs.py gpurun_out/c10_trace_swap_t3d.txt 20 > gpurun_out/c10_d.log 2>&1
tail -n 2 gpurun_out/c10_?.log
diff gpurun_out/c10_trace_swap.txt gpurun_out/c10_trace_copy.txt | head -12
echo "--- copy vs copy2"; diff gpurun_out/c10_trace_copy.txt gpurun_out/c10_trace_copy2.txt | head -4
