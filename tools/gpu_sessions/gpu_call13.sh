#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 60 python -m pytest tests/test_api_gpu.py tests/test_zz_wave2_gpu.py -x -q -m gpu -k "lbfgs" > gpurun_out/c13_tests.log 2>&1
echo "rc=$?" >> gpurun_out/c13_tests.log
tail -n 3 gpurun_out/c13_tests.log
