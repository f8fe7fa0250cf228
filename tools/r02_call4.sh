#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "groupnorm or layernorm or conv3x3 or linear" > gpurun_out/r02_tests4.log 2>&1; tail -5 gpurun_out/r02_tests4.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run A=base
run B200_GN_ONEPASS=0
run B200_EMBED_OVERLAP=0
run B200_GN_ONEPASS=0 B200_EMBED_OVERLAP=0
} > gpurun_out/r02_exp4.log 2>&1
cat gpurun_out/r02_exp4.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -q -m gpu > gpurun_out/r02_tests4b.log 2>&1; tail -5 gpurun_out/r02_tests4b.log
