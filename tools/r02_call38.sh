#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200_FFPROJ=0
run A=new
run B200_FFPROJ=0
run A=new
} > gpurun_out/r02_exp38.log 2>&1
cat gpurun_out/r02_exp38.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -q -m gpu -x > gpurun_out/r02_tests38.log 2>&1; tail -3 gpurun_out/r02_tests38.log
