#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vae.py -q -m gpu -x > gpurun_out/r02_tests12.log 2>&1; tail -5 gpurun_out/r02_tests12.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run A=base
} > gpurun_out/r02_exp12.log 2>&1
cat gpurun_out/r02_exp12.log
timeout 600 python tools/shape_profile.py > gpurun_out/r02_shape_profile_d.log 2>&1; head -30 gpurun_out/r02_shape_profile_d.log
