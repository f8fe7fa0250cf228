#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200_TILING_MODEL=1
run B200_TILING_MODEL=2
run B200_TILING_MODEL=1
run B200_TILING_MODEL=2
} > gpurun_out/r02_exp31.log 2>&1
cat gpurun_out/r02_exp31.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -q -m gpu -x > gpurun_out/r02_tests31.log 2>&1; tail -3 gpurun_out/r02_tests31.log
