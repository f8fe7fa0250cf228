#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "groupnorm or geglu or unet_forward" > gpurun_out/r02_tests11.log 2>&1; tail -5 gpurun_out/r02_tests11.log
{
python tools/gemm_only.py 16000 1024 256 1 1 256 0
python tools/gemm_only.py 1024 2560 640 1 1 128 0
python tools/gemm_only.py 4032 1536 384 1 1 256 0
} > gpurun_out/r02_geglu_after.log 2>&1
cat gpurun_out/r02_geglu_after.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run A=base
run B200_GN_RESIDENT=0
run A=base
} > gpurun_out/r02_exp11.log 2>&1
cat gpurun_out/r02_exp11.log
