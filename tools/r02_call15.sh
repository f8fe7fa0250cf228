#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vocoder.py -q -m gpu -x > gpurun_out/r02_tests15.log 2>&1; tail -5 gpurun_out/r02_tests15.log
{
for ex in "" r v rv; do python tools/gemm_only.py 16000 256 2304 9 0 256 1 50 "$ex"; done
for ex in "" r; do python tools/gemm_only.py 64000 128 1152 9 0 128 1 50 "$ex"; done
for ex in "" r; do python tools/gemm_only.py 1024 640 2560 1 0 64 0 50 "$ex"; done
for ex in "" r; do python tools/gemm_only.py 16000 256 1024 1 0 128 0 50 "$ex"; done
} > gpurun_out/r02_res_ab2.log 2>&1
cat gpurun_out/r02_res_ab2.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run A=base
run B200_RES_TMA=0
} > gpurun_out/r02_exp15.log 2>&1
cat gpurun_out/r02_exp15.log
