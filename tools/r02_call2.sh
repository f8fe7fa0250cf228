#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -q -m gpu -k "logmel or rank16 or groupnorm or free_running" > gpurun_out/r02_tests2.log 2>&1; tail -5 gpurun_out/r02_tests2.log
V=audioldm_with_lora_b200/variants
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run A=base
run B200_NO_PDL=1
run B200_GEMM_SMEM_KB=180
run B200LDM_LIB=$V/libb200ldm_r128.so
run B200LDM_LIB=$V/libb200ldm_r128.so B200_GEMM_SMEM_KB=180
run B200LDM_LIB=$V/libb200ldm_r128.so B200_GEMM_SMEM_KB=150
run B200LDM_LIB=$V/libb200ldm_r128.so B200_GEMM_SMEM_KB=120
} > gpurun_out/r02_exp2.log 2>&1
cat gpurun_out/r02_exp2.log
