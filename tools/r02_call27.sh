#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vocoder.py tests/test_gpu_vae.py -q -m gpu -x > gpurun_out/r02_tests27.log 2>&1; tail -5 gpurun_out/r02_tests27.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prev.so
run A=new
} > gpurun_out/r02_exp27.log 2>&1
cat gpurun_out/r02_exp27.log
B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prof.so B200_GEMM_DEBUG=12 timeout 300 python tools/gemm_timeline.py > gpurun_out/r02_gemm_timeline2.log 2>&1
grep -A16 "lin L3" gpurun_out/r02_gemm_timeline2.log | head -17
