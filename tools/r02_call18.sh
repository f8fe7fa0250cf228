#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vocoder.py tests/test_gpu_vae.py -q -m gpu -x > gpurun_out/r02_tests18.log 2>&1; tail -3 gpurun_out/r02_tests18.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prev.so
run A=new
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prev.so
run A=new
} > gpurun_out/r02_exp18.log 2>&1
cat gpurun_out/r02_exp18.log
