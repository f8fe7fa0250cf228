#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_head.so
run A=new
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_head.so
run A=new
} > gpurun_out/r02_exp32.log 2>&1
cat gpurun_out/r02_exp32.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "conv or linear or geglu or unet" > gpurun_out/r02_tests32.log 2>&1; tail -3 gpurun_out/r02_tests32.log
