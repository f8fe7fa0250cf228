#!/bin/bash
mkdir -p gpurun_out
B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prof.so B200_GEMM_DEBUG=12 timeout 300 python tools/gemm_timeline.py > gpurun_out/r02_gemm_timeline.log 2>&1
cat gpurun_out/r02_gemm_timeline.log | grep -v "^    *[0-9]*:" 
grep -A14 "lin L3" gpurun_out/r02_gemm_timeline.log | tail -14
