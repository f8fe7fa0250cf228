#!/bin/bash
mkdir -p gpurun_out
B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_attnprof.so timeout 300 python tools/attn_timeline.py > gpurun_out/r02_attn_timeline.log 2>&1
cat gpurun_out/r02_attn_timeline.log
