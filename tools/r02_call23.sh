#!/bin/bash
mkdir -p gpurun_out
{
echo "TMA:"; timeout 300 python tools/attn_only.py 0 4
echo "cp.async:"; B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_cpasync.so timeout 300 python tools/attn_only.py 0 4
} > gpurun_out/r02_attn_only4.log 2>&1; cat gpurun_out/r02_attn_only4.log
