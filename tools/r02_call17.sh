#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prev.so
run A=new
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prev.so
run A=new
run B200_RES_TMA=0 B200_EPI_VEC=0
} > gpurun_out/r02_exp17.log 2>&1
cat gpurun_out/r02_exp17.log
