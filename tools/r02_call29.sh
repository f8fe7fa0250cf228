#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run A=base
run B200_GN_ONEPASS=1
run B200_LN_FUSED=1
run B200_LN_FUSED=1 B200_LN_FUSED_QKV_MT=8 B200_LN_FUSED_FF_MT=40
run B200_LN_FUSED=1 B200_LN_FUSED_QKV_MT=40 B200_LN_FUSED_FF_MT=40
run B200_LN_FUSED=1 B200_LN_FUSED_QKV_MT=0 B200_LN_FUSED_FF_MT=-1
run B200_LORA_FUSED=0
} > gpurun_out/r02_exp29.log 2>&1
cat gpurun_out/r02_exp29.log
