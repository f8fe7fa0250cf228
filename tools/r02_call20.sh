#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "attention" > gpurun_out/r02_tests20.log 2>&1; tail -3 gpurun_out/r02_tests20.log
B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_attnprof.so timeout 300 python tools/attn_timeline.py > gpurun_out/r02_attn_timeline2.log 2>&1
cat gpurun_out/r02_attn_timeline2.log | cut -c1-400
timeout 300 python tools/attn_only.py 0 > gpurun_out/r02_attn_only.log 2>&1; cat gpurun_out/r02_attn_only.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prev.so
run A=new
} > gpurun_out/r02_exp20.log 2>&1
cat gpurun_out/r02_exp20.log
