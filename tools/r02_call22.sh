#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "attention" > gpurun_out/r02_tests22.log 2>&1; tail -3 gpurun_out/r02_tests22.log
B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_attnprof.so timeout 300 python tools/attn_timeline.py 2>&1 | grep -v "^attn.*-[0-9]\{10\}" > gpurun_out/r02_attn_timeline3.log
cut -c1-330 gpurun_out/r02_attn_timeline3.log
timeout 300 python tools/attn_only.py 0 > gpurun_out/r02_attn_only3.log 2>&1; cat gpurun_out/r02_attn_only3.log
B200_ATTN_STAGES=2 timeout 300 python tools/attn_only.py 0 >> gpurun_out/r02_attn_only3.log 2>&1; tail -1 gpurun_out/r02_attn_only3.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prev.so
run A=new
run B200_ATTN_STAGES=2
} > gpurun_out/r02_exp22.log 2>&1
cat gpurun_out/r02_exp22.log
