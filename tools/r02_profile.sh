#!/bin/bash
# round-2 ncu evidence: launch list of the bench command (plain run first), then ncu --set full of three launches:
# the level-0 3x3 convolution, the level-1 GEGLU GEMM and the level-1 attention
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --ddim-steps 3 --legs c2"
timeout 300 $B > gpurun_out/r02_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log
python tools/ncu_step.py gpurun_out/r02_launches_bench.csv > gpurun_out/r02_ncu_step.txt 2>&1; head -14 gpurun_out/r02_ncu_step.txt
python tools/conv_only.py 1 128 > gpurun_out/r02_conv_plain.log 2>&1 && \
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -s 8 -c 1 -o gpurun_out/r02_prof_conv_l0 python tools/conv_only.py 1 128 > gpurun_out/r02_ncu_conv.log 2>&1
python tools/gemm_only.py 16000 1024 256 1 1 256 0 5 > gpurun_out/r02_geglu_plain.log 2>&1 && \
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -s 8 -c 1 -o gpurun_out/r02_prof_geglu_l1_after python tools/gemm_only.py 16000 1024 256 1 1 256 0 5 > gpurun_out/r02_ncu_geglu2.log 2>&1
python tools/attn_only.py 0 > gpurun_out/r02_attn_plain.log 2>&1 && \
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention_kernel -s 8 -c 1 -o gpurun_out/r02_prof_attn_l1 python tools/attn_only.py 0 > gpurun_out/r02_ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
