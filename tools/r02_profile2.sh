#!/bin/bash
# round-2 (final kernels) ncu evidence: launch list of the bench command (plain run first), then ncu --set full of the
# level-1 attention launch (pipelined, P through TMEM) and of the level-0 3x3 convolution
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --ddim-steps 3 --legs c2"
timeout 300 $B > gpurun_out/r02_plain_bench_v4.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_bench_v4.csv $B > gpurun_out/r02_ncu_bench_v4.log 2>&1
tail -2 gpurun_out/r02_ncu_bench_v4.log
python tools/ncu_step.py gpurun_out/r02_launches_bench_v4.csv > gpurun_out/r02_ncu_step_v4.txt 2>&1; head -16 gpurun_out/r02_ncu_step_v4.txt
python tools/attn_only.py 0 > gpurun_out/r02_attn_plain_v4.log 2>&1 && \
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention_kernel -s 8 -c 1 -o gpurun_out/r02_prof_attn_l1_ts python tools/attn_only.py 0 > gpurun_out/r02_ncu_attn_v4.log 2>&1
python tools/conv_only.py 1 128 > gpurun_out/r02_conv_plain_v4.log 2>&1 && \
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -s 8 -c 1 -o gpurun_out/r02_prof_conv_l0_v4 python tools/conv_only.py 1 128 > gpurun_out/r02_ncu_conv_v4.log 2>&1
ls -la gpurun_out/*.ncu-rep
