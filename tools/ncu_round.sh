#!/bin/bash
# One gpurun call: ncu launch lists (sampling step, training step) + --set full captures of the top kernels.
# Every command runs plain first (same arguments) and only then under ncu.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --ddim-steps 3"
$B > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
T="python tools/train_bench.py --eager --steps 1 --warmup 1 --batch 8"
$T > gpurun_out/plain_train.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv $T > gpurun_out/ncu_train.log 2>&1
S="python tools/step_time.py 1"
$S > gpurun_out/plain_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 1300 -c 4 -o gpurun_out/prof_gemm $S > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 40 -c 2 -o gpurun_out/prof_attn $S > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_bwd_kernel -s 4 -c 2 -o gpurun_out/prof_attn_bwd $T > gpurun_out/ncu_attn_bwd.log 2>&1
ls -la gpurun_out | tail -20
