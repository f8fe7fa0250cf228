#!/bin/bash
# One gpurun call: ncu launch lists of the sampling bench and of one training step.
# Every command runs plain first (same arguments) and only then under ncu.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --ddim-steps 3"
$B > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
T="python tools/train_bench.py --eager --steps 1 --warmup 1 --batch 8"
$T > gpurun_out/plain_train.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_train.csv $T > gpurun_out/ncu_train.log 2>&1
ls -la gpurun_out | tail -8
