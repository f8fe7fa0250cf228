#!/bin/bash
mkdir -p gpurun_out
{
for K in 1152 2304 4608; do for cfg in "256 1" "256 0" "128 1" "128 0"; do python tools/gemm_only.py 16000 256 $K 9 0 $cfg; done; done
for K in 1152 2304; do for cfg in "128 1" "128 0" "64 0"; do python tools/gemm_only.py 64000 128 $K 9 0 $cfg; done; done
for cfg in "128 1" "128 0" "192 0" "64 0"; do python tools/gemm_only.py 4032 384 3456 9 0 $cfg; done
} > gpurun_out/r02_conv_ab.log 2>&1
cat gpurun_out/r02_conv_ab.log
