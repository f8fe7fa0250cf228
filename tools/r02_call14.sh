#!/bin/bash
mkdir -p gpurun_out
{
for ex in "" r v rv; do python tools/gemm_only.py 16000 256 2304 9 0 256 1 50 "$ex"; done
for ex in "" r; do python tools/gemm_only.py 16000 256 2304 9 0 128 0 50 "$ex"; done
for ex in "" r; do python tools/gemm_only.py 64000 128 1152 9 0 128 1 50 "$ex"; done
for ex in "" r; do python tools/gemm_only.py 1024 640 2560 1 0 64 0 50 "$ex"; done
for ex in "" r; do python tools/gemm_only.py 16000 256 1024 1 0 256 0 50 "$ex"; done
for ex in "" r; do python tools/gemm_only.py 16000 256 1024 1 0 128 0 50 "$ex"; done
} > gpurun_out/r02_res_ab.log 2>&1
cat gpurun_out/r02_res_ab.log
