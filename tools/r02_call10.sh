#!/bin/bash
mkdir -p gpurun_out
{
for bn in 128 256; do for pr in 0 1; do python tools/gemm_only.py 16000 1024 256 1 1 $bn $pr; done; done
for bn in 64 128 192 256; do python tools/gemm_only.py 16000 2048 256 1 0 $bn 0; done
python tools/gemm_only.py 16000 2048 256 1 0 256 1
python tools/gemm_only.py 16000 256 256 1 0 256 0
python tools/gemm_only.py 16000 256 256 1 0 128 0
python tools/gemm_only.py 16000 256 256 1 0 64 0
python tools/gemm_only.py 1024 640 640 1 0 64 0
python tools/gemm_only.py 1024 640 640 1 0 128 0
python tools/gemm_only.py 1024 2560 640 1 1 128 0
python tools/gemm_only.py 4032 1536 384 1 1 256 0
python tools/gemm_only.py 4032 1536 384 1 1 128 0
} > gpurun_out/r02_gemm_ab.log 2>&1
cat gpurun_out/r02_gemm_ab.log
python tools/gemm_only.py 16000 1024 256 1 1 256 0 5 > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:conv_gemm -s 8 -c 1 -o gpurun_out/prof_geglu_l1 python tools/gemm_only.py 16000 1024 256 1 1 256 0 5 > gpurun_out/ncu_geglu.log 2>&1
tail -3 gpurun_out/ncu_geglu.log
