#!/bin/bash
mkdir -p gpurun_out
{
for m in 0 1 2; do for ex in "" v; do echo "EPI_VEC=$m"; B200_EPI_VEC=$m python tools/gemm_only.py 16000 256 2304 9 0 256 1 50 "$ex"; done; done
for m in 0 1 2; do for ex in "" v; do echo "EPI_VEC=$m"; B200_EPI_VEC=$m python tools/gemm_only.py 64000 128 1152 9 0 128 1 50 "$ex"; done; done
} > gpurun_out/r02_vec_ab.log 2>&1
cat gpurun_out/r02_vec_ab.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run B200_EPI_VEC=0
run B200_EPI_VEC=1
run B200_EPI_VEC=2
run B200_EPI_VEC=0
} > gpurun_out/r02_exp16.log 2>&1
cat gpurun_out/r02_exp16.log
