#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "attention" > gpurun_out/r02_tests30.log 2>&1; tail -3 gpurun_out/r02_tests30.log
timeout 300 python tools/attn_only.py 0 > gpurun_out/r02_attn_only5.log 2>&1; cat gpurun_out/r02_attn_only5.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run A=new
} > gpurun_out/r02_exp30.log 2>&1
cat gpurun_out/r02_exp30.log
