#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu -x > gpurun_out/r02_tests25.log 2>&1; tail -5 gpurun_out/r02_tests25.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
{
run A=new
run B200_CTA_PAIR=0
run B200_PAIR_MIN_BN=256
run A=new
} > gpurun_out/r02_exp25.log 2>&1
cat gpurun_out/r02_exp25.log
