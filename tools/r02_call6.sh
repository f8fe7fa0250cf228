#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vae.py -q -m gpu -x > gpurun_out/r02_tests6.log 2>&1; tail -30 gpurun_out/r02_tests6.log
run() { echo "== $*"; env "$@" timeout 300 python tools/step_time.py 1 2>&1 | tail -1; }
run A=base
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tensor_maps or layernorm" 2>&1 | tail -3
