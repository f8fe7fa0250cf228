#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vocoder.py -q -m gpu -x > gpurun_out/r02_tests36.log 2>&1; tail -5 gpurun_out/r02_tests36.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
