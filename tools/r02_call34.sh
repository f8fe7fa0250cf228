#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -q -m gpu -x -s -k "bf16_storage or match_oracle" > gpurun_out/r02_tests34.log 2>&1; tail -8 gpurun_out/r02_tests34.log
cat gpurun_out/train_parity_report.json
