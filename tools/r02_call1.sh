#!/bin/bash
# round 2, call 1: GPU test suite (without -x: see every failure), then the bench line with the extra legs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_tests1.log 2>&1; tail -15 gpurun_out/r02_tests1.log
timeout 900 python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; tail -c 3000 gpurun_out/r02_bench1.json; tail -5 gpurun_out/r02_bench1.err
