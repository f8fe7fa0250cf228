#!/bin/bash
# One gpurun call at the end of a round: full GPU test suite, smoke, the bench line, the fine-tuning step, and the
# ncu launch list of the same bench command (plain first, then under ncu).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 2500 gpurun_out/final_bench.json
timeout 300 python tools/train_bench.py > gpurun_out/final_train.log 2>&1; tail -2 gpurun_out/final_train.log
B="python bench.py --steps 1 --warmup 3 --ddim-steps 3"
timeout 300 $B > gpurun_out/plain_bench.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -6
