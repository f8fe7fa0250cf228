#!/bin/bash
# 2-GPU bench (torchrun, as the driver launches it): headline + legs c3 / c4 (NCCL all-reduce) / c5
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 2500 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
