#!/bin/bash
# First-contact GPU run: every kernel case in its own process with a timeout; results -> gpurun_out/kcheck.jsonl
mkdir -p gpurun_out
OUT=gpurun_out/kcheck.jsonl
: > $OUT
run() { timeout 180 python tools/kcheck.py "$@" >> $OUT 2>> gpurun_out/kcheck.err || echo "{\"case\": \"$1\", \"args\": \"$*\", \"ok\": false, \"error\": \"exit $?\"}" >> $OUT; }
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run ln
run gn
run gn c0=384 c1=256 hw=252
run upsample
run linear m=1000 n=256 k=256
run linear m=128 n=128 k=64 bias=0 resid=0 fp32=1
run linear m=4032 n=1152 k=384
run linear m=1024 n=640 k=640 bn=160
run conv
run conv nb=16 h=32 w=2 ci=640 co=640
run conv nb=4 h=63 w=4 ci=384 co=384 resid=1
run conv nb=2 h=125 w=8 ci=256 co=256 stride=2
run conv nb=2 h=50 w=16 ci=8 co=128
run segments
run geglu
run attn variant=0
run attn variant=1
run attn variant=2
run attn variant=3
run attn d=48 s=252 variant=0
run attn d=80 s=64 variant=0
run attn d=32 s=1000 b=16 variant=0 time=1
run unet
cat $OUT
