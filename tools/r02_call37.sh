#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_only.py 0 16 0 16 4 > gpurun_out/r02_attn_only6.log 2>&1; cat gpurun_out/r02_attn_only6.log
