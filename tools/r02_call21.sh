#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_only.py 0 8 0 8 > gpurun_out/r02_attn_only2.log 2>&1; cat gpurun_out/r02_attn_only2.log
