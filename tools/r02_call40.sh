#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/c5_profile.py > gpurun_out/r02_c5_profile.log 2>&1; head -40 gpurun_out/r02_c5_profile.log
