#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/shape_profile.py > gpurun_out/r02_shape_profile_e.log 2>&1; head -60 gpurun_out/r02_shape_profile_e.log
