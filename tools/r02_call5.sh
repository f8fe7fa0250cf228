#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/shape_profile.py > gpurun_out/r02_shape_profile_b.log 2>&1; head -2 gpurun_out/r02_shape_profile_b.log
B200_GN_ONEPASS=0 timeout 600 python tools/shape_profile.py > gpurun_out/r02_shape_profile_c.log 2>&1; head -2 gpurun_out/r02_shape_profile_c.log
