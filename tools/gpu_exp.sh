#!/bin/bash
mkdir -p gpurun_out
python tools/profile_one.py 2 1024 1024 200 336 3 && \
ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 1 -c 1 -f -o gpurun_out/r2_conv3_pair python tools/profile_one.py 2 1024 1024 200 336 3 > gpurun_out/ncu_conv3_pair.log 2>&1
echo "ncu exit $?"
