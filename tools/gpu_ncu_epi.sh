#!/bin/bash
mkdir -p gpurun_out
python tools/quick_time.py bf16 1 > gpurun_out/plain_qt.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 22 -c 5 -o gpurun_out/prof_epi python tools/quick_time.py bf16 1 > gpurun_out/ncu_epi.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_epi.log
