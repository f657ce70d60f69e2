#!/bin/bash
mkdir -p gpurun_out
python tools/quick_time.py bf16 1 > gpurun_out/plain_qt.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/quick_time.py bf16 1 > gpurun_out/ncu_qt.log 2>&1
echo "ncu launches exit $?"; tail -3 gpurun_out/plain_qt.log
