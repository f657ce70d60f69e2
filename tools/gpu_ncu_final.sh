#!/bin/bash
# ncu evidence of the final build: launch list of one bf16 step and one --set full capture of the dominant kernel (run after the plain programs exit 0)
mkdir -p gpurun_out
python tools/quick_time.py bf16 1 > gpurun_out/plain_qt.log 2>&1 && \
AFIGAN_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16.csv python tools/quick_time.py bf16 1 > gpurun_out/ncu_qt.log 2>&1
echo "launch list exit $?"
python tools/profile_one.py 2 1024 1024 200 336 3 bf16 > gpurun_out/plain_p1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_conv_halo -s 1 -c 1 -o gpurun_out/r02_conv3_tb3 -f python tools/profile_one.py 2 1024 1024 200 336 3 bf16 > gpurun_out/ncu_p1.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/r02_conv3_tb3.ncu-rep --page raw --csv > gpurun_out/r02_conv3_tb3_raw.csv 2>/dev/null
tail -2 gpurun_out/plain_p1.log
