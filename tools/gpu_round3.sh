#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log; }
run all_gpu python -m pytest tests -q -m gpu
run bench python bench.py --steps 10 --warmup 3
run bench_ref python bench.py --impl reference --steps 1 --warmup 0
run profconv python tools/profile_conv.py bf16 3
python tools/quick_time.py bf16 1 > gpurun_out/plain_qt.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/quick_time.py bf16 1 > gpurun_out/ncu_qt.log 2>&1
echo "ncu launches exit $?"
python tools/profile_conv.py bf16 1 1 > gpurun_out/plain_pc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_conv_tc|k_wgrad_tc" -c 3 -o gpurun_out/prof_r1_conv3 python tools/profile_conv.py bf16 1 1 > gpurun_out/ncu_pc.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
