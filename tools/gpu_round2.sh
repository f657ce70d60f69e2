#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 30 gpurun_out/$name.log; }
run all_gpu python -m pytest tests -q -s -m gpu
run time_bf16 python tools/quick_time.py bf16 3
run time_bf16_simt python tools/quick_time.py bf16_simt 1
run smoke python __graft_entry__.py smoke
