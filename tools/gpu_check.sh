#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run all_gpu python -m pytest tests -q -m gpu -x
run time_bf16 python tools/quick_time.py bf16 5
TAILN=60 run stepprof python tools/step_profile.py bf16
