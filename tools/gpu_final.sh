#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-900; }
run all_gpu python -m pytest tests -q -m gpu
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 20 --warmup 3
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
TAILN=34 run stepprof python tools/step_profile.py bf16
AFIGAN_OVERLAP=0 python tools/quick_time.py bf16 1 > gpurun_out/plain_qt.log 2>&1 && \
AFIGAN_OVERLAP=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python tools/quick_time.py bf16 1 > gpurun_out/ncu_qt.log 2>&1
echo "ncu launch list exit $?"
