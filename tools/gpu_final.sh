#!/bin/bash
# round-end check on one B200 (gpurun --timeout 1800 -- 'bash tools/gpu_final.sh'): tests, smoke, every bench line, profiles
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-900; }
run all_gpu python -m pytest tests -q -m gpu
run smoke python __graft_entry__.py smoke
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench exit $?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm exit $?"
for w in g_only pafpn_c4 stage2_c3 infer_c5; do python bench.py --workload $w > gpurun_out/final_bench_$w.json 2> gpurun_out/final_bench_$w.err; echo "$w exit $?"; done
python bench.py --precision split --no-cpu-baseline > gpurun_out/final_bench_split.json 2> gpurun_out/final_bench_split.err; echo "split exit $?"
AFIGAN_IN_FLIGHT=8 python bench.py --workload infer_c5 > gpurun_out/final_bench_infer_c5_8_in_flight.json 2> /dev/null; echo "infer_c5 x8 exit $?"
TAILN=3 run soak_bf16 python tools/quick_time.py bf16 200
TAILN=24 run gprof_bf16 python tools/g_profile.py bf16
TAILN=34 run stepprof_bf16 python tools/step_profile.py bf16
TAILN=34 run stepprof_split python tools/step_profile.py split
