#!/bin/bash
# first GPU pass: each group in its own process (a trap in the tcgen05 kernel must not poison the others)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run conv_simt python -m pytest tests/test_gpu_conv.py -q -x -k "fp32 or bf16_simt or strided"
run modules_simt python -m pytest tests/test_gpu_modules.py -q -s -k "fp32 or bf16_simt or zero or cpu_tensor"
run stage1_fp32 python -m pytest tests/test_gpu_stage1.py -q -s -k "fp32"
run conv_tc_fwd python -m pytest tests/test_gpu_conv.py -q -k "forward and bf16 and not simt"
run conv_tc_bwd python -m pytest tests/test_gpu_conv.py -q -k "backward and bf16 and not simt"
run time_fp32 python tools/quick_time.py fp32 1
