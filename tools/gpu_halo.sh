#!/bin/bash
# halo-tile conv experiment: parity per mode, then step timing per mode
mkdir -p gpurun_out
for m in 1 3; do
  echo "== conv parity, AFIGAN_CONV_HALO=$m"
  AFIGAN_CONV_HALO=$m timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -4
done
echo "== full GPU suite (default mode)"
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for m in 0 1 3; do
  echo "== step time, AFIGAN_CONV_HALO=$m"
  AFIGAN_CONV_HALO=$m timeout 300 python tools/quick_time.py bf16 10 2>&1 | tail -2
done
for m in 0 1; do
  echo "== gemm profile, AFIGAN_CONV_HALO=$m"
  AFIGAN_CONV_HALO=$m timeout 300 python tools/step_profile.py bf16 2>&1 | tail -34
done
