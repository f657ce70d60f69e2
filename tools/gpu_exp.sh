#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_stage1.py -x -q -m gpu 2>&1 | tail -3
for m in 0 2; do AFIGAN_CONV_HALO=$m timeout 300 python tools/quick_time.py bf16 10 2>&1 | tail -1; done
timeout 300 python tools/step_profile.py bf16 2>&1 | tail -34 | head -14
