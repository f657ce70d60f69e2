#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for f in 0 1 0 1; do AFIGAN_WGRAD_PAIR=$f timeout 300 python tools/quick_time.py bf16 10 2>&1 | tail -1; done
timeout 300 python tools/step_profile.py bf16 2>&1 | grep wgrad
