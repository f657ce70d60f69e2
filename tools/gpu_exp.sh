#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_modules.py tests/test_gpu_configs.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | tail -3
for f in 0 1 0 1; do AFIGAN_DHEAD_MMA=$f timeout 300 python tools/quick_time.py bf16 10 2>&1 | tail -1; done
