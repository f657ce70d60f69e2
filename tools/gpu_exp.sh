#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/quick_time.py bf16 10 2>&1 | tail -1
timeout 300 python tools/step_profile.py bf16 2>&1 | tail -30 | grep "conv_tc  *\(32\|128\)"
