#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for r in 0 1; do AFIGAN_REUSE_G_FORWARD=$r timeout 300 python tools/quick_time.py bf16 10 2>&1 | tail -1; done
