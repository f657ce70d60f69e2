#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/quick_time.py bf16 10 | tail -1
AFIGAN_OVERLAP=0 python tools/quick_time.py bf16 1 | tail -1
