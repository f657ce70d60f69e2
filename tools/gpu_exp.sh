#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/quick_time.py bf16 10 2>&1 | tail -1
