#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "dpt_tail" 2>&1 | grep -v "^$" | grep "FAILED\|assert\|passed\|failed\|Error" | cut -c1-300 | head -30
