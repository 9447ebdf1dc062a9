#!/bin/bash
cd "$(dirname "$0")/.."
python tools/gpu_diag.py "attention" 2>&1 | grep -E "^(PASS|FAIL)|timeout|Error" | head -10
KB=attn timeout 200 python tools/kbench.py 2>&1 | grep -E "^attn"
python tools/gpu_diag.py "golden or full_width or determinism or graph" tests/test_model_gpu.py 2>&1 | grep -E "^(PASS|FAIL)|Error|Warn" | head -16
python bench.py --no-cpu-baseline 2>&1 | tail -1
OVG_CUDA_GRAPH=0 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220
