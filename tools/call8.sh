#!/bin/bash
cd "$(dirname "$0")/.."
python tools/gpu_diag.py "attention or upsample" 2>&1 | grep -E "^(PASS|FAIL)|timeout|Error" | head -14
if grep -q "^FAIL.*attention" gpurun_out/diag.txt; then
  echo "### attention v5 FAILED -> libovg_prev.so"; export OVG_LIB_PATH=$PWD/omnivggt-official_b200/libovg_prev.so
fi
KB=attn timeout 200 python tools/kbench.py 2>&1 | grep -E "^attn|sdpa"
python tools/gpu_diag.py "golden or full_width or determinism" tests/test_model_gpu.py 2>&1 | grep -E "^(PASS|FAIL)|Error" | head -14
python bench.py --no-cpu-baseline 2>&1 | tail -1
