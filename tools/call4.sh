#!/bin/bash
# A/B driver for a GPU box: validate the new attention + CTA-pair GEMM first; fall back to the v1 library if they fail.
cd "$(dirname "$0")/.."
python tools/gpu_diag.py attention 2>&1 | tail -12
if grep -q "^FAIL" gpurun_out/diag.txt; then
  echo "### attention v2 FAILED -> falling back to libovg_v1.so for the rest"
  export OVG_LIB_PATH=$PWD/omnivggt-official_b200/libovg_v1.so
  python tools/gpu_diag.py tests/test_model_gpu.py 2>&1 | grep -E "^(PASS|FAIL)|^conv_|^dino_|^full_width"
  exit 0
fi
cp gpurun_out/diag.txt gpurun_out/diag_attn.txt
python tools/gpu_diag.py "gemm or conv" 2>&1 | grep -E "^(PASS|FAIL)|Error|timeout" | head -60
cp gpurun_out/diag.txt gpurun_out/diag_gemm.txt
timeout 600 python tools/kbench.py 2>&1 | tail -40
python tools/gpu_diag.py tests/test_model_gpu.py 2>&1 | grep -E "^(PASS|FAIL)|^conv_|^dino_|^full_width|Error" | head -40
cp gpurun_out/diag.txt gpurun_out/diag_model.txt
python bench.py --no-cpu-baseline 2>&1 | tail -1
OVG_GEMM_PAIR=1 python bench.py --no-cpu-baseline 2>&1 | tail -1
