#!/bin/bash
cd "$(dirname "$0")/.."
python tools/gpu_diag.py "attention" 2>&1 | grep -E "^(PASS|FAIL)|timeout|Error" | head -10
if grep -q "^FAIL" gpurun_out/diag.txt; then
  echo "### attention v6 FAILED -> libovg_prev.so (no staged epilogue either)"; export OVG_LIB_PATH=$PWD/omnivggt-official_b200/libovg_prev.so
fi
KB=attn timeout 200 python tools/kbench.py 2>&1 | grep -E "^attn"
python tools/gpu_diag.py "gemm or conv or resid" 2>&1 | grep -E "^(PASS|FAIL)|Error|timeout" | grep -v "^PASS" | head -20
echo "gemm diag: $(grep -c '^PASS' gpurun_out/diag.txt) pass, $(grep -c '^FAIL' gpurun_out/diag.txt) fail"
if grep -q "^FAIL" gpurun_out/diag.txt; then echo "### staged epilogue FAILED -> OVG_GEMM_STAGE=0"; export OVG_GEMM_STAGE=0; fi
KB=gemm timeout 300 python tools/kbench.py 2>&1 | grep -E "bn512|bn384|cublas"
OVG_GEMM_STAGE=0 KB=gemm timeout 300 python tools/kbench.py 2>&1 | grep -E "bn512" | sed 's/^/nostage /'
python tools/gpu_diag.py "golden or full_width or determinism" tests/test_model_gpu.py 2>&1 | grep -E "^(PASS|FAIL)|Error" | head -14
python bench.py --no-cpu-baseline 2>&1 | tail -1
