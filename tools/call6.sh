#!/bin/bash
# validate attention v4 (early S issue) + pair-GEMM fix + DINO-on-libovg; fall back to libovg_prev.so if attention fails
cd "$(dirname "$0")/.."
python tools/gpu_diag.py attention 2>&1 | grep -E "^(PASS|FAIL)|timeout|Error" | head -12
if grep -q "^FAIL" gpurun_out/diag.txt; then
  echo "### attention v4 FAILED -> libovg_prev.so"
  export OVG_LIB_PATH=$PWD/omnivggt-official_b200/libovg_prev.so
fi
KB=attn timeout 200 python tools/kbench.py 2>&1 | grep -E "^attn|sdpa"
python tools/gpu_diag.py "gemm or conv or layernorm or qkv or resid" 2>&1 | grep -E "^(PASS|FAIL)|Error|timeout" | grep -v "^PASS" | head -30
echo "kernel diag done: $(grep -c '^PASS' gpurun_out/diag.txt) pass, $(grep -c '^FAIL' gpurun_out/diag.txt) fail"
python tools/pair_probe.py 2>&1 | tail -6
KB=gemm timeout 300 python tools/kbench.py 2>&1 | grep -E "bn512|bn256|cublas|conv"
python tools/gpu_diag.py tests/test_model_gpu.py 2>&1 | grep -E "^(PASS|FAIL)|^conv_|^dino_|^full_width|Error" | head -40
cp gpurun_out/diag.txt gpurun_out/diag_model.txt
python bench.py --no-cpu-baseline 2>&1 | tail -1
