#!/bin/bash
cd "$(dirname "$0")/.."
python tools/gpu_diag.py "gemm_bf16 or qkv or resid" 2>&1 | grep -E "^(PASS|FAIL)|Error" | head -30
KB=gemm timeout 300 python tools/kbench.py 2>&1 | grep -E "gemm_|cublas" 
for e in 4 6 8; do
  echo "### attention emu pairs = $e"
  export OVG_LIB_PATH=$PWD/omnivggt-official_b200/libovg_e$e.so
  python tools/gpu_diag.py attention 2>&1 | grep -E "^(PASS|FAIL)" | sort | uniq -c | head
  KB=attn timeout 200 python tools/kbench.py 2>&1 | grep -E "^attn"
done
unset OVG_LIB_PATH
echo "### pair probe"
python tools/pair_probe.py 2>&1 | tail -8
ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 40 -c 1 -f -o gpurun_out/gemm2_r01 python tools/pair_probe.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 1 -f -o gpurun_out/gemm1_r01 python tools/pair_probe.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
