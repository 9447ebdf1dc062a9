#!/bin/bash
cd "$(dirname "$0")/.."
echo "### TMEM / MUFU microbenchmark"; timeout 120 tools/micro/tmem_bw
python tools/gpu_diag.py "gemm_bf16 or conv3x3" 2>&1 | grep -E "^(PASS|FAIL)|Error|timeout" | grep -v "^PASS" | head
echo "kernel diag: $(grep -c '^PASS' gpurun_out/diag.txt) pass, $(grep -c '^FAIL' gpurun_out/diag.txt) fail"
KB=gemm timeout 300 python tools/kbench.py 2>&1 | grep -E "conv|proj"
python tools/gpu_diag.py "golden or full_width" tests/test_model_gpu.py 2>&1 | grep -E "^(PASS|FAIL)|Error" | head -12
python bench.py 2>&1 | tail -1
OVG_GEMM_PAIR=0 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv python tools/profile_step.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:attn_kernel --profile-from-start off -s 49 -c 1 -f -o gpurun_out/attn_r01b python tools/profile_step.py 2>&1 | tail -1
