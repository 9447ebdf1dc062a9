#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== tests (kernels)"; timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -5
echo "== kbench"; KB=gemm timeout 300 python tools/kbench.py 2>&1 | grep "gemm_\|layernorm"
echo "== ncu"; timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r01_block_kernels python tools/ncu_kernels.py 2>&1 | tail -5
} > gpurun_out/call20.txt 2>&1
tail -40 gpurun_out/call20.txt
