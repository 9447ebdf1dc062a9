#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== tests"; timeout 500 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -8
echo "== kbench split"; KB=gemm timeout 200 python tools/kbench.py 2>&1 | grep "gemm_"
echo "== kbench nosplit"; OVG_GEMM_SPLIT=0 KB=gemm timeout 200 python tools/kbench.py 2>&1 | grep "gemm_proj\|gemm_fc"
echo "== bench"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
} > gpurun_out/call27.txt 2>&1
tail -60 gpurun_out/call27.txt
