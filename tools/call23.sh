#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== kbench"; KB=gemm timeout 300 python tools/kbench.py 2>&1 | grep "qkv"
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1
} > gpurun_out/call23.txt 2>&1
tail -40 gpurun_out/call23.txt
