#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== tests"; timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -3
echo "== bench"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "== headtail time"; timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:headtail python tools/ncu_kernels.py 2>&1 | grep -A3 "headtail_kernel" | grep "gpu__time" | head -2
} > gpurun_out/call31.txt 2>&1
tail -20 gpurun_out/call31.txt
