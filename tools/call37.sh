#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== tests"; timeout 400 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -4
echo "== bench rows-upsample"; timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "== bench direct-upsample"; OVG_UPSAMPLE_ROWS=0 timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "== upsample kernel times"; timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:upsample python tools/ncu_kernels.py 2>&1 | grep -B2 "gpu__time_duration" | grep "upsample\|gpu__time" | head -4
} > gpurun_out/call37.txt 2>&1
tail -30 gpurun_out/call37.txt
