#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== full-size config tests"; timeout 400 python -m pytest tests/test_model_gpu.py -m gpu -x -q --timeout 180 -k full_size 2>&1 | tail -5
echo "== bench S=4"; timeout 200 python bench.py --views 4 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2000
echo "== bench S=24"; timeout 300 python bench.py --views 24 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2000
} > gpurun_out/call32.txt 2>&1
tail -20 gpurun_out/call32.txt | cut -c1-400
