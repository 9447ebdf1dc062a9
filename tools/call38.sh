#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench"; timeout 300 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_final2.json | cut -c1-200
echo "== durations"; timeout 200 python -m pytest tests/test_model_gpu.py -m gpu -q --timeout 120 --durations=6 2>&1 | tail -12
} > gpurun_out/call38.txt 2>&1
tail -30 gpurun_out/call38.txt
