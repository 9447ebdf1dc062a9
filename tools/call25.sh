#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== repro"; timeout 120 python tools/qkv_repro.py 2>&1 | tail -5
echo "== tests"; timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -8
echo "== kbench"; timeout 200 python tools/kbench.py 2>&1 | grep "qkv\|attn\|sdpa"
echo "== attn single"; OVG_ATTN_SINGLE=1 KB=attn timeout 120 python tools/kbench.py 2>&1 | grep "^attn\|rror"
echo "== attn single tests"; OVG_ATTN_SINGLE=1 timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 60 -k attention 2>&1 | tail -4
echo "== bench"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1
echo "== bench single"; OVG_ATTN_SINGLE=1 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
} > gpurun_out/call25.txt 2>&1
tail -60 gpurun_out/call25.txt
