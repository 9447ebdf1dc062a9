#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== tests attention (half)"; OVG_ATTN_SINGLE=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 60 -k attention 2>&1 | tail -3
echo "== attn half"; OVG_ATTN_SINGLE=2 KB=attn timeout 120 python tools/kbench.py 2>&1 | grep "^attn"
echo "== attn single (default)"; KB=attn timeout 120 python tools/kbench.py 2>&1 | grep "^attn\|sdpa"
echo "== attn paired"; OVG_ATTN_SINGLE=0 KB=attn timeout 120 python tools/kbench.py 2>&1 | grep "^attn"
} > gpurun_out/call33.txt 2>&1
tail -20 gpurun_out/call33.txt
