#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=omnivggt-official_b200/variants
{
echo "== tests attention"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 60 -k attention 2>&1 | tail -3
for r in 1 2; do
echo "== attn sumcheck run $r"; KB=attn timeout 120 python tools/kbench.py 2>&1 | grep "^attn"
echo "== attn prev (max tracking) run $r"; OVG_LIB_PATH=$V/libovg_prev.so KB=attn timeout 120 python tools/kbench.py 2>&1 | grep "^attn"
done
for e in 2 6; do echo "== attn emu $e"; OVG_LIB_PATH=$V/libovg_emu$e.so KB=attn timeout 120 python tools/kbench.py 2>&1 | grep "^attn"; done
echo "== tests all"; timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -3
echo "== bench"; timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
} > gpurun_out/call28.txt 2>&1
tail -60 gpurun_out/call28.txt
