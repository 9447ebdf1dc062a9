#!/bin/bash
# correctness of the new attention / epilogue / LN code, then A/B numbers
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=omnivggt-official_b200/variants
{
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== kbench"; timeout 300 python tools/kbench.py 2>&1 | grep -v "^cublas"
for e in 0 2 3; do echo "== attn emu $e"; OVG_LIB_PATH=$V/libovg_emu$e.so KB=attn timeout 200 python tools/kbench.py 2>&1 | grep "^attn"; done
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1
} > gpurun_out/call19.txt 2>&1
tail -80 gpurun_out/call19.txt
