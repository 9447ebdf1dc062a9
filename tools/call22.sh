#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=omnivggt-official_b200/variants
{
for r in 1 2; do
echo "== attn v10 (default) run $r"; KB=attn timeout 200 python tools/kbench.py 2>&1 | grep "^attn"
echo "== attn v11 run $r"; OVG_LIB_PATH=$V/libovg_v11.so KB=attn timeout 200 python tools/kbench.py 2>&1 | grep "^attn"
done
echo "== qkv default"; KB=gemm timeout 200 python tools/kbench.py 2>&1 | grep "qkv"
echo "== qkv nostore"; OVG_LIB_PATH=$V/libovg_nostore.so KB=gemm timeout 200 python tools/kbench.py 2>&1 | grep "qkv"
echo "== launches"; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_call22.csv python tools/profile_step.py 2>&1 | tail -2
} > gpurun_out/call22.txt 2>&1
tail -40 gpurun_out/call22.txt
