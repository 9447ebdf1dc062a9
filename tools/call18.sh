#!/bin/bash
# A/B sweep: GEMM tile choices, fc1 GELU cost, LayerNorm launch geometry, attention exp-emulation share
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=omnivggt-official_b200/variants
{
echo "== gemm"; KB=gemm timeout 300 python tools/kbench.py 2>&1 | grep -v "^attn\|sdpa"
for t in 128 512; do echo "== LN threads $t"; OVG_LN_THREADS=$t KB=attn timeout 200 python tools/kbench.py 2>&1 | grep layernorm; done
for b in 2 4 5 8; do echo "== LN persist $b"; OVG_LN_PERSIST=$b KB=attn timeout 200 python tools/kbench.py 2>&1 | grep layernorm; done
for e in 2 6; do echo "== attn emu $e"; OVG_LIB_PATH=$V/libovg_emu$e.so KB=attn timeout 200 python tools/kbench.py 2>&1 | grep "^attn"; done
echo "== attn default"; KB=attn timeout 200 python tools/kbench.py 2>&1 | grep "^attn\|sdpa"
} > gpurun_out/call18.txt 2>&1
tail -60 gpurun_out/call18.txt
