#!/bin/bash
# context-parallel check on N GPUs of one box: tools/call_r2_cp.sh N [views_full]
cd "$(dirname "$0")/.."
N=${1:-2}; V=${2:-8}
mkdir -p gpurun_out
OUT=gpurun_out/r2cp_n$N
rm -f $OUT.*
nvidia-smi topo -m 2>/dev/null | head -12 > $OUT.topo.txt
echo "== reduced model, bit-exactness"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/cp_check.py > $OUT.small.txt 2>&1; grep -v "^\s*$" $OUT.small.txt | grep -B30 -m1 "Error\|error\|CP_CHECK" | tail -45
echo "== full model, $V views @518"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/cp_check.py --full --views $V > $OUT.full.txt 2>&1; grep -B25 -m1 "Error\|CP_CHECK" $OUT.full.txt | tail -40
