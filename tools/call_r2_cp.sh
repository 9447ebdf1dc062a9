#!/bin/bash
# context-parallel check on N GPUs of one box: tools/call_r2_cp.sh N [views_full]
cd "$(dirname "$0")/.."
N=${1:-2}; V=${2:-8}
mkdir -p gpurun_out
OUT=gpurun_out/r2cp_n$N
rm -f $OUT.*
nvidia-smi topo -m 2>/dev/null | head -12 > $OUT.topo.txt
echo "== reduced model, bit-exactness"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/cp_check.py 2>&1 | tail -12 | tee $OUT.small.txt
echo "== full model, $V views @518"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/cp_check.py --full --views $V 2>&1 | tail -12 | tee $OUT.full.txt
