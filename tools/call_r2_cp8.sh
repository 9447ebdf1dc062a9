#!/bin/bash
# context parallelism on all GPUs of the box: one 24-view scene (BASELINE.json configs[4]) sharded over N ranks
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
OUT=gpurun_out/r2cp_n$N
rm -f $OUT.*
echo "== full model, 24 views @518, N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/cp_check.py --full --views 24 --steps 3 > $OUT.full.txt 2>&1; grep -B25 -m1 "Error\|CP_CHECK" $OUT.full.txt | tail -12 | cut -c1-500
echo "== bench --cp cfg5"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --config cfg5 --cp --steps 10 --warmup 3 > $OUT.bench.txt 2>&1; tail -1 $OUT.bench.txt | cut -c1-1800
