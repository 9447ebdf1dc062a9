#!/bin/bash
# 2-GPU checks in one call: the pytest CP test, the full-size CP check, and the 2-rank weak-scaling bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest cp"; timeout 900 python -m pytest tests/test_cp_gpu.py -q 2>&1 | tail -3
echo "== full model, 8 views, 2 ranks"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/cp_check.py --full --views 8 2>&1 | grep -B12 -m1 "Error\|CP_CHECK" | tail -16
echo "== bench N=2 weak"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-cpu-baseline --no-gpu-torch-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n2.json | cut -c1-260
} > gpurun_out/cp2.txt 2>&1
tail -40 gpurun_out/cp2.txt
