#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/attn_ab.jsonl
rm -f $OUT
echo "== tests (persistent)"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 2>&1 | tail -3
echo "== default"; timeout 300 python tools/attn_bench.py persistent 2>&1 | tail -1 | tee -a $OUT
for v in "$@"; do
  echo "== $v"; OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so timeout 300 python tools/attn_bench.py $v 2>&1 | tail -1 | tee -a $OUT
done
