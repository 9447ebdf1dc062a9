#!/bin/bash
# round 2, GPU call 7: attention without the running maximum / interleaved polynomial lanes (A/B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c7
rm -f $OUT.*
echo "== default"; ATTN_SHAPES=global8,frame8 timeout 300 python tools/attn_bench.py base 2>&1 | tail -1 | tee -a $OUT.attn.jsonl
for v in nomax il nomaxil nomaxe5 nomaxe6; do
  echo "== $v"; OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so ATTN_SHAPES=global8,frame8 timeout 300 python tools/attn_bench.py $v 2>&1 | tail -1 | tee -a $OUT.attn.jsonl
done
echo "== attention tests on the nomax build"; OVG_LIB_PATH=$PWD/build_ab/libovg_nomax.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention or im2col" 2>&1 | tail -4
