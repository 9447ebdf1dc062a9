#!/bin/bash
# attention A/B: tools/call_attn_ab.sh tag1 tag2 ... (libraries build_ab/libovg_<tag>.so, built by tools/build_ab.sh); the default library first
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/attn_ab.jsonl
rm -f $OUT
echo "== default"; ATTN_SHAPES=global8,frame8 timeout 300 python tools/attn_bench.py base 2>&1 | tail -1 | tee -a $OUT
for v in "$@"; do
  echo "== $v"; OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so ATTN_SHAPES=global8,frame8 timeout 300 python tools/attn_bench.py $v 2>&1 | tail -1 | tee -a $OUT
  OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" 2>&1 | tail -1
done
