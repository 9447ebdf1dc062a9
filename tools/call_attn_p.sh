#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/attn_ab.jsonl
rm -f $OUT
echo "== tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 2>&1 | grep "passed\|failed\|FAILED\|^E  \|rel-L2" | cut -c1-250 | head
echo "== split"; ATTN_SHAPES=global8,global4,global24 timeout 300 python tools/attn_bench.py split 2>&1 | tail -1 | tee -a $OUT
echo "== nosplit"; ATTN_SPLIT=0 ATTN_SHAPES=global8,global4,global24 ATTN_SDPA=0 timeout 300 python tools/attn_bench.py nosplit 2>&1 | tail -1 | tee -a $OUT
