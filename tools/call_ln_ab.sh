#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/ln_ab.jsonl; rm -f $OUT
timeout 300 python tools/ln_bench.py base 2>&1 | tail -1 | tee -a $OUT
for v in "$@"; do OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so timeout 300 python tools/ln_bench.py $v 2>&1 | tail -1 | tee -a $OUT; done
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "layernorm" 2>&1 | tail -2
