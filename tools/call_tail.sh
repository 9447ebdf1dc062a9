#!/bin/bash
# fused DPT tail: tests, micro-benchmark against the two-kernel path, clock64 timeline of CTA 0 (one gpurun call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "dpt_tail" 2>&1 | tail -3
timeout 600 python tools/tail_bench.py base 2>&1 | tail -1 | tee gpurun_out/tail_ab.jsonl
for v in "$@"; do OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so timeout 300 python tools/tail_bench.py $v 2>&1 | tail -1 | tee -a gpurun_out/tail_ab.jsonl; done
TAIL_PROF=1 timeout 300 python tools/tail_ncu.py 2>&1 | tail -60
