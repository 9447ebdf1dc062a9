#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels added last (fused DPT tail, fp16 mode) at small shapes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -x -k "dpt_tail and not 296 or fp16_conv or fp16_upsample or fp16_layernorm" 2>&1 | tail -15 > gpurun_out/sanitize.txt
echo "rc=$?" >> gpurun_out/sanitize.txt
tail -20 gpurun_out/sanitize.txt
