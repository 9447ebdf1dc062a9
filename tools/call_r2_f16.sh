#!/bin/bash
# fp16 DPT heads: kernel tests, model goldens, error budget, bench with the GPU library baseline (one gpurun call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/model_parity_full.txt
{
echo "== fp16 kernel tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "fp16 or conv or head_tail or upsample or layernorm or gemm_bf16" 2>&1 | tail -6
echo "== model tests"; timeout 1500 python -m pytest tests/test_model_gpu.py -q 2>&1 | tail -6
cat gpurun_out/model_parity_full.txt
echo "== error budget"; timeout 900 python tools/error_budget.py 2>&1 | grep -v "^sim"
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_f16.json | cut -c1-300
} > gpurun_out/f16.txt 2>&1
tail -60 gpurun_out/f16.txt
