#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/model_parity_full.txt gpurun_out/model_parity_mini.txt
{
echo "== kernel tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "ln_fold or resid or inject or layernorm or qkv" 2>&1 | grep "passed\|failed\|FAILED\|^E  " | cut -c1-300 | head -20
echo "== model tests"; timeout 1500 python -m pytest tests/test_model_gpu.py -q 2>&1 | grep "passed\|failed\|FAILED\|^E  " | cut -c1-400 | head -20
cat gpurun_out/model_parity_full.txt
echo "== bench fold"; timeout 900 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fold.json | cut -c1-200
echo "== bench no fold"; OVG_LN_FOLD=0 timeout 900 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>&1 | tail -1 | tee gpurun_out/bench_nofold.json | cut -c1-200
echo "== bench fold again"; timeout 900 python bench.py --no-cpu-baseline --no-gpu-torch-baseline 2>&1 | tail -1 | cut -c1-200
} > gpurun_out/fold.txt 2>&1
tail -40 gpurun_out/fold.txt
