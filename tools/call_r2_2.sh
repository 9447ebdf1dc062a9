#!/bin/bash
# round 2, GPU call 2: attn4 (packed f16 softmax) variants vs attn1; post-processing tests; bench line with the GPU library baseline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c2
rm -f $OUT.*
echo "== attn4 E0 (default lib)"; timeout 300 python tools/attn_bench.py attn4_e0 2>&1 | tail -2 | tee -a $OUT.attn.jsonl
for v in a4e2 a4e4; do
  echo "== $v"; OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so ATTN_SDPA=0 timeout 300 python tools/attn_bench.py $v 2>&1 | tail -2 | tee -a $OUT.attn.jsonl
done
echo "== attn1 E0"; OVG_ATTN_KERNEL=1 OVG_LIB_PATH=$PWD/build_ab/libovg_a1e0.so ATTN_SDPA=0 timeout 300 python tools/attn_bench.py attn1_e0 2>&1 | tail -2 | tee -a $OUT.attn.jsonl
echo "== pytest attention + postprocess"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_postprocess.py -m gpu -q -k "attention or postprocess or pose or percentile" 2>&1 | tail -15 | tee $OUT.pytest.txt
echo "== bench cfg2"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $OUT.bench.json
