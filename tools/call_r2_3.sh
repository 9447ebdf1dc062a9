#!/bin/bash
# round 2, GPU call 3: attention with the deferred PV(j-1) wait (A/B over the number of chunks held), post-processing tests, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c3
rm -f $OUT.*
echo "== default lib (LW1 E4)"; timeout 300 python tools/attn_bench.py lw1e4 2>&1 | tail -4 | tee -a $OUT.attn.jsonl
for v in lw0 lw2 lw3 lw2e5 lw2e6 lw1e3; do
  echo "== $v"; OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so ATTN_SDPA=0 timeout 300 python tools/attn_bench.py $v 2>&1 | tail -4 | tee -a $OUT.attn.jsonl
done
echo "== pytest attention + postprocess"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_postprocess.py -m gpu -q -k "attention or postprocess or pose or percentile" 2>&1 | tail -30 | tee $OUT.pytest.txt
echo "== bench cfg2"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $OUT.bench.json
