#!/bin/bash
# round 2, GPU call 1: attention A/B (new 4-tiles-per-SM kernel vs round-1 kernel vs torch SDPA), GPU test-suite, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c1
rm -f $OUT.*
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT.smi.txt 2>&1
echo "== attn default (attn3 E4)"; timeout 300 python tools/attn_bench.py attn3_e4 2>&1 | tail -2 | tee -a $OUT.attn.jsonl
echo "== attn1 (round-1 kernel)"; OVG_ATTN_KERNEL=1 ATTN_SDPA=0 timeout 300 python tools/attn_bench.py attn1 2>&1 | tail -2 | tee -a $OUT.attn.jsonl
for v in e2 e6 e8 safe; do
  echo "== attn3 $v"; OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so ATTN_SDPA=$([ $v = safe ] && echo 1 || echo 0) timeout 300 python tools/attn_bench.py attn3_$v 2>&1 | tail -2 | tee -a $OUT.attn.jsonl
done
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee $OUT.pytest.txt
echo "== bench cfg2"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $OUT.bench.json
