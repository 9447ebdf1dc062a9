#!/bin/bash
# round 2, GPU call 6: attention without the running maximum / interleaved polynomial lanes (A/B), kernel tests, ncu of the gather/scatter kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c6
rm -f $OUT.*
echo "== default"; ATTN_SHAPES=global8,frame8 timeout 300 python tools/attn_bench.py base 2>&1 | tail -2 | tee -a $OUT.attn.jsonl
for v in nomax il nomaxil nomaxe5 nomaxe6; do
  echo "== $v"; OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so ATTN_SHAPES=global8,frame8 timeout 300 python tools/attn_bench.py $v 2>&1 | tail -2 | tee -a $OUT.attn.jsonl
done
echo "== attention tests on the nomax build"; OVG_LIB_PATH=$PWD/build_ab/libovg_nomax.so timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k attention 2>&1 | tail -5
echo "== pytest kernels"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $OUT.pytest.txt
echo "== ncu scatter"; NCU_ONLY=scatter timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r02_scatter python tools/ncu_kernels.py 2>&1 | tail -2
echo "== bench cfg2"; timeout 900 python bench.py --steps 10 --warmup 3 --no-gpu-torch-baseline 2>&1 | tail -3 | tee $OUT.bench.json | cut -c1-1500
