#!/bin/bash
# round 2, GPU call 4: full GPU test-suite, ncu --set full (attention with source view; block + gather/scatter kernels), bench lines of cfg1/3/4/5
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c4
rm -f $OUT.*
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT.pytest.txt
echo "== ncu attention"; NCU_ONLY=attn timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r02_attn python tools/ncu_kernels.py 2>&1 | tail -2
echo "== ncu kernels"; timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r02_kernels python tools/ncu_kernels.py 2>&1 | tail -2
for c in cfg1 cfg3 cfg5; do
  echo "== bench $c"; timeout 900 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT.bench_$c.json | cut -c1-400
done
echo "== bench cfg4"; timeout 900 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT.bench_cfg4.json | cut -c1-400
