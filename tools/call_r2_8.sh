#!/bin/bash
# round 2, GPU call 8: full GPU test-suite, smoke, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c8
rm -f $OUT.* gpurun_out/model_parity_full.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT.pytest.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -7
echo "== bench cfg2"; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | tee $OUT.bench.json | cut -c1-600
