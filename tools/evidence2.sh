#!/bin/bash
# round-2 final evidence, part 2: the other bench configurations (cfg1, cfg3, cfg5 on one GPU; cfg4 = 32 scenes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in cfg1 cfg3 cfg5 cfg4; do
  echo "== $c"; timeout 900 python bench.py --config $c --no-cpu-baseline --no-gpu-torch-baseline 2>&1 | tail -1 | tee gpurun_out/r02_bench_$c.json | cut -c1-220
done
