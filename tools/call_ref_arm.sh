#!/bin/bash
# the reference arm of bench.py on the box's host cores (the driver runs this at round end)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_reference_arm.json | cut -c1-600
