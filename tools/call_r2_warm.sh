#!/bin/bash
# launch list of one bench step WITHOUT the profiler's cache flush between kernels (in-situ L2 state)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OVG_BENCH_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-torch-baseline 2>&1 | tail -1 | cut -c1-200
