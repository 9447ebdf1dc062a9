#!/bin/bash
# round-2 evidence in one gpurun call: GPU tests, bench line, ncu launch list of one step of the bench command itself,
# ncu --set full of the block-level + gather/scatter kernels.  Summaries: tools/ncu_summary.py (run where ncu is, no GPU needed).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/r02_bench_cfg2.json | cut -c1-300
echo "== launches of one step of the bench command"; OVG_BENCH_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_cmd.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-torch-baseline 2>&1 | tail -1 | cut -c1-200
echo "== ncu full"; timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r02_kernels python tools/ncu_kernels.py 2>&1 | tail -2
} > gpurun_out/evidence.txt 2>&1
tail -30 gpurun_out/evidence.txt
