#!/bin/bash
# round-1 evidence: GPU tests, bench line, ncu launch list of one forward, ncu --set full of the block-level kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== tests"; timeout 400 python -m pytest tests -m gpu -q --timeout 90 2>&1 | tail -4
echo "== bench"; timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-300
echo "== launches"; OVG_CUDA_GRAPH=0 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py 2>&1 | tail -1
echo "== ncu full"; timeout 500 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r01_block_kernels_final python tools/ncu_kernels.py 2>&1 | tail -2
} > gpurun_out/evidence.txt 2>&1
tail -30 gpurun_out/evidence.txt
