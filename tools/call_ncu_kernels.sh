#!/bin/bash
# ncu --set full of the block / DPT / gather-scatter kernels (tools/ncu_kernels.py); summary: tools/ncu_summary.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r02_kernels python tools/ncu_kernels.py 2>&1 | tail -2
