#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/tail python tools/tail_ncu.py 2>&1 | tail -3
