#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
bash tools/call_ab_bench.sh nopdl
