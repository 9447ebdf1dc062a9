#!/bin/bash
# error budget of the dense outputs + streaming-pipeline test + a bench line (one gpurun call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== error budget"; timeout 900 python tools/error_budget.py 2>&1 | tail -30
echo "== pipe check"; timeout 600 python tools/pipe_check.py 2>&1 | tail -12

} > gpurun_out/eb.txt 2>&1
tail -50 gpurun_out/eb.txt
