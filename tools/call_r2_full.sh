#!/bin/bash
# full GPU suite + smoke() + one bench line (one gpurun call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/model_parity_full.txt gpurun_out/model_parity_mini.txt
{
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_now.json | cut -c1-300
} > gpurun_out/full.txt 2>&1
tail -30 gpurun_out/full.txt
