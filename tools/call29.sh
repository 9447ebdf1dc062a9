#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
good=0
for mode in 2 1; do
  echo "== head_tail mode $mode"
  if OVG_HT_ROWSHIFT=$mode timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q --timeout 60 -k head_tail 2>&1 | tail -3 | tee /tmp/ht_$mode.txt | grep -q "passed" && ! grep -q failed /tmp/ht_$mode.txt; then good=$mode; break; fi
  cat /tmp/ht_$mode.txt
done
echo "== working mode: $good"
if [ "$good" != "0" ]; then
  echo "== tests all"; OVG_HT_ROWSHIFT=$good timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 2>&1 | tail -3
  echo "== bench rowshift"; OVG_HT_ROWSHIFT=$good timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
fi
echo "== bench generic"; OVG_HT_ROWSHIFT=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
} > gpurun_out/call29.txt 2>&1
tail -40 gpurun_out/call29.txt
