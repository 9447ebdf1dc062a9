#!/bin/bash
# retry a gpurun call while the pod answers "busy" (nothing is charged for those); usage: tools/gpurun_retry.sh <timeout> <script>
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $1 -- "bash $2" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|retry in a few minutes"; then sleep 120; continue; fi
  break
done
