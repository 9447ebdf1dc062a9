#!/bin/bash
# same-box A/B of whole-step time, alternating: "default", library builds build_ab/libovg_<tag>.so (argument "<tag>") and
# environment variants (argument "NAME=VALUE", applied to the default library)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_bench.txt
for rep in 1 2; do
  for v in default "$@"; do
    unset OVG_LIB_PATH
    envs=""
    if [[ $v == *=* ]]; then envs="$v"; elif [ $v != default ]; then export OVG_LIB_PATH=$PWD/build_ab/libovg_$v.so; fi
    env $envs timeout 600 python bench.py --no-cpu-baseline --no-gpu-torch-baseline --steps 20 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), 'ms/step  e2e', round(d['e2e']['ms_per_step'],3), ' attn avg ms', round(d['roofline']['avg_launch_ms'],4), 'clk', d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab_bench.txt
  done
done
