#!/bin/bash
# A/B builds of libovg for one gpurun call: build_ab/libovg_<tag>.so with extra -D flags.  Usage: tools/build_ab.sh tag "-DX=1 -DY=2" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build_ab
while [ $# -gt 0 ]; do
  tag=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC $flags \
    -o build_ab/libovg_$tag.so omnivggt-official_b200/csrc/ovg.cu &
done
wait
ls -la build_ab
