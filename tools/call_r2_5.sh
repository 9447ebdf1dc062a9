#!/bin/bash
# round 2, GPU call 5: multicast quad GEMM (cluster of two CTA pairs) -- parity tests and kernel bench vs pairs / cuBLAS
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2c5
rm -f $OUT.*
echo "== pytest gemm"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" 2>&1 | tail -12 | tee $OUT.pytest.txt
echo "== kbench"; timeout 600 python tools/kbench.py 2>&1 | grep -E "gemm_|cublas_" | tee $OUT.kbench.txt
