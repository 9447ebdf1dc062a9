// Microbenchmark: TMEM read bandwidth (tcgen05.ld) and MUFU.EX2 throughput per SM on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../omnivggt-official_b200/csrc/ptx.cuh"
using namespace ovg;

template <int NWARPS, int MODE>   // MODE 0: ld32 + wait each; 1: 4x ld32 then one wait; 2: MUFU only; 3: ld + MUFU interleaved
__global__ void __launch_bounds__(NWARPS * 32, 1) k(long long* out, float* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { uint32_t r[32]; tmem_ld32(base + c * 32, r); tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(r[i]); }
    } else if (MODE == 1) {
      uint32_t r[128];
      tmem_ld32(base, r); tmem_ld32(base + 32, r + 32); tmem_ld32(base + 64, r + 64); tmem_ld32(base + 96, r + 96);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 128; ++i) acc += __uint_as_float(r[i]);
    } else if (MODE == 2) {
      float x = acc + it;
#pragma unroll
      for (int i = 0; i < 128; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x + i)); acc += y; }
    } else {
      uint32_t r[128];
      tmem_ld32(base, r); tmem_ld32(base + 32, r + 32); tmem_ld32(base + 64, r + 64); tmem_ld32(base + 96, r + 96);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 128; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(r[i]))); acc += y; }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int NW, int MODE> void run(const char* name) {
  long long* d; float* s; cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 148 * NW * 32 * 4);
  const int iters = 2000;
  k<NW, MODE><<<148, NW * 32>>>(d, s, iters); cudaDeviceSynchronize();
  k<NW, MODE><<<148, NW * 32>>>(d, s, iters); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = double(h[0]) / iters;
  printf("%-34s warps=%2d  %8.1f cyc / iteration (128 elems per thread) -> %.1f elems/clk/SM%s\n", name, NW, cyc,
         NW * 32 * 128.0 / cyc, cudaGetLastError() == cudaSuccess ? "" : "  [CUDA ERROR]");
  cudaFree(d); cudaFree(s);
}
int main() {
  run<4, 0>("tmem ld32+wait x4");  run<8, 0>("tmem ld32+wait x4");
  run<4, 1>("tmem 4x ld32, one wait"); run<8, 1>("tmem 4x ld32, one wait"); run<16, 1>("tmem 4x ld32, one wait");
  run<4, 2>("MUFU.EX2 only"); run<8, 2>("MUFU.EX2 only");
  run<4, 3>("tmem ld + MUFU"); run<8, 3>("tmem ld + MUFU"); run<16, 3>("tmem ld + MUFU");
  return 0;
}
