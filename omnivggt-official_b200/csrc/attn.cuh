// Fused softmax(Q K^T) V for head_dim 64, non-causal, ragged sequence length (sm_100a, tcgen05 + TMA).
// Replaces F.scaled_dot_product_attention at reference layers/attention.py:61-66 for both the frame-wise
// (batch = B*S, N = 1374) and the global (batch = B, N = S*1374) attention of models/aggregator.py:312-341.
//
// q is pre-scaled by (1/sqrt(64))*log2(e) in the QKV GEMM epilogue, so probabilities are exp2(s - m).
//
// attn3_kernel (end of this file) is the product kernel; attn1_kernel is the round-1 kernel kept for one A/B run.
//
// Online softmax with a stale reference and lazy rescaling (both kernels): P(j) = exp2(S - m_ref) is computed against
// the reference left by earlier steps while this step's row maximum is gathered in the same pass; O / l are rescaled
// (and the chunk redone from the S values still in registers) only when the maximum exceeds the reference by > 8 (log2
// units) -- exact, the reference cancels in O / l, and P stays <= 256.
#pragma once
#include "ptx.cuh"

namespace ovg {

struct AttnParams {
  int n;        // sequence length (keys == queries)
  int heads;
  int C;        // heads * 64 (row stride of `out`)
  __nv_bfloat16* out;  // [batch, n, C]
  long long* prof;     // optional [2][8] cycle counters (OVG_ATT_PROFILE builds only)
};
#ifdef OVG_ATT_PROFILE
#define ATT_T(var) const long long var = clock64()
#define ATT_ACC(slot, a, b) do { if (prof_on) prof_acc[slot] += (b) - (a); } while (0)
#else
#define ATT_T(var)
#define ATT_ACC(slot, a, b)
#endif

constexpr int ATT_TILE_BYTES = 128 * 64 * 2;

template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// exp2 on the FMA pipes for a pair of values (x <= ~8): Cody-Waite split x = n + r, r in [-0.5, 0.5], 2^r by a
// degree-3 minimax polynomial (max rel. error 1.0e-4, far below bf16 resolution of P), exponent inserted with one IMAD.
// At head_dim 64 the softmax needs 16 384 exp2 per 128x128 tile against 512 tensor-pipe clocks; the MUFU unit alone
// (16/clk/SM) caps the kernel at 50% tensor utilisation, so a fraction of the exponentials is moved here.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  const float kMagic = 12582912.0f;   // 1.5 * 2^23: adding it rounds x to the nearest integer in the low mantissa bits
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 t = fadd2(x, make_float2(kMagic, kMagic));
  const float2 n = fadd2(t, make_float2(-kMagic, -kMagic));
  const float2 r = fadd2(x, make_float2(-n.x, -n.y));
  float2 p = ffma2(make_float2(0.05500871f, 0.05500871f), r, make_float2(0.24221068f, 0.24221068f));
  p = ffma2(p, r, make_float2(0.69328292f, 0.69328292f));
  p = ffma2(p, r, make_float2(1.0f, 1.0f));
  float2 o;
  o.x = __int_as_float(__float_as_int(t.x) * 8388608 + __float_as_int(p.x));
  o.y = __int_as_float(__float_as_int(t.y) * 8388608 + __float_as_int(p.y));
  return o;
}
#ifndef OVG_ATT_EMU_PAIRS
#define OVG_ATT_EMU_PAIRS 4   // of every 16 element pairs, how many take the polynomial path (0 = all MUFU); 4 measured best
#endif
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------------------------------
// Single-tile variant: CTA = 128 query rows, 256 threads, TWO CTAs per SM (256 TMEM columns, ~112 KB smem and half of
// the register file each).  The two tiles that share an SM are
// independent CTAs: their phases drift freely, one CTA's prologue (barrier init, TMEM alloc, Q / first K loads) and
// epilogue (O read-out, global stores, TMEM free) run under the other CTA's main loop, and the grid quantises in
// 128-row units.  Cost: every CTA streams K/V for itself (2x the L2->SM bytes of the paired kernel).
//   warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-3 idle, warps 4-7 softmax (one query row per thread).
// TMEM (256 cols): S [0,128) P [128,192) O [192,256).
constexpr int ATT1_THREADS = 256;
constexpr int ATT1_KV_STAGES = 3;
constexpr int ATT1_SMEM_BYTES = (1 + 2 * ATT1_KV_STAGES) * ATT_TILE_BYTES + 256;   // base must be 1024-aligned (checked)

__global__ void __launch_bounds__(ATT1_THREADS, 2)
attn1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int NS = ATT1_KV_STAGES;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = sK + NS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;             // [1]
  uint64_t* k_full = bars + 1;         // [NS]
  uint64_t* k_empty = k_full + NS;     // [NS]
  uint64_t* v_full = k_empty + NS;     // [NS]
  uint64_t* v_empty = v_full + NS;     // [NS]
  uint64_t* s_full = v_empty + NS;     // [1]
  uint64_t* p_full = s_full + 1;       // [1]
  uint64_t* o_ready = p_full + 1;      // [1]
  uint64_t* s_taken = o_ready + 1;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_taken + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int head = blockIdx.y;
  const int bh = blockIdx.z * p.heads + head;
  const int nkv = (p.n + 127) / 128;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("ovg attn1: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_ready, 1);
    mbar_init(s_taken, 4);
    for (int i = 0; i < NS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // register budget: launch allocation 128 * 256 = 32768 per CTA; after the split 128*40 + 128*208 = 31744
  if (warp < 4) reg_dealloc<40>();
  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, 0, q0, bh);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], ATT_TILE_BYTES);
        tma_load_3d(sK + s * ATT_TILE_BYTES, &tmK, &k_full[s], 0, j * 128, bh);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], ATT_TILE_BYTES);
        tma_load_3d(sV + s * ATT_TILE_BYTES, &tmV, &v_full[s], 0, j * 128, bh);
        if (++s == NS) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (=V) is MN-major
      const uint32_t tS = tmem_base;
      const uint32_t tP = tmem_base + 128;
      const uint32_t tO = tmem_base + 192;
      const uint64_t qdesc = make_sw128_desc(smem_u32(sQ));
      const uint64_t kdesc0 = make_sw128_desc(smem_u32(sK));
      const uint64_t vdesc0 = make_sw128_desc(smem_u32(sV));
      constexpr uint64_t kStageStep = ATT_TILE_BYTES >> 4;
      auto issue_S = [&](int stage) {
        const uint64_t bdesc = kdesc0 + stage * kStageStep;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS, qdesc + 2 * k, bdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      auto issue_PV = [&](int stage, int j) {
        const uint64_t bdesc = vdesc0 + stage * kStageStep;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tO, tP + 8 * k, bdesc + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        umma_commit(o_ready);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_S(0);
      umma_commit(&k_empty[0]);
      int s = 0, sn = 1 % NS;
      uint32_t ph = 0, phn = (NS == 1) ? 1u : 0u;
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {              // S(j+1) as soon as the softmax warps hold S(j) in registers
          mbar_wait(&k_full[sn], phn);
          mbar_wait(s_taken, j & 1);
          tc_fence_after();
          issue_S(sn);
          umma_commit(&k_empty[sn]);
        }
        mbar_wait(&v_full[s], ph);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        issue_PV(s, j);
        umma_commit(&v_empty[s]);
        s = sn;
        ph = phn;
        if (++sn == NS) {
          sn = 0;
          phn ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    reg_alloc<208>();
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qrow = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tS = tmem_base + lane_off;
    const uint32_t tP = tmem_base + 128 + lane_off;
    const uint32_t tO = tmem_base + 192 + lane_off;
    float m_used = -INFINITY;
    float l = 0.f;
    for (int j = 0; j < nkv; ++j) {
      const int kv_valid = min(128, p.n - j * 128);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t raw[128];
      tmem_ld32(tS, raw);
      tmem_ld32(tS + 32, raw + 32);
      tmem_ld32(tS + 64, raw + 64);
      tmem_ld32(tS + 96, raw + 96);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_taken);
      if (kv_valid != 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_valid) raw[i] = 0xff800000u;
      }
      // stale-reference softmax step (see the file header)
      float2 acc = make_float2(0.f, 0.f);
      bool slow = (j == 0);
      float m_new = m_used;
      if (j > 0) {
        const float2 negm = make_float2(-m_used, -m_used);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float r0 = __uint_as_float(raw[c * 32 + 2 * i]), r1 = __uint_as_float(raw[c * 32 + 2 * i + 1]);
            if (i & 1) mx1 = fmaxf(fmaxf(mx1, r0), r1); else mx0 = fmaxf(fmaxf(mx0, r0), r1);
            float2 x = fadd2(make_float2(r0, r1), negm);
            if (i >= 16 - OVG_ATT_EMU_PAIRS) {
              x = exp2_poly2(x);
            } else {
              x.x = ex2_approx(x.x);
              x.y = ex2_approx(x.y);
            }
            acc = fadd2(acc, x);
            pk[i] = pack_bf16(x.x, x.y);
          }
          if (c == 0) {
            mbar_wait(o_ready, (j - 1) & 1);   // PV(j-1) complete: P buffer reusable, O stable
            tc_fence_after();
          }
          tmem_st16(tP + c * 16, pk);
        }
        m_new = fmaxf(m_used, fmaxf(mx0, mx1));
        slow = __any_sync(0xffffffffu, (m_new - m_used) > 8.0f);
      }
      if (slow) {
        if (j == 0) {
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 128; i += 4) {
            mx0 = fmaxf(fmaxf(mx0, __uint_as_float(raw[i])), __uint_as_float(raw[i + 1]));
            mx1 = fmaxf(fmaxf(mx1, __uint_as_float(raw[i + 2])), __uint_as_float(raw[i + 3]));
          }
          m_used = fmaxf(mx0, mx1);
        } else {
          const bool need = (m_new - m_used) > 8.0f;
          const float alpha = need ? ex2_approx(m_used - m_new) : 1.0f;
          if (need) {
            m_used = m_new;
            l *= alpha;
          }
          tmem_st_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tO + c * 32, o);
          }
        }
        const float2 negm = make_float2(-m_used, -m_used);
        acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 x = make_float2(__uint_as_float(raw[c * 32 + 2 * i]), __uint_as_float(raw[c * 32 + 2 * i + 1]));
            x = fadd2(x, negm);
            x.x = ex2_approx(x.x);
            x.y = ex2_approx(x.y);
            acc = fadd2(acc, x);
            pk[i] = pack_bf16(x.x, x.y);
          }
          tmem_st16(tP + c * 16, pk);
        }
      }
      l += acc.x + acc.y;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> out[b, qrow, head*64 .. +64)
    mbar_wait(o_ready, (nkv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l;
    uint32_t o[64];
    tmem_ld32(tO, o);
    tmem_ld32(tO + 32, o + 32);
    tmem_ld_wait();
    if (qrow < p.n) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(blockIdx.z) * p.n + qrow) * p.C + head * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
        w.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
        w.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
        w.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
        dst[i] = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// attn3_kernel: four query tiles per SM.  CTA = two 128-row query tiles of one (batch, head) that share one K/V ring,
// KV steps of 64 keys, 384 threads, TWO CTAs per SM.
//   warp 0        TMA producer: both Q tiles once, K / V tiles of 64 keys through a 4-stage ring
//   warps 1, 2    MMA issuers, one per query tile: S_t = Q_t K^T (SS, M128 N64 K64), O_t += P_t V (TS: P from TMEM, V as
//                 MN-major smem operand, M128 N64 K64); issue order PV_t(j), S_t(j+1)
//   warp 3        idle (setmaxnreg is per warpgroup)
//   warps 4-7     softmax of tile 0, warps 8-11 softmax of tile 1: one query row per thread, the 64-wide S row in two
//                 32-column chunks (32 + 16 live registers instead of the 128 + 16 of attn1_kernel)
// TMEM (256 columns per CTA): tile t owns [128 t, 128 t + 128): S_t [0, 64), O_t [64, 128).  P_t (bf16, 32 columns) is
// written IN PLACE over the first half of S_t, so one tile needs 128 columns and four tiles fit an SM.
// Why: at head_dim 64 the kernel is bound by the softmax warps (MUFU.EX2 16/clk/SM + FMA-pipe emulation + in-order issue),
// not by the tensor pipe.  attn1_kernel runs 2 softmax warps per SM sub-partition (208 registers each); ncu showed the
// issue slots at 50% and MUFU at 55% with the warps stalled on fixed-latency dependencies.  Here 4 softmax warps share a
// sub-partition (104 registers each), so one warp's dependency stalls, TMEM reads and barrier waits are filled by three
// others, and two tiles share every K/V byte staged from L2.  In-place P serialises S(j+1) behind PV(j) inside a tile
// (tcgen05.mma instructions of one thread execute in issue order); the other three tiles cover that latency.
constexpr int ATT3_THREADS = 384;
constexpr int ATT3_NS = 4;
constexpr int ATT3_KV_BYTES = 64 * 64 * 2;
constexpr int ATT3_SMEM_BYTES = 2 * ATT_TILE_BYTES + 2 * ATT3_NS * ATT3_KV_BYTES + 256;   // base must be 1024-aligned (checked)
#ifndef OVG_ATT3_SAFE_ORDER
#define OVG_ATT3_SAFE_ORDER 0    // 1: the issuer waits for PV(j) to complete before S(j+1) overwrites P(j) (validation builds)
#endif

__global__ void __launch_bounds__(ATT3_THREADS, 2)
attn3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int NS = ATT3_NS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + NS * ATT3_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NS * ATT3_KV_BYTES);
  uint64_t* q_full = bars;             // [2]
  uint64_t* k_full = bars + 2;         // [NS]
  uint64_t* k_empty = k_full + NS;     // [NS]
  uint64_t* v_full = k_empty + NS;     // [NS]
  uint64_t* v_empty = v_full + NS;     // [NS]
  uint64_t* s_full = v_empty + NS;     // [2]  S_t(j) complete (implies PV_t(j-1) complete: same issuing thread)
  uint64_t* p_full = s_full + 2;       // [2]  P_t(j) stored by the 4 softmax warps of tile t
  uint64_t* o_done = p_full + 2;       // [2]  last PV_t complete
  uint64_t* pv_done = o_done + 2;      // [2]  OVG_ATT3_SAFE_ORDER only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int head = blockIdx.y;
  const int bh = blockIdx.z * p.heads + head;
  const int nkv = (p.n + 63) / 64;
  const bool two = (q0 + 128) < p.n;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("ovg attn3: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_done[i], 1);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], two ? 2 : 1);   // one commit per MMA issuer (tile)
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], two ? 2 : 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // register budget: launch allocation 80 * 384 = 30720 per CTA; after the split 128 * 32 + 256 * 104 = 30720
  if (warp < 4) {
    reg_dealloc<32>();
    if (warp == 0) {
      if (lane == 0) {
        mbar_expect_tx(&q_full[0], ATT_TILE_BYTES);
        tma_load_3d(sQ, &tmQ, &q_full[0], 0, q0, bh);
        if (two) {
          mbar_expect_tx(&q_full[1], ATT_TILE_BYTES);
          tma_load_3d(sQ + ATT_TILE_BYTES, &tmQ, &q_full[1], 0, q0 + 128, bh);
        }
        int s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < nkv; ++j) {
          mbar_wait_quiet(&k_empty[s], ph ^ 1);
          mbar_expect_tx(&k_full[s], ATT3_KV_BYTES);
          tma_load_3d(sK + s * ATT3_KV_BYTES, &tmK, &k_full[s], 0, j * 64, bh);
          mbar_wait_quiet(&v_empty[s], ph ^ 1);
          mbar_expect_tx(&v_full[s], ATT3_KV_BYTES);
          tma_load_3d(sV + s * ATT3_KV_BYTES, &tmV, &v_full[s], 0, j * 64, bh);
          if (++s == NS) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    } else if (warp <= 2) {
      const int t = warp - 1;
      if (lane == 0 && (t == 0 || two)) {
        constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
        const uint32_t tS = tmem_base + t * 128;                      // P aliases the first 32 columns of S
        const uint32_t tO = tmem_base + t * 128 + 64;
        const uint64_t qdesc = make_sw128_desc(smem_u32(sQ + t * ATT_TILE_BYTES));
        const uint64_t kdesc0 = make_sw128_desc(smem_u32(sK));
        const uint64_t vdesc0 = make_sw128_desc(smem_u32(sV));
        constexpr uint64_t kStageStep = ATT3_KV_BYTES >> 4;           // descriptor address units are 16 B
        mbar_wait_quiet(&q_full[t], 0);
        mbar_wait_quiet(&k_full[0], 0);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS, qdesc + 2 * k, kdesc0 + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[0]);
        int s = 0, sn = 1;
        uint32_t ph = 0, phn = 0;
        for (int j = 0; j < nkv; ++j) {
          mbar_wait_quiet(&v_full[s], ph);
          mbar_wait_quiet(&p_full[t], j & 1);
          tc_fence_after();
          {
            const uint64_t bdesc = vdesc0 + s * kStageStep;
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 16 keys per MMA: P advances 8 columns (bf16 pairs), V advances 16 rows = 2048 B
              umma_ts(tO, tS + 8 * k, bdesc + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&v_empty[s]);
          if (j + 1 < nkv) {
#if OVG_ATT3_SAFE_ORDER
            umma_commit(&pv_done[t]);
            mbar_wait_quiet(&pv_done[t], j & 1);
            tc_fence_after();
#endif
            mbar_wait_quiet(&k_full[sn], phn);
            tc_fence_after();
            const uint64_t bdesc = kdesc0 + sn * kStageStep;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tS, qdesc + 2 * k, bdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
            umma_commit(&s_full[t]);
            umma_commit(&k_empty[sn]);
          } else {
            umma_commit(&o_done[t]);
          }
          s = sn;
          ph = phn;
          if (++sn == NS) {
            sn = 0;
            phn ^= 1;
          }
        }
      }
    }
  } else {
    reg_alloc<104>();
    const int t = (warp - 4) >> 2;
    if (t == 0 || two) {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;
      const int qrow = q0 + t * 128 + r;
      const uint32_t tS = tmem_base + t * 128 + (static_cast<uint32_t>(quarter * 32) << 16);
      const uint32_t tO = tS + 64;
      float m_ref = -INFINITY;   // integer-valued once set: rescaling factors are exact powers of two
      float l = 0.f;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait_quiet(&s_full[t], j & 1);
        tc_fence_after();
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t raw[32];
          tmem_ld32(tS + 32 * c, raw);
          tmem_ld_wait();
          const int kv_left = p.n - j * 64 - c * 32;     // valid keys in this chunk (TMA zero-fills the rest)
          if (kv_left < 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= kv_left) raw[i] = 0xff800000u;    // -inf: excluded from the max, exp2 -> 0
          }
          uint32_t pk[16];
          float2 a2 = make_float2(0.f, 0.f);
          bool slow = (j == 0 && c == 0);
          float mx = -INFINITY;
          if (!slow) {
            const float2 negm = make_float2(-m_ref, -m_ref);
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float r0 = __uint_as_float(raw[2 * i]), r1 = __uint_as_float(raw[2 * i + 1]);
              if (i & 1) mx1 = fmaxf(fmaxf(mx1, r0), r1); else mx0 = fmaxf(fmaxf(mx0, r0), r1);
              float2 x = fadd2(make_float2(r0, r1), negm);
              if (i >= 16 - OVG_ATT_EMU_PAIRS) {
                x = exp2_poly2(x);
              } else {
                x.x = ex2_approx(x.x);
                x.y = ex2_approx(x.y);
              }
              a2 = fadd2(a2, x);
              pk[i] = pack_bf16(x.x, x.y);
            }
            mx = fmaxf(mx0, mx1);
            slow = __any_sync(0xffffffffu, mx > m_ref + 8.0f);
          }
          if (slow) {
            if (j == 0 && c == 0) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
            }
            // Per-row decision inside a warp-uniform block (the TMEM accesses below are warp-collective).
            const bool need = mx > m_ref + 8.0f;
            float alpha = 1.0f;
            if (need) {
              const float m_new = ceilf(mx);
              const float d = m_ref - m_new;                       // negative integer, or -inf at the very first chunk
              alpha = d < -126.0f ? 0.0f : __int_as_float((127 + static_cast<int>(d)) << 23);
              m_ref = m_new;
              l *= alpha;
              acc.x *= alpha;
              acc.y *= alpha;
            }
            tmem_st_wait();                                        // chunk 0's P stores have landed
            if (j > 0) {                                           // O holds PV(0..j-1): stable until p_full(j) is signalled
#pragma unroll 1
              for (int h = 0; h < 4; ++h) {
                uint32_t o[16];
                tmem_ld16(tO + h * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st16(tO + h * 16, o);
              }
            }
            if (c == 1) {                                          // P of chunk 0 of this step, already stored: same factor
              uint32_t q16[16];
              tmem_ld16(tS, q16);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) q16[i] = pack_bf16(bf16_lo(q16[i]) * alpha, bf16_hi(q16[i]) * alpha);
              tmem_st16(tS, q16);
            }
            const float2 negm = make_float2(-m_ref, -m_ref);
            a2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float2 x = fadd2(make_float2(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1])), negm);
              x.x = ex2_approx(x.x);
              x.y = ex2_approx(x.y);
              a2 = fadd2(a2, x);
              pk[i] = pack_bf16(x.x, x.y);
            }
          }
          tmem_st16(tS + 16 * c, pk);
          acc = fadd2(acc, a2);
        }
        l += acc.x + acc.y;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ---- epilogue: O / l -> bf16 -> out[b, qrow, head*64 .. +64)
      mbar_wait_quiet(&o_done[t], 0);
      tc_fence_after();
      const float inv = 1.0f / l;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t o[32];
        tmem_ld32(tO + h * 32, o);
        tmem_ld_wait();
        if (qrow < p.n) {
          uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<long long>(blockIdx.z) * p.n + qrow) * p.C + head * 64 +
                                                h * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
            w.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
            w.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
            w.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
            dst[i] = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace ovg
