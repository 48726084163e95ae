// Spatial self-attention of the UNet transformer blocks (K6 in SURVEY §2.3): flash-attention forward,
// head dim 64, fp16 operands, on tcgen05 tensor cores with TMEM-resident score tiles.
//   diffusers attention_processor.py AttnProcessor2_0 as used by BasicTransformerBlock.attn1:
//   softmax(Q K^T / sqrt(64)) V per (frame, head) over S = h*w tokens.
//
// Kernel generations in this file (all behind evw_spatial_attention_f16; evw_set_attention_variant / EVW_ATTN_* pick one,
// tests/test_gpu_unet_ops.py::test_spatial_attention_variants checks every one against the same fp32 reference):
//   v1  spatial_attn_kernel    128 queries per CTA, 4 softmax warps, P through shared memory            (EVW_ATTN_V1)
//   v2  spatial_attn2_kernel   two 128-query groups per CTA sharing each K/V tile, setmaxnreg            (EVW_ATTN_V2)
//   v3  spatial_attn3_kernel   O and row sums accumulate in tensor memory, lazy rescale                  (variant -1)
//   v4  spatial_attn4_kernel   P aliased onto S in tensor memory (serialised S/PV chain: slower)         (EVW_ATTN_V4)
//   v5  spatial_attn5_kernel   v3 + the two groups take turns on the MUFU pipe (named-barrier token)     (variants 0-4)
//   v6  spatial_attn6_kernel   v3 + two threads per query row (16 softmax warps)                         (variants 5-8)
//   v7  spatial_attn7_kernel   P in its OWN tensor-memory columns, A operand of PV from TMEM, CUDA-core row sums (9-13)
//   v8  spatial_attn8_kernel   v7 + two threads per query row — the DEFAULT (variant 14)
// Measured at 28 frames x 9216 tokens x 5 heads (profiles/r01*_attn_bench*.log): v3 4.99 ms, v5 4.66, v6 4.62, v7 3.87,
// v8 3.69 ms.  What moved the needle was taking P out of shared memory (v3-v6 saturate the shared-memory data pipe:
// P written once and read twice per tile); v8 is bound by the MUFU pipe (77 % busy).
//
// Common structure: warp 0 = TMA producer (Q once, then a ring of (K_j, V_j) 128x64 tiles from the fused qkv buffer
// [F*S, 3C] through a 3-D tensor map; rows past S are zero-filled), warps 1-2 = MMA issuers (S_j = Q K_j^T, M128 N128 K64;
// O += P_j V_j, M128 N64 K128), remaining warps = softmax (one or two threads per query row: tcgen05.ld S_j -> running
// max in the log2 domain -> P_j in fp16).
#include "common.h"
#include "tc_common.cuh"
#include "unet_elem.h"
#include <cstdlib>

namespace evw {
using namespace tc;
namespace {

constexpr int kBQ = 128, kBKV = 128, kD = 64;
constexpr int kKvStages = 3;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KiB: Q, K_j, V_j tiles and each half of P
constexpr int kAttnThreads = 192;
// smem map (offsets from the 1024-aligned base)
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kTileBytes;
constexpr int kOffV = kOffK + kKvStages * kTileBytes;
constexpr int kOffP = kOffV + kKvStages * kTileBytes;
constexpr int kOffBar = kOffP + 2 * 2 * kTileBytes;
constexpr int kAttnSmem = kOffBar + 256 + 1024;
// TMEM columns
constexpr int kTmemS = 0, kTmemO = 256;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
spatial_attn_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out, int S, int C, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, head = blockIdx.y, frame = blockIdx.z;
  const int q0 = q_tile * kBQ;
  const int n_kv = (S + kBKV - 1) / kBKV;

  const uint32_t bar = base + kOffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kKvStages + s); };
  auto s_full = [&](int b) { return bar + 8u * (1 + 2 * kKvStages + b); };
  auto s_empty = [&](int b) { return bar + 8u * (3 + 2 * kKvStages + b); };
  auto p_full = [&](int b) { return bar + 8u * (5 + 2 * kKvStages + b); };
  auto o_full = [&](int b) { return bar + 8u * (7 + 2 * kKvStages + b); };
  const uint32_t tmem_slot = bar + 8u * (9 + 2 * kKvStages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kOffBar + 8 * (9 + 2 * kKvStages));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), 4);
      mbar_init(p_full(b), 4);
      mbar_init(o_full(b), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_3d(&tmap, base + kOffQ, q_full, head * kD, q0, frame);
    }
    for (int j = 0; j < n_kv; ++j) {
      const int st = j % kKvStages;
      mbar_wait(kv_empty(st), ((j / kKvStages) & 1) ^ 1u);
      if (lane == 0) {
        mbar_arrive_expect_tx(kv_full(st), 2 * kTileBytes);
        tma_load_3d(&tmap, base + kOffK + st * kTileBytes, kv_full(st), C + head * kD, j * kBKV, frame);
        tma_load_3d(&tmap, base + kOffV + st * kTileBytes, kv_full(st), 2 * C + head * kD, j * kBKV, frame);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0, 0);
    const uint32_t idesc_o = make_idesc_f16(kBQ, kD, 0, 0, 1);  // B (= V) is MN-major
    auto issue_s = [&](int i) {
      const int st = i % kKvStages, b = i & 1;
      mbar_wait(kv_full(st), (i / kKvStages) & 1);
      mbar_wait(s_empty(b), ((i >> 1) & 1) ^ 1u);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t dq = make_desc_k_sw128(base + kOffQ);
        const uint64_t dk = make_desc_k_sw128(base + kOffK + st * kTileBytes);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_f16_ss(tmem_base + kTmemS + b * kBKV, dq + 2ull * k, dk + 2ull * k, idesc_s, k != 0);
        tc_commit(s_full(b));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_s(j + 1);
      const int st = j % kKvStages, b = j & 1;
      mbar_wait(p_full(b), (j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int ks = 0; ks < kBKV / 16; ++ks) {
          const uint64_t dp = make_desc_k_sw128(base + kOffP + b * 2 * kTileBytes + (ks >> 2) * kTileBytes) + 2ull * (ks & 3);
          const uint64_t dv = make_desc_mn_sw128(base + kOffV + st * kTileBytes + ks * 2048, 1024);
          umma_f16_ss(tmem_base + kTmemO + b * kD, dp, dv, idesc_o, ks != 0);
        }
        tc_commit(o_full(b));
        tc_commit(kv_empty(st));
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / output (warps 2..5) =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
    float o_acc[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) o_acc[d] = 0.f;

    auto fold_o = [&](int i, float alpha) {
      const int b = i & 1;
      mbar_wait(o_full(b), (i >> 1) & 1);
      tc_fence_after();
      uint32_t lo[32], hi[32];
      tmem_ld_32x32b_x32(lane_addr + kTmemO + b * kD, lo);
      tmem_ld_32x32b_x32(lane_addr + kTmemO + b * kD + 32, hi);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        o_acc[d] = fmaf(o_acc[d], alpha, __uint_as_float(lo[d]));
        o_acc[32 + d] = fmaf(o_acc[32 + d], alpha, __uint_as_float(hi[d]));
      }
    };

    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      mbar_wait(s_full(b), (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[128];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
        uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
        tmem_ld_32x32b_x32(lane_addr + kTmemS + b * kBKV + 0, s0);
        tmem_ld_32x32b_x32(lane_addr + kTmemS + b * kBKV + 32, s1);
        tmem_ld_32x32b_x32(lane_addr + kTmemS + b * kBKV + 64, s2);
        tmem_ld_32x32b_x32(lane_addr + kTmemS + b * kBKV + 96, s3);
        tmem_ld_wait();
      }
      // the score tile now lives in registers: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(b));

      const int kv_valid = S - j * kBKV;  // >= 1; < 128 only in the last tile
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 128; ++c) {
        float v = __uint_as_float(s[c]) * scale_log2e;
        if (c >= kv_valid) v = -INFINITY;
        s[c] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = fast_exp2(m_run - m_new);  // first tile: exp2(-inf) = 0
      float sum = 0.f;
      uint8_t* prow = base_ptr + kOffP + b * 2 * kTileBytes + r * 128;
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {  // 16-byte chunks of 8 probabilities
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p0 = fast_exp2(__uint_as_float(s[ch * 8 + 2 * i]) - m_new);
          const float p1 = fast_exp2(__uint_as_float(s[ch * 8 + 2 * i + 1]) - m_new);
          sum += p0 + p1;
          __half2 h = __floats2half2_rn(p0, p1);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        const int atom = ch >> 3, cc = ch & 7;
        *reinterpret_cast<uint4*>(prow + atom * kTileBytes + ((cc ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      l_run = l_run * alpha + sum;
      m_run = m_new;
      fence_proxy_async();  // make the generic-proxy smem writes visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(b));
      if (j > 0) fold_o(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    fold_o(n_kv - 1, alpha_prev);
    const int qrow = q0 + r;
    if (qrow < S) {
      const float inv = 1.0f / l_run;
      uint4* op = reinterpret_cast<uint4*>(out + ((long long)frame * S + qrow) * C + head * kD);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __half2 h0 = __floats2half2_rn(o_acc[8 * i + 0] * inv, o_acc[8 * i + 1] * inv);
        __half2 h1 = __floats2half2_rn(o_acc[8 * i + 2] * inv, o_acc[8 * i + 3] * inv);
        __half2 h2 = __floats2half2_rn(o_acc[8 * i + 4] * inv, o_acc[8 * i + 5] * inv);
        __half2 h3 = __floats2half2_rn(o_acc[8 * i + 6] * inv, o_acc[8 * i + 7] * inv);
        uint4 v;
        v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
        v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
        op[i] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ==========================================================================================
// v2: two 128-query tiles per CTA (two softmax warpgroups ping-pong on the MUFU pipe and share every
// K/V tile), single-buffered S/O in TMEM, P through shared memory, event-driven MMA issue.
//   warp 0 TMA | warp 1 MMA | warps 4..7 softmax group A (rows q0..q0+127) | warps 8..11 group B (+128)
// Registers are rebalanced with setmaxnreg: the control warpgroup drops to 40, the softmax warpgroups grow to 232.
// ==========================================================================================
constexpr int kA2Threads = 384;  // warpgroup 0: TMA + MMA (+2 idle warps), warpgroups 1, 2: softmax groups A, B
constexpr int kA2OffQ = 0;                                   // 2 x 16 KiB
constexpr int kA2OffK = kA2OffQ + 2 * kTileBytes;            // 3 x 16 KiB
constexpr int kA2OffV = kA2OffK + kKvStages * kTileBytes;    // 3 x 16 KiB
constexpr int kA2OffP = kA2OffV + kKvStages * kTileBytes;    // 2 groups x 32 KiB
constexpr int kA2OffBar = kA2OffP + 2 * 2 * kTileBytes;
constexpr int kA2Smem = kA2OffBar + 256 + 1024;
#ifndef EVW_POLY_EVERY
#define EVW_POLY_EVERY 4
#endif
constexpr int kPolyEvery = EVW_POLY_EVERY;  // every n-th exponential is evaluated on the FMA pipe instead of MUFU

// 2^x for x <= 0 on the FMA/ALU pipes: round-to-nearest split x = n + r, |r| <= 0.5, degree-4 polynomial for 2^r
// (relative error < 5e-5, below the fp16 rounding of P), exponent patched in with an integer add.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;  // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float n = t - 12582912.0f;
  const float r = x - n;
  float p = fmaf(9.618129e-3f, r, 5.550411e-2f);
  p = fmaf(p, r, 2.402265e-1f);
  p = fmaf(p, r, 6.931472e-1f);
  p = fmaf(p, r, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__global__ void __launch_bounds__(kA2Threads, 1)
spatial_attn2_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out, int S, int C, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, frame = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kBQ;
  const int n_kv = (S + kBKV - 1) / kBKV;
  const bool b_active = q0 + kBQ < S;

  const uint32_t bar = base + kA2OffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kKvStages + s); };
  auto s_full = [&](int g) { return bar + 8u * (1 + 2 * kKvStages + g); };
  auto s_empty = [&](int g) { return bar + 8u * (3 + 2 * kKvStages + g); };
  auto p_full = [&](int g) { return bar + 8u * (5 + 2 * kKvStages + g); };
  auto o_full = [&](int g) { return bar + 8u * (7 + 2 * kKvStages + g); };
  const uint32_t tmem_slot = bar + 8u * (9 + 2 * kKvStages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kA2OffBar + 8 * (9 + 2 * kKvStages));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_empty(g), 4);
      mbar_init(p_full(g), 4);
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    // control warpgroup: give registers back, the softmax warpgroups take them (register budgets are per branch,
    // so the setmaxnreg has to dominate the whole role body)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      tma_load_3d(&tmap, base + kA2OffQ, q_full, head * kD, q0, frame);
      tma_load_3d(&tmap, base + kA2OffQ + kTileBytes, q_full, head * kD, q0 + kBQ, frame);
    }
    for (int j = 0; j < n_kv; ++j) {
      const int st = j % kKvStages;
      mbar_wait(kv_empty(st), ((j / kKvStages) & 1) ^ 1u);
      if (lane == 0) {
        mbar_arrive_expect_tx(kv_full(st), 2 * kTileBytes);
        tma_load_3d(&tmap, base + kA2OffK + st * kTileBytes, kv_full(st), C + head * kD, j * kBKV, frame);
        tma_load_3d(&tmap, base + kA2OffV + st * kTileBytes, kv_full(st), 2 * C + head * kD, j * kBKV, frame);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (event driven) =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(kBQ, kD, 0, 0, 1);
      mbar_wait(q_full, 0);
      int s_next[2] = {0, 0}, pv_next[2] = {0, 0};
      const int n_g[2] = {n_kv, b_active ? n_kv : 0};
      long long t0 = clock64();
      while (pv_next[0] < n_g[0] || pv_next[1] < n_g[1]) {
        bool progress = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          // ---- S_g(i) = Q_g K_i^T : needs K_i in smem and the previous S_g read out by the softmax group
          int i = s_next[g];
          if (i < n_g[g] && mbar_test(kv_full(i % kKvStages), (i / kKvStages) & 1) &&
              (i == 0 || mbar_test(s_empty(g), (i - 1) & 1))) {
            tc_fence_after();
            const uint64_t dq = make_desc_k_sw128(base + kA2OffQ + g * kTileBytes);
            const uint64_t dk = make_desc_k_sw128(base + kA2OffK + (i % kKvStages) * kTileBytes);
#pragma unroll
            for (int k = 0; k < kD / 16; ++k)
              umma_f16_ss(tmem_base + g * kBKV, dq + 2ull * k, dk + 2ull * k, idesc_s, k != 0);
            tc_commit(s_full(g));
            s_next[g] = i + 1;
            progress = true;
          }
          // ---- O_g(i) = P_g(i) V_i : needs P_g(i) in smem (which also implies O_g(i-1) was folded)
          i = pv_next[g];
          if (i < n_g[g] && i < s_next[g] && mbar_test(p_full(g), i & 1)) {
            tc_fence_after();
            const int st = i % kKvStages;
#pragma unroll
            for (int ks = 0; ks < kBKV / 16; ++ks) {
              const uint64_t dp = make_desc_k_sw128(base + kA2OffP + g * 2 * kTileBytes + (ks >> 2) * kTileBytes) + 2ull * (ks & 3);
              const uint64_t dv = make_desc_mn_sw128(base + kA2OffV + st * kTileBytes + ks * 2048, 1024);
              umma_f16_ss(tmem_base + 256 + g * kD, dp, dv, idesc_o, ks != 0);
            }
            tc_commit(o_full(g));
            pv_next[g] = i + 1;
            // K_i / V_i are free once both groups have issued their PV for tile i
            const int other = g ^ 1;
            if (n_g[other] == 0 || pv_next[other] > i) tc_commit(kv_empty(st));
            progress = true;
          }
        }
        if (progress) t0 = clock64();
        else if (clock64() - t0 > 8000000000ll) __trap();
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===================== softmax groups =====================
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    if (g == 0 || b_active) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t s_addr = lane_addr + g * kBKV, o_addr = lane_addr + 256 + g * kD;
      float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;
      float o_acc[kD];
#pragma unroll
      for (int d = 0; d < kD; ++d) o_acc[d] = 0.f;
      uint8_t* prow = base_ptr + kA2OffP + g * 2 * kTileBytes + r * 128;

      auto fold_o = [&](int i, float alpha) {
        mbar_wait(o_full(g), i & 1);
        tc_fence_after();
        uint32_t lo[32], hi[32];
        tmem_ld_32x32b_x32(o_addr, lo);
        tmem_ld_32x32b_x32(o_addr + 32, hi);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          o_acc[d] = fmaf(o_acc[d], alpha, __uint_as_float(lo[d]));
          o_acc[32 + d] = fmaf(o_acc[32 + d], alpha, __uint_as_float(hi[d]));
        }
      };

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full(g), j & 1);
        tc_fence_after();
        uint32_t s[128];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
          uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
          tmem_ld_32x32b_x32(s_addr + 0, s0);
          tmem_ld_32x32b_x32(s_addr + 32, s1);
          tmem_ld_32x32b_x32(s_addr + 64, s2);
          tmem_ld_32x32b_x32(s_addr + 96, s3);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(g));  // S_g(j+1) may now overwrite the TMEM tile
        // fold the previous PV result: also proves PV_g(j-1) is done, i.e. the P buffer is free again
        if (j > 0) fold_o(j - 1, alpha_prev);

        const int kv_valid = S - j * kBKV;
        float mx = -INFINITY;
        if (kv_valid >= kBKV) {
#pragma unroll
          for (int c = 0; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            if (c >= kv_valid) s[c] = 0xff800000u;  // -inf
            mx = fmaxf(mx, __uint_as_float(s[c]));
          }
        }
        const float m_new = fmaxf(m_run, mx * scale_log2e);
        const float alpha = fast_exp2(m_run - m_new);
        const float neg_m = -m_new;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c0 = ch * 8 + 2 * i;
            const float x0 = fmaf(__uint_as_float(s[c0]), scale_log2e, neg_m);
            const float x1 = fmaf(__uint_as_float(s[c0 + 1]), scale_log2e, neg_m);
            const float p0 = fast_exp2(x0);
            const float p1 = ((c0 + 1) % kPolyEvery == kPolyEvery - 1) ? exp2_poly(x1) : fast_exp2(x1);
            sum0 += p0;
            sum1 += p1;
            __half2 h = __floats2half2_rn(p0, p1);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
          }
          const int atom = ch >> 3, cc = ch & 7;
          *reinterpret_cast<uint4*>(prow + atom * kTileBytes + ((cc ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        l_run = l_run * alpha + (sum0 + sum1);
        m_run = m_new;
        alpha_prev = alpha;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      fold_o(n_kv - 1, alpha_prev);
      if (qrow < S) {
        const float inv = 1.0f / l_run;
        uint4* op = reinterpret_cast<uint4*>(out + ((long long)frame * S + qrow) * C + head * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __half2 h0 = __floats2half2_rn(o_acc[8 * i + 0] * inv, o_acc[8 * i + 1] * inv);
          __half2 h1 = __floats2half2_rn(o_acc[8 * i + 2] * inv, o_acc[8 * i + 3] * inv);
          __half2 h2 = __floats2half2_rn(o_acc[8 * i + 4] * inv, o_acc[8 * i + 5] * inv);
          __half2 h3 = __floats2half2_rn(o_acc[8 * i + 6] * inv, o_acc[8 * i + 7] * inv);
          uint4 v;
          v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
          v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
          op[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// v3: as v2, plus (a) O and the softmax denominator accumulate in TENSOR MEMORY across all K/V tiles — the
// denominator comes from a second tiny MMA of P against an all-ones [16 x 128] operand, so the CUDA cores neither
// sum probabilities nor fold partial outputs; (b) lazy rescaling: the exponent reference m_ref of a row only moves
// when the tile maximum exceeds it by more than 2^8, and only then are O and L rescaled in place (tcgen05.ld/st);
// P = exp2(s c - m_ref) <= 256 stays well inside fp16.  Per score element the softmax warps issue ~3 instructions
// (FMNMX3/2, FFMA, MUFU|poly, F2FP/2) instead of ~10.
// ==========================================================================================
constexpr int kA3OffOnes = kA2OffBar + 256;                  // 4 KiB of fp16 1.0 (K-major [16 x 128], swizzle-invariant)
constexpr int kA3Smem = kA3OffOnes + 4096 + 1024;
constexpr int kTmemO3 = 256, kTmemL3 = 384;                  // O_g at 256 + 64 g, L_g at 384 + 16 g
constexpr float kLazyTau = 8.0f;

__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t units16) {  // advance the 14-bit start address field
  return d + units16;  // never carries out of the address field for in-range tiles
}

#ifndef EVW_ATTN_ONES_MMA
#define EVW_ATTN_ONES_MMA 1
#endif
#ifndef EVW_EXP_F16X2
#define EVW_EXP_F16X2 0
#endif
// two exponentials per MUFU op: the arguments (<= 0) are rounded to fp16 first — the resulting relative error of p is
// <= 0.07% for p >= 1/16 and shrinks in absolute terms below that, under the fp16 rounding P gets anyway
__device__ __forceinline__ uint32_t exp2_pair_f16x2(float x0, float x1) {
  __half2 h = __floats2half2_rn(x0, x1);
  uint32_t in = *reinterpret_cast<uint32_t*>(&h), o;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(o) : "r"(in));
  return o;
}

__global__ void __launch_bounds__(kA2Threads, 1)
spatial_attn3_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out, int S, int C, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, frame = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kBQ;
  const int n_kv = (S + kBKV - 1) / kBKV;
  const bool b_active = q0 + kBQ < S;

  const uint32_t bar = base + kA2OffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kKvStages + s); };
  auto s_full = [&](int g) { return bar + 8u * (1 + 2 * kKvStages + g); };
  auto s_empty = [&](int g) { return bar + 8u * (3 + 2 * kKvStages + g); };
  auto p_full = [&](int g) { return bar + 8u * (5 + 2 * kKvStages + g); };
  auto o_full = [&](int g) { return bar + 8u * (7 + 2 * kKvStages + g); };
  const uint32_t tmem_slot = bar + 8u * (9 + 2 * kKvStages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kA2OffBar + 8 * (9 + 2 * kKvStages));

  // all-ones operand for the row-sum MMA
  for (int i = threadIdx.x; i < 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base_ptr + kA3OffOnes)[i] = 0x3C003C00u;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), b_active ? 2 : 1);  // one tcgen05.commit per active group
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_empty(g), 4);
      mbar_init(p_full(g), 4);
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
  }
  fence_proxy_async();  // ones tile + barrier inits visible to the async proxy
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        tma_load_3d(&tmap, base + kA2OffQ, q_full, head * kD, q0, frame);
        tma_load_3d(&tmap, base + kA2OffQ + kTileBytes, q_full, head * kD, q0 + kBQ, frame);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kKvStages;
        mbar_wait(kv_empty(st), ((j / kKvStages) & 1) ^ 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(kv_full(st), 2 * kTileBytes);
          tma_load_3d(&tmap, base + kA2OffK + st * kTileBytes, kv_full(st), C + head * kD, j * kBKV, frame);
          tma_load_3d(&tmap, base + kA2OffV + st * kTileBytes, kv_full(st), 2 * C + head * kD, j * kBKV, frame);
        }
        __syncwarp();
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: warp 1 drives group A, warp 2 drives group B (event driven) =====================
      const int g = warp - 1;
      if (lane == 0 && (g == 0 || b_active)) {
        const uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0, 0);
        const uint32_t idesc_o = make_idesc_f16(kBQ, kD, 0, 0, 1);
        const uint32_t idesc_l = make_idesc_f16(kBQ, 16, 0, 0, 0);
        // loop-invariant descriptors; per-k-step offsets are compile-time constants added to the address field
        const uint64_t dq = make_desc_k_sw128(base + kA2OffQ + g * kTileBytes);
        const uint64_t dp = make_desc_k_sw128(base + kA2OffP + g * 2 * kTileBytes);
        const uint64_t d1 = make_desc_k_sw128(base + kA3OffOnes);
        const uint64_t dk0 = make_desc_k_sw128(base + kA2OffK);
        const uint64_t dv0 = make_desc_mn_sw128(base + kA2OffV, 1024);
        const uint32_t t_s = tmem_base + g * kBKV, t_o = tmem_base + kTmemO3 + g * kD, t_l = tmem_base + kTmemL3 + g * 16;
        mbar_wait(q_full, 0);
        int s_next = 0, pv_next = 0;
        long long t0 = clock64();
        while (pv_next < n_kv) {
          bool progress = false;
          if (s_next < n_kv && mbar_test(kv_full(s_next % kKvStages), (s_next / kKvStages) & 1) &&
              (s_next == 0 || mbar_test(s_empty(g), (s_next - 1) & 1))) {
            tc_fence_after();
            const uint64_t dk = desc_add(dk0, (s_next % kKvStages) * (kTileBytes >> 4));
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) umma_f16_ss(t_s, desc_add(dq, 2 * k), desc_add(dk, 2 * k), idesc_s, k != 0);
            tc_commit(s_full(g));
            ++s_next;
            progress = true;
          }
          if (pv_next < s_next && mbar_test(p_full(g), pv_next & 1)) {
            tc_fence_after();
            const int st = pv_next % kKvStages;
            const uint64_t dv = desc_add(dv0, st * (kTileBytes >> 4));
            const uint32_t acc = pv_next != 0;
#pragma unroll
            for (int ks = 0; ks < kBKV / 16; ++ks) {
              const uint64_t a = desc_add(dp, (ks >> 2) * (kTileBytes >> 4) + 2 * (ks & 3));
              umma_f16_ss(t_o, a, desc_add(dv, ks * (2048 >> 4)), idesc_o, acc | (ks != 0));                         // O += P V
#if EVW_ATTN_ONES_MMA
              umma_f16_ss(t_l, a, desc_add(d1, (ks >> 2) * (2048 >> 4) + 2 * (ks & 3)), idesc_l, acc | (ks != 0));   // L += P 1
#endif
            }
            tc_commit(o_full(g));
            tc_commit(kv_empty(st));
            ++pv_next;
            progress = true;
          }
          if (progress) t0 = clock64();
          else if (clock64() - t0 > 8000000000ll) __trap();
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== softmax groups =====================
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    if (g == 0 || b_active) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t s_addr = lane_addr + g * kBKV, o_addr = lane_addr + kTmemO3 + g * kD, l_addr = lane_addr + kTmemL3 + g * 16;
      float m_ref = -INFINITY;
      float l_run = 0.f;
      uint8_t* prow = base_ptr + kA2OffP + g * 2 * kTileBytes + r * 128;

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full(g), j & 1);
        tc_fence_after();
        uint32_t s[128];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
          uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
          tmem_ld_32x32b_x32(s_addr + 0, s0);
          tmem_ld_32x32b_x32(s_addr + 32, s1);
          tmem_ld_32x32b_x32(s_addr + 64, s2);
          tmem_ld_32x32b_x32(s_addr + 96, s3);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(g));  // S_g(j+1) may now overwrite the TMEM tile

        const int kv_valid = S - j * kBKV;
        float mx = -INFINITY;
        if (kv_valid >= kBKV) {
#pragma unroll
          for (int c = 0; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            if (c >= kv_valid) s[c] = 0xff800000u;
            mx = fmaxf(mx, __uint_as_float(s[c]));
          }
        }
        const float m_tile = mx * scale_log2e;
        const bool need = m_tile > m_ref + kLazyTau;  // first tile: m_ref = -inf -> true
        const float m_old = m_ref;
        if (need) m_ref = m_tile;
        const float neg_m = -m_ref;
        // probabilities first (registers only): this phase overlaps the PV MMA of the previous tile
        uint32_t w[64];
#if !EVW_ATTN_ONES_MMA
        float sum0 = 0.f, sum1 = 0.f;
#endif
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float x0 = fmaf(__uint_as_float(s[2 * i]), scale_log2e, neg_m);
          const float x1 = fmaf(__uint_as_float(s[2 * i + 1]), scale_log2e, neg_m);
#if EVW_EXP_F16X2
          w[i] = exp2_pair_f16x2(fmaxf(x0, -60000.0f), fmaxf(x1, -60000.0f));  // -inf (masked) -> 0
#else
          const float p0 = fast_exp2(x0);
          const float p1 = ((2 * i + 1) % kPolyEvery == kPolyEvery - 1) ? exp2_poly(x1) : fast_exp2(x1);
          __half2 h = __floats2half2_rn(p0, p1);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
#if !EVW_ATTN_ONES_MMA
          sum0 += p0;
          sum1 += p1;
#endif
#endif
        }
#if !EVW_ATTN_ONES_MMA
        l_run = l_run * (need ? fast_exp2(m_old - m_ref) : 1.0f) + (sum0 + sum1);
#endif
        // P buffer / O, L accumulators of this group are free once PV_g(j-1) has completed
        if (j > 0) {
          mbar_wait(o_full(g), (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? fast_exp2(m_old - m_ref) : 1.0f;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
              uint32_t v[16];
              tmem_ld_32x32b_x16(o_addr + part * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st_32x32b_x16(o_addr + part * 16, v);
            }
#if EVW_ATTN_ONES_MMA
            {
              uint32_t v[16];
              tmem_ld_32x32b_x16(l_addr, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st_32x32b_x16(l_addr, v);
            }
#endif
            tmem_st_wait();
          }
        }
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          const int atom = ch >> 3, cc = ch & 7;
          *reinterpret_cast<uint4*>(prow + atom * kTileBytes + ((cc ^ (r & 7)) << 4)) =
              make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      // final: O / L
      mbar_wait(o_full(g), (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t lo[32], hi[32], lv[16];
      tmem_ld_32x32b_x32(o_addr, lo);
      tmem_ld_32x32b_x32(o_addr + 32, hi);
      tmem_ld_32x32b_x16(l_addr, lv);
      tmem_ld_wait();
#if !EVW_ATTN_ONES_MMA
      lv[0] = __float_as_uint(l_run);
#endif
      if (qrow < S) {
        const float inv = 1.0f / __uint_as_float(lv[0]);
        uint4* op = reinterpret_cast<uint4*>(out + ((long long)frame * S + qrow) * C + head * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t* src = (i < 4) ? &lo[8 * i] : &hi[8 * (i - 4)];
          __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          uint4 v;
          v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
          v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
          op[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// v4: as v3, but the probability tile P never touches shared memory: the softmax warps write it (fp16, two per
// 32-bit column) over the first 64 columns of their S tile with tcgen05.st and the PV / row-sum MMAs take their A
// operand from TENSOR MEMORY.  Shared memory then only carries Q, K, V (v3 was bound by smem bandwidth: P written,
// read by the PV MMA and read again by the row-sum MMA).  S_g(j+1) is issued right behind PV_g(j) in the same
// in-order MMA stream, so it may overwrite the aliased S/P columns; the two query groups keep the tensor pipe busy.
// ==========================================================================================
constexpr int kA4Stages = 4;
constexpr int kA4OffQ = 0;
constexpr int kA4OffK = kA4OffQ + 2 * kTileBytes;
constexpr int kA4OffV = kA4OffK + kA4Stages * kTileBytes;
constexpr int kA4OffOnes = kA4OffV + kA4Stages * kTileBytes;
constexpr int kA4OffBar = kA4OffOnes + 4096;
constexpr int kA4Smem = kA4OffBar + 256 + 1024;

__global__ void __launch_bounds__(kA2Threads, 1)
spatial_attn4_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out, int S, int C, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, frame = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kBQ;
  const int n_kv = (S + kBKV - 1) / kBKV;
  const bool b_active = q0 + kBQ < S;

  const uint32_t bar = base + kA4OffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kA4Stages + s); };
  auto s_full = [&](int g) { return bar + 8u * (1 + 2 * kA4Stages + g); };
  auto p_full = [&](int g) { return bar + 8u * (3 + 2 * kA4Stages + g); };
  auto o_full = [&](int g) { return bar + 8u * (5 + 2 * kA4Stages + g); };
  const uint32_t tmem_slot = bar + 8u * (7 + 2 * kA4Stages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kA4OffBar + 8 * (7 + 2 * kA4Stages));

  for (int i = threadIdx.x; i < 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base_ptr + kA4OffOnes)[i] = 0x3C003C00u;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kA4Stages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), b_active ? 2 : 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 4);
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        tma_load_3d(&tmap, base + kA4OffQ, q_full, head * kD, q0, frame);
        tma_load_3d(&tmap, base + kA4OffQ + kTileBytes, q_full, head * kD, q0 + kBQ, frame);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kA4Stages;
        mbar_wait(kv_empty(st), ((j / kA4Stages) & 1) ^ 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(kv_full(st), 2 * kTileBytes);
          tma_load_3d(&tmap, base + kA4OffK + st * kTileBytes, kv_full(st), C + head * kD, j * kBKV, frame);
          tma_load_3d(&tmap, base + kA4OffV + st * kTileBytes, kv_full(st), 2 * C + head * kD, j * kBKV, frame);
        }
        __syncwarp();
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: one thread per query group =====================
      const int g = warp - 1;
      if (lane == 0 && (g == 0 || b_active)) {
        const uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0, 0);
        const uint32_t idesc_o = make_idesc_f16(kBQ, kD, 0, 0, 1);
        const uint32_t idesc_l = make_idesc_f16(kBQ, 16, 0, 0, 0);
        const uint64_t dq = make_desc_k_sw128(base + kA4OffQ + g * kTileBytes);
        const uint64_t d1 = make_desc_k_sw128(base + kA4OffOnes);
        const uint64_t dk0 = make_desc_k_sw128(base + kA4OffK);
        const uint64_t dv0 = make_desc_mn_sw128(base + kA4OffV, 1024);
        const uint32_t t_s = tmem_base + g * kBKV;  // S tile; P aliases its first 64 columns
        const uint32_t t_o = tmem_base + kTmemO3 + g * kD, t_l = tmem_base + kTmemL3 + g * 16;
        mbar_wait(q_full, 0);
        for (int i = 0; i < n_kv; ++i) {
          const int st = i % kA4Stages;
          mbar_wait(kv_full(st), (i / kA4Stages) & 1);
          tc_fence_after();
          const uint64_t dk = desc_add(dk0, st * (kTileBytes >> 4));
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) umma_f16_ss(t_s, desc_add(dq, 2 * k), desc_add(dk, 2 * k), idesc_s, k != 0);
          tc_commit(s_full(g));
          mbar_wait(p_full(g), i & 1);
          tc_fence_after();
          const uint64_t dv = desc_add(dv0, st * (kTileBytes >> 4));
          const uint32_t acc = i != 0;
#pragma unroll
          for (int ks = 0; ks < kBKV / 16; ++ks) {
            umma_f16_ts(t_o, t_s + ks * 8, desc_add(dv, ks * (2048 >> 4)), idesc_o, acc | (ks != 0));                       // O += P V
            umma_f16_ts(t_l, t_s + ks * 8, desc_add(d1, (ks >> 2) * (2048 >> 4) + 2 * (ks & 3)), idesc_l, acc | (ks != 0)); // L += P 1
          }
          tc_commit(kv_empty(st));
          if (i == n_kv - 1) tc_commit(o_full(g));
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== softmax groups =====================
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    if (g == 0 || b_active) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t s_addr = lane_addr + g * kBKV, o_addr = lane_addr + kTmemO3 + g * kD, l_addr = lane_addr + kTmemL3 + g * 16;
      float m_ref = -INFINITY;

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full(g), j & 1);  // also implies PV_g(j-1) and L_g(j-1) completed (in-order MMA stream)
        tc_fence_after();
        uint32_t s[128];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
          uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
          tmem_ld_32x32b_x32(s_addr + 0, s0);
          tmem_ld_32x32b_x32(s_addr + 32, s1);
          tmem_ld_32x32b_x32(s_addr + 64, s2);
          tmem_ld_32x32b_x32(s_addr + 96, s3);
          tmem_ld_wait();
        }
        const int kv_valid = S - j * kBKV;
        float mx = -INFINITY;
        if (kv_valid >= kBKV) {
#pragma unroll
          for (int c = 0; c < 128; ++c) mx = fmaxf(mx, __uint_as_float(s[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 128; ++c) {
            if (c >= kv_valid) s[c] = 0xff800000u;
            mx = fmaxf(mx, __uint_as_float(s[c]));
          }
        }
        const float m_tile = mx * scale_log2e;
        const bool need = m_tile > m_ref + kLazyTau;
        const float m_old = m_ref;
        if (need) m_ref = m_tile;
        const float neg_m = -m_ref;
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          const float f = need ? fast_exp2(m_old - m_ref) : 1.0f;
#pragma unroll
          for (int part = 0; part < 4; ++part) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(o_addr + part * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st_32x32b_x16(o_addr + part * 16, v);
          }
          {
            uint32_t v[16];
            tmem_ld_32x32b_x16(l_addr, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st_32x32b_x16(l_addr, v);
          }
        }
        // probabilities, packed two per 32-bit TMEM column (even kv index in the low half), written over S columns 0..63
#pragma unroll
        for (int part = 0; part < 4; ++part) {
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int c0 = part * 32 + 2 * i;
            const float x0 = fmaf(__uint_as_float(s[c0]), scale_log2e, neg_m);
            const float x1 = fmaf(__uint_as_float(s[c0 + 1]), scale_log2e, neg_m);
            const float p0 = fast_exp2(x0);
            const float p1 = ((c0 + 1) % kPolyEvery == kPolyEvery - 1) ? exp2_poly(x1) : fast_exp2(x1);
            __half2 h = __floats2half2_rn(p0, p1);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
          }
          tmem_st_32x32b_x16(s_addr + part * 16, w);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      mbar_wait(o_full(g), 0);
      tc_fence_after();
      uint32_t lo[32], hi[32], lv[16];
      tmem_ld_32x32b_x32(o_addr, lo);
      tmem_ld_32x32b_x32(o_addr + 32, hi);
      tmem_ld_32x32b_x16(l_addr, lv);
      tmem_ld_wait();
      if (qrow < S) {
        const float inv = 1.0f / __uint_as_float(lv[0]);
        uint4* op = reinterpret_cast<uint4*>(out + ((long long)frame * S + qrow) * C + head * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t* src = (i < 4) ? &lo[8 * i] : &hi[8 * (i - 4)];
          __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          uint4 v;
          v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
          v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
          op[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// v5: v3 with the two softmax groups explicitly out of phase.  ncu (profiles/r01_attn3_stalls.txt) showed v3's groups
// running in lockstep: both hit their 96 MUFU.EX2 per tile at the same time (pipe saturated, mio throttle) and then both
// did their FMA / TMEM / smem work with the MUFU pipe idle — xu 42 %, tensor 31 %, issue 46 % busy, nothing saturated.
// Here a token (two named barriers) serialises only the MUFU-dense phase between the groups, and the exponent
// arguments + polynomial exponentials are computed before the token is taken.
// ==========================================================================================
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_volatile(uint32_t addr, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
constexpr int kA5OffToken = kA3OffOnes + 4096;  // 256 words: one per softmax thread
constexpr int kA5Smem = kA5OffToken + 1024 + 1024;

template <int kPolyN, bool kStagger>
__global__ void __launch_bounds__(kA2Threads, 1)
spatial_attn5_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out, int S, int C, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, frame = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kBQ;
  const int n_kv = (S + kBKV - 1) / kBKV;
  const bool b_active = q0 + kBQ < S;

  const uint32_t bar = base + kA2OffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kKvStages + s); };
  auto s_full = [&](int g) { return bar + 8u * (1 + 2 * kKvStages + g); };
  auto s_empty = [&](int g) { return bar + 8u * (3 + 2 * kKvStages + g); };
  auto p_full = [&](int g) { return bar + 8u * (5 + 2 * kKvStages + g); };
  auto o_full = [&](int g) { return bar + 8u * (7 + 2 * kKvStages + g); };
  const uint32_t tmem_slot = bar + 8u * (9 + 2 * kKvStages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kA2OffBar + 8 * (9 + 2 * kKvStages));

  // all-ones operand for the row-sum MMA
  for (int i = threadIdx.x; i < 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base_ptr + kA3OffOnes)[i] = 0x3C003C00u;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), b_active ? 2 : 1);  // one tcgen05.commit per active group
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_empty(g), 4);
      mbar_init(p_full(g), 4);
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
  }
  fence_proxy_async();  // ones tile + barrier inits visible to the async proxy
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        tma_load_3d(&tmap, base + kA2OffQ, q_full, head * kD, q0, frame);
        tma_load_3d(&tmap, base + kA2OffQ + kTileBytes, q_full, head * kD, q0 + kBQ, frame);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kKvStages;
        mbar_wait(kv_empty(st), ((j / kKvStages) & 1) ^ 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(kv_full(st), 2 * kTileBytes);
          tma_load_3d(&tmap, base + kA2OffK + st * kTileBytes, kv_full(st), C + head * kD, j * kBKV, frame);
          tma_load_3d(&tmap, base + kA2OffV + st * kTileBytes, kv_full(st), 2 * C + head * kD, j * kBKV, frame);
        }
        __syncwarp();
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: warp 1 drives group A, warp 2 drives group B (event driven) =====================
      const int g = warp - 1;
      if (lane == 0 && (g == 0 || b_active)) {
        const uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0, 0);
        const uint32_t idesc_o = make_idesc_f16(kBQ, kD, 0, 0, 1);
        const uint32_t idesc_l = make_idesc_f16(kBQ, 16, 0, 0, 0);
        // loop-invariant descriptors; per-k-step offsets are compile-time constants added to the address field
        const uint64_t dq = make_desc_k_sw128(base + kA2OffQ + g * kTileBytes);
        const uint64_t dp = make_desc_k_sw128(base + kA2OffP + g * 2 * kTileBytes);
        const uint64_t d1 = make_desc_k_sw128(base + kA3OffOnes);
        const uint64_t dk0 = make_desc_k_sw128(base + kA2OffK);
        const uint64_t dv0 = make_desc_mn_sw128(base + kA2OffV, 1024);
        const uint32_t t_s = tmem_base + g * kBKV, t_o = tmem_base + kTmemO3 + g * kD, t_l = tmem_base + kTmemL3 + g * 16;
        mbar_wait(q_full, 0);
        int s_next = 0, pv_next = 0;
        long long t0 = clock64();
        while (pv_next < n_kv) {
          bool progress = false;
          if (s_next < n_kv && mbar_test(kv_full(s_next % kKvStages), (s_next / kKvStages) & 1) &&
              (s_next == 0 || mbar_test(s_empty(g), (s_next - 1) & 1))) {
            tc_fence_after();
            const uint64_t dk = desc_add(dk0, (s_next % kKvStages) * (kTileBytes >> 4));
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) umma_f16_ss(t_s, desc_add(dq, 2 * k), desc_add(dk, 2 * k), idesc_s, k != 0);
            tc_commit(s_full(g));
            ++s_next;
            progress = true;
          }
          if (pv_next < s_next && mbar_test(p_full(g), pv_next & 1)) {
            tc_fence_after();
            const int st = pv_next % kKvStages;
            const uint64_t dv = desc_add(dv0, st * (kTileBytes >> 4));
            const uint32_t acc = pv_next != 0;
#pragma unroll
            for (int ks = 0; ks < kBKV / 16; ++ks) {
              const uint64_t a = desc_add(dp, (ks >> 2) * (kTileBytes >> 4) + 2 * (ks & 3));
              umma_f16_ss(t_o, a, desc_add(dv, ks * (2048 >> 4)), idesc_o, acc | (ks != 0));                         // O += P V
#if EVW_ATTN_ONES_MMA
              umma_f16_ss(t_l, a, desc_add(d1, (ks >> 2) * (2048 >> 4) + 2 * (ks & 3)), idesc_l, acc | (ks != 0));   // L += P 1
#endif
            }
            tc_commit(o_full(g));
            tc_commit(kv_empty(st));
            ++pv_next;
            progress = true;
          }
          if (progress) t0 = clock64();
          else if (clock64() - t0 > 8000000000ll) __trap();
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== softmax groups =====================
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    if (g == 0 || b_active) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t s_addr = lane_addr + g * kBKV, o_addr = lane_addr + kTmemO3 + g * kD, l_addr = lane_addr + kTmemL3 + g * 16;
      float m_ref = -INFINITY;
      float l_run = 0.f;
      const uint32_t token_word = base + kA5OffToken + 4u * (threadIdx.x - 128);
      if (kStagger) st_shared_volatile(token_word, 0u);
      if (kStagger && b_active && g == 1) named_bar_arrive(1, 256);  // group A goes first
      uint8_t* prow = base_ptr + kA2OffP + g * 2 * kTileBytes + r * 128;

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full(g), j & 1);
        tc_fence_after();
        uint32_t s[128];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
          uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
          tmem_ld_32x32b_x32(s_addr + 0, s0);
          tmem_ld_32x32b_x32(s_addr + 32, s1);
          tmem_ld_32x32b_x32(s_addr + 64, s2);
          tmem_ld_32x32b_x32(s_addr + 96, s3);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(g));  // S_g(j+1) may now overwrite the TMEM tile

        const int kv_valid = S - j * kBKV;
        float mx = -INFINITY;
        if (kv_valid < kBKV) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= kv_valid) s[c] = 0xff800000u;
        }
        {  // four independent maxima: the serial FMNMX chain of v3 cost ~400 cycles per tile
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 128; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(s[c]));
          mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        }
        const float m_tile = mx * scale_log2e;
        const bool need = m_tile > m_ref + kLazyTau;  // first tile: m_ref = -inf -> true
        const float m_old = m_ref;
        if (need) m_ref = m_tile;
        const float neg_m = -m_ref;
        // (1) FMA-pipe phase, outside the MUFU token: the polynomial share of the exponentials.  It overlaps the other
        // group's MUFU phase on the same SM sub-partitions.
        if (kPolyN > 0) {
#pragma unroll
          for (int c = kPolyN - 1; c < 128; c += (kPolyN > 0 ? kPolyN : 128))
            s[c] = __float_as_uint(exp2_poly(fmaf(__uint_as_float(s[c]), scale_log2e, neg_m)));
        }
        // (2) MUFU phase: the two query groups take turns (named barriers 1 / 2), so one group's exponentials run while
        // the other group loads, scales, stores and waits — instead of both saturating the MUFU pipe in lockstep and
        // both leaving it idle afterwards.  ptxas only honours data dependencies, so the phase is pinned between the
        // barriers by one: its exponent arguments depend on a shared-memory word read after the acquire (always 0), and a
        // word derived from its results is stored before the release.
        float neg_m_dep = neg_m;
        if (kStagger && b_active) {
          named_bar_sync(1 + g, 256);
          neg_m_dep = neg_m + __uint_as_float(ld_shared_volatile(token_word));
        }
        uint32_t w[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const bool poly0 = kPolyN > 0 && ((2 * i) % (kPolyN > 0 ? kPolyN : 1)) == kPolyN - 1;
          const bool poly1 = kPolyN > 0 && ((2 * i + 1) % (kPolyN > 0 ? kPolyN : 1)) == kPolyN - 1;
          const float p0 = poly0 ? __uint_as_float(s[2 * i]) : fast_exp2(fmaf(__uint_as_float(s[2 * i]), scale_log2e, neg_m_dep));
          const float p1 = poly1 ? __uint_as_float(s[2 * i + 1])
                                 : fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2e, neg_m_dep));
          __half2 h = __floats2half2_rn(p0, p1);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (kStagger && b_active) {
          // probabilities are >= 0: the two fp16 sign bits are 0 at run time, which the compiler cannot know
          st_shared_volatile(token_word, (w[15] | w[31] | w[47] | w[63]) & 0x80008000u);
          if (!(g == 1 && j + 1 == n_kv)) named_bar_arrive(2 - g, 256);
        }
        // P buffer / O, L accumulators of this group are free once PV_g(j-1) has completed
        if (j > 0) {
          mbar_wait(o_full(g), (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
            const float f = need ? fast_exp2(m_old - m_ref) : 1.0f;
#pragma unroll
            for (int part = 0; part < 4; ++part) {
              uint32_t v[16];
              tmem_ld_32x32b_x16(o_addr + part * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st_32x32b_x16(o_addr + part * 16, v);
            }
#if EVW_ATTN_ONES_MMA
            {
              uint32_t v[16];
              tmem_ld_32x32b_x16(l_addr, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st_32x32b_x16(l_addr, v);
            }
#endif
            tmem_st_wait();
          }
        }
#pragma unroll
        for (int ch = 0; ch < 16; ++ch) {
          const int atom = ch >> 3, cc = ch & 7;
          *reinterpret_cast<uint4*>(prow + atom * kTileBytes + ((cc ^ (r & 7)) << 4)) =
              make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      // final: O / L
      mbar_wait(o_full(g), (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t lo[32], hi[32], lv[16];
      tmem_ld_32x32b_x32(o_addr, lo);
      tmem_ld_32x32b_x32(o_addr + 32, hi);
      tmem_ld_32x32b_x16(l_addr, lv);
      tmem_ld_wait();
#if !EVW_ATTN_ONES_MMA
      lv[0] = __float_as_uint(l_run);
#endif
      if (qrow < S) {
        const float inv = 1.0f / __uint_as_float(lv[0]);
        uint4* op = reinterpret_cast<uint4*>(out + ((long long)frame * S + qrow) * C + head * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t* src = (i < 4) ? &lo[8 * i] : &hi[8 * (i - 4)];
          __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          uint4 v;
          v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
          v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
          op[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ==========================================================================================
// v6: two threads per query row.  ncu on v5 (profiles/r01c_attn5_*.txt) shows each softmax warp's per-tile chain —
// wait S, tcgen05.ld, max, 128 exponentials, wait PV, store P, fence — is ~3300 cycles long with only 2 softmax warps
// per SM sub-partition to overlap it (issue 54 %, MUFU 48 %, tensor 36 % busy).  Here a 128-row group is served by EIGHT
// warps: warp (quarter, half) owns TMEM lanes 32 quarter .. +31 and score columns 64 half .. +63, so every sub-partition
// holds 4 softmax warps (two per group) with half the per-thread work and registers (setmaxnreg 112 instead of 224).
// The two halves of a row exchange their partial maxima through shared memory (one 256-thread named barrier per tile);
// everything else — lazy rescale, TMEM-resident O and row sums, P through swizzled shared memory — is v3/v5.
//   warp 0 TMA | warp 1 MMA group A | warp 2 MMA group B | warp 3 idle | warps 4..11 group A | warps 12..19 group B
// ==========================================================================================
constexpr int kA6Threads = 640;
constexpr int kA6OffToken = kA3OffOnes + 4096;       // 512 words: one per softmax thread (stagger token dependency)
constexpr int kA6OffMax = kA6OffToken + 2048;        // [2 parities][2 groups][2 halves][128 rows] fp32 partial maxima
constexpr int kA6Smem = kA6OffMax + 4096 + 1024;

template <int kPolyN, bool kStagger>
__global__ void __launch_bounds__(kA6Threads, 1)
spatial_attn6_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out, int S, int C, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, frame = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kBQ;
  const int n_kv = (S + kBKV - 1) / kBKV;
  const bool b_active = q0 + kBQ < S;

  const uint32_t bar = base + kA2OffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kKvStages + s); };
  auto s_full = [&](int g) { return bar + 8u * (1 + 2 * kKvStages + g); };
  auto s_empty = [&](int g) { return bar + 8u * (3 + 2 * kKvStages + g); };
  auto p_full = [&](int g) { return bar + 8u * (5 + 2 * kKvStages + g); };
  auto o_full = [&](int g) { return bar + 8u * (7 + 2 * kKvStages + g); };
  const uint32_t tmem_slot = bar + 8u * (9 + 2 * kKvStages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kA2OffBar + 8 * (9 + 2 * kKvStages));

  for (int i = threadIdx.x; i < 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base_ptr + kA3OffOnes)[i] = 0x3C003C00u;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kKvStages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), b_active ? 2 : 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_empty(g), 8);  // one arrival per softmax warp of the group
      mbar_init(p_full(g), 8);
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        tma_load_3d(&tmap, base + kA2OffQ, q_full, head * kD, q0, frame);
        tma_load_3d(&tmap, base + kA2OffQ + kTileBytes, q_full, head * kD, q0 + kBQ, frame);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kKvStages;
        mbar_wait(kv_empty(st), ((j / kKvStages) & 1) ^ 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(kv_full(st), 2 * kTileBytes);
          tma_load_3d(&tmap, base + kA2OffK + st * kTileBytes, kv_full(st), C + head * kD, j * kBKV, frame);
          tma_load_3d(&tmap, base + kA2OffV + st * kTileBytes, kv_full(st), 2 * C + head * kD, j * kBKV, frame);
        }
        __syncwarp();
      }
    } else if (warp == 1 || warp == 2) {
      const int g = warp - 1;
      if (lane == 0 && (g == 0 || b_active)) {
        const uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0, 0);
        const uint32_t idesc_o = make_idesc_f16(kBQ, kD, 0, 0, 1);
        const uint32_t idesc_l = make_idesc_f16(kBQ, 16, 0, 0, 0);
        const uint64_t dq = make_desc_k_sw128(base + kA2OffQ + g * kTileBytes);
        const uint64_t dp = make_desc_k_sw128(base + kA2OffP + g * 2 * kTileBytes);
        const uint64_t d1 = make_desc_k_sw128(base + kA3OffOnes);
        const uint64_t dk0 = make_desc_k_sw128(base + kA2OffK);
        const uint64_t dv0 = make_desc_mn_sw128(base + kA2OffV, 1024);
        const uint32_t t_s = tmem_base + g * kBKV, t_o = tmem_base + kTmemO3 + g * kD, t_l = tmem_base + kTmemL3 + g * 16;
        mbar_wait(q_full, 0);
        int s_next = 0, pv_next = 0;
        long long t0 = clock64();
        while (pv_next < n_kv) {
          bool progress = false;
          if (s_next < n_kv && mbar_test(kv_full(s_next % kKvStages), (s_next / kKvStages) & 1) &&
              (s_next == 0 || mbar_test(s_empty(g), (s_next - 1) & 1))) {
            tc_fence_after();
            const uint64_t dk = desc_add(dk0, (s_next % kKvStages) * (kTileBytes >> 4));
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) umma_f16_ss(t_s, desc_add(dq, 2 * k), desc_add(dk, 2 * k), idesc_s, k != 0);
            tc_commit(s_full(g));
            ++s_next;
            progress = true;
          }
          if (pv_next < s_next && mbar_test(p_full(g), pv_next & 1)) {
            tc_fence_after();
            const int st = pv_next % kKvStages;
            const uint64_t dv = desc_add(dv0, st * (kTileBytes >> 4));
            const uint32_t acc = pv_next != 0;
#pragma unroll
            for (int ks = 0; ks < kBKV / 16; ++ks) {
              const uint64_t a = desc_add(dp, (ks >> 2) * (kTileBytes >> 4) + 2 * (ks & 3));
              umma_f16_ss(t_o, a, desc_add(dv, ks * (2048 >> 4)), idesc_o, acc | (ks != 0));                         // O += P V
              umma_f16_ss(t_l, a, desc_add(d1, (ks >> 2) * (2048 >> 4) + 2 * (ks & 3)), idesc_l, acc | (ks != 0));   // L += P 1
            }
            tc_commit(o_full(g));
            tc_commit(kv_empty(st));
            ++pv_next;
            progress = true;
          }
          if (progress) t0 = clock64();
          else if (clock64() - t0 > 8000000000ll) __trap();
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================== softmax: 8 warps per 128-row group =====================
    const int sw = warp - 4;
    const int g = sw >> 3;
    const int half = (sw >> 2) & 1;
    const int quarter = warp & 3;  // = sw & 3: the TMEM lane quarter this warp may address
    const int r = quarter * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    if (g == 0 || b_active) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t s_addr = lane_addr + g * kBKV + half * 64;
      const uint32_t o_addr = lane_addr + kTmemO3 + g * kD + half * 32;
      const uint32_t l_addr = lane_addr + kTmemL3 + g * 16;
      float m_ref = -INFINITY;
      const uint32_t token_word = base + kA6OffToken + 4u * (threadIdx.x - 128);
      if (kStagger) st_shared_volatile(token_word, 0u);
      if (kStagger && b_active && g == 1) named_bar_arrive(1, 512);  // group A goes first
      uint8_t* prow = base_ptr + kA2OffP + g * 2 * kTileBytes + half * kTileBytes + r * 128;
      float* mxbuf = reinterpret_cast<float*>(base_ptr + kA6OffMax);

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full(g), j & 1);
        tc_fence_after();
        uint32_t s[64];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          tmem_ld_32x32b_x32(s_addr + 0, s0);
          tmem_ld_32x32b_x32(s_addr + 32, s1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(g));  // 8 arrivals: S_g(j+1) may now overwrite the TMEM tile

        const int kv_valid = S - j * kBKV - half * 64;  // valid columns of this half
        if (kv_valid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= kv_valid) s[c] = 0xff800000u;
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 64; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(s[c]));
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // exchange the half-row maxima (double-buffered by tile parity: a slot is rewritten two barriers later)
        {
          float* slot = mxbuf + (((j & 1) * 2 + g) * 2) * 128;
          slot[half * 128 + r] = mx;
          named_bar_sync(3 + g, 256);
          mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);
        }
        const float m_tile = mx * scale_log2e;
        const bool need = m_tile > m_ref + kLazyTau;  // identical in both threads of the row
        const float m_old = m_ref;
        if (need) m_ref = m_tile;
        const float neg_m = -m_ref;
        if (kPolyN > 0) {
#pragma unroll
          for (int c = kPolyN - 1; c < 64; c += (kPolyN > 0 ? kPolyN : 64))
            s[c] = __float_as_uint(exp2_poly(fmaf(__uint_as_float(s[c]), scale_log2e, neg_m)));
        }
        float neg_m_dep = neg_m;
        if (kStagger && b_active) {
          named_bar_sync(1 + g, 512);
          neg_m_dep = neg_m + __uint_as_float(ld_shared_volatile(token_word));
        }
        uint32_t w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool poly0 = kPolyN > 0 && ((2 * i) % (kPolyN > 0 ? kPolyN : 1)) == kPolyN - 1;
          const bool poly1 = kPolyN > 0 && ((2 * i + 1) % (kPolyN > 0 ? kPolyN : 1)) == kPolyN - 1;
          const float p0 = poly0 ? __uint_as_float(s[2 * i]) : fast_exp2(fmaf(__uint_as_float(s[2 * i]), scale_log2e, neg_m_dep));
          const float p1 = poly1 ? __uint_as_float(s[2 * i + 1])
                                 : fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2e, neg_m_dep));
          __half2 h = __floats2half2_rn(p0, p1);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (kStagger && b_active) {
          st_shared_volatile(token_word, (w[7] | w[15] | w[23] | w[31]) & 0x80008000u);
          if (!(g == 1 && j + 1 == n_kv)) named_bar_arrive(2 - g, 512);
        }
        // P buffer / O, L accumulators of this group are free once PV_g(j-1) has completed
        if (j > 0) {
          mbar_wait(o_full(g), (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {  // same rows, same decision in both warps of the quarter
            const float f = need ? fast_exp2(m_old - m_ref) : 1.0f;
#pragma unroll
            for (int part = 0; part < 2; ++part) {  // this warp's 32 of the 64 output columns
              uint32_t v[16];
              tmem_ld_32x32b_x16(o_addr + part * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st_32x32b_x16(o_addr + part * 16, v);
            }
            if (half == 0) {
              uint32_t v[16];
              tmem_ld_32x32b_x16(l_addr, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
              tmem_st_32x32b_x16(l_addr, v);
            }
            tmem_st_wait();
          }
        }
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
          *reinterpret_cast<uint4*>(prow + ((cc ^ (r & 7)) << 4)) = make_uint4(w[4 * cc], w[4 * cc + 1], w[4 * cc + 2], w[4 * cc + 3]);
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      // final: this warp's 32 output columns, normalised by the row sum
      mbar_wait(o_full(g), (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t ov[32], lv[16];
      tmem_ld_32x32b_x32(o_addr, ov);
      tmem_ld_32x32b_x16(l_addr, lv);
      tmem_ld_wait();
      if (qrow < S) {
        const float inv = 1.0f / __uint_as_float(lv[0]);
        uint4* op = reinterpret_cast<uint4*>(out + ((long long)frame * S + qrow) * C + head * kD + half * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t* src = &ov[8 * i];
          __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          uint4 v;
          v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
          v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
          op[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ==========================================================================================
// v7: P never touches shared memory, and does not alias S.  ncu on v5/v6 (profiles/r01c_ncu_full_attn5.txt:
// l1tex__data_pipe_tc_wavefronts_mem_shared 56 % + LSU shared stores 16 %) shows the shared-memory data pipe is what
// every earlier generation saturates: per (2 x 128 query, 128 key) tile pair the MMAs read P twice (PV and the
// row-sum MMA, 128 KB) plus Q, K, V (96 KB) and the softmax warps write P (64 KB) — ~2600 of the ~3300 cycles a tile
// pair takes.  Here the softmax warps write P (fp16 pairs) with tcgen05.st into its OWN 64 tensor-memory columns
// per group and the PV MMA takes its A operand from tensor memory, so shared memory only carries Q, K, V (96 KB read,
// 32 KB TMA-written per tile pair).  Unlike v4, S(j+1) can still be issued while the softmax of tile j runs.
// Tensor memory is then full (S 2 x 128, O 2 x 64, P 2 x 64 columns), so the softmax denominator is accumulated by
// the CUDA cores (fp32, 4 partial sums) instead of the ones-MMA.  Groups are staggered on the MUFU pipe as in v5.
// ==========================================================================================
constexpr int kA7Stages = 4;
constexpr int kA7OffQ = 0;                                   // 2 x 16 KiB
constexpr int kA7OffK = kA7OffQ + 2 * kTileBytes;            // 4 x 16 KiB
constexpr int kA7OffV = kA7OffK + kA7Stages * kTileBytes;    // 4 x 16 KiB
constexpr int kA7OffBar = kA7OffV + kA7Stages * kTileBytes;
constexpr int kA7OffToken = kA7OffBar + 256;                 // 256 words
constexpr int kA7Smem = kA7OffToken + 1024 + 1024;
constexpr int kTmemO7 = 256, kTmemP7 = 384;                  // O_g at 256 + 64 g, P_g at 384 + 64 g

template <int kPolyN, bool kStagger>
__global__ void __launch_bounds__(kA2Threads, 1)
spatial_attn7_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out, int S, int C, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, frame = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kBQ;
  const int n_kv = (S + kBKV - 1) / kBKV;
  const bool b_active = q0 + kBQ < S;

  const uint32_t bar = base + kA7OffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kA7Stages + s); };
  auto s_full = [&](int g) { return bar + 8u * (1 + 2 * kA7Stages + g); };
  auto s_empty = [&](int g) { return bar + 8u * (3 + 2 * kA7Stages + g); };
  auto p_full = [&](int g) { return bar + 8u * (5 + 2 * kA7Stages + g); };
  auto o_full = [&](int g) { return bar + 8u * (7 + 2 * kA7Stages + g); };
  const uint32_t tmem_slot = bar + 8u * (9 + 2 * kA7Stages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kA7OffBar + 8 * (9 + 2 * kA7Stages));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kA7Stages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), b_active ? 2 : 1);  // one tcgen05.commit per active group
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_empty(g), 4);
      mbar_init(p_full(g), 4);
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        tma_load_3d(&tmap, base + kA7OffQ, q_full, head * kD, q0, frame);
        tma_load_3d(&tmap, base + kA7OffQ + kTileBytes, q_full, head * kD, q0 + kBQ, frame);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kA7Stages;
        mbar_wait(kv_empty(st), ((j / kA7Stages) & 1) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(kv_full(st), 2 * kTileBytes);
          tma_load_3d(&tmap, base + kA7OffK + st * kTileBytes, kv_full(st), C + head * kD, j * kBKV, frame);
          tma_load_3d(&tmap, base + kA7OffV + st * kTileBytes, kv_full(st), 2 * C + head * kD, j * kBKV, frame);
        }
        __syncwarp();
      }
    } else if (warp == 1 || warp == 2) {
      // ===================== MMA issuers: warp 1 drives group A, warp 2 drives group B (event driven) =====================
      const int g = warp - 1;
      if (elect_one_sync() && (g == 0 || b_active)) {
        const uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0, 0);
        const uint32_t idesc_o = make_idesc_f16(kBQ, kD, 0, 0, 1);
        const uint64_t dq = make_desc_k_sw128(base + kA7OffQ + g * kTileBytes);
        const uint64_t dk0 = make_desc_k_sw128(base + kA7OffK);
        const uint64_t dv0 = make_desc_mn_sw128(base + kA7OffV, 1024);
        const uint32_t t_s = tmem_base + g * kBKV, t_o = tmem_base + kTmemO7 + g * kD, t_p = tmem_base + kTmemP7 + g * 64;
        mbar_wait(q_full, 0);
        int s_next = 0, pv_next = 0;
        long long t0 = clock64();
        while (pv_next < n_kv) {
          bool progress = false;
          if (s_next < n_kv && mbar_test(kv_full(s_next % kA7Stages), (s_next / kA7Stages) & 1) &&
              (s_next == 0 || mbar_test(s_empty(g), (s_next - 1) & 1))) {
            tc_fence_after();
            const uint64_t dk = desc_add(dk0, (s_next % kA7Stages) * (kTileBytes >> 4));
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) umma_f16_ss(t_s, desc_add(dq, 2 * k), desc_add(dk, 2 * k), idesc_s, k != 0);
            tc_commit(s_full(g));
            ++s_next;
            progress = true;
          }
          if (pv_next < s_next && mbar_test(p_full(g), pv_next & 1)) {
            tc_fence_after();
            const int st = pv_next % kA7Stages;
            const uint64_t dv = desc_add(dv0, st * (kTileBytes >> 4));
            const uint32_t acc = pv_next != 0;
#pragma unroll
            for (int ks = 0; ks < kBKV / 16; ++ks)  // O += P V, A operand (16 keys = 8 packed columns) from tensor memory
              umma_f16_ts(t_o, t_p + ks * 8, desc_add(dv, ks * (2048 >> 4)), idesc_o, acc | (ks != 0));
            tc_commit(o_full(g));
            tc_commit(kv_empty(st));
            ++pv_next;
            progress = true;
          }
          if (progress) t0 = clock64();
          else if (clock64() - t0 > 8000000000ll) __trap();
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ===================== softmax groups =====================
    const int g = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    if (g == 0 || b_active) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t s_addr = lane_addr + g * kBKV, o_addr = lane_addr + kTmemO7 + g * kD, p_addr = lane_addr + kTmemP7 + g * 64;
      float m_ref = -INFINITY;
      float l_run = 0.f;
      const uint32_t token_word = base + kA7OffToken + 4u * (threadIdx.x - 128);
      if (kStagger) st_shared_volatile(token_word, 0u);
      if (kStagger && b_active && g == 1) named_bar_arrive(1, 256);  // group A goes first

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full(g), j & 1);
        tc_fence_after();
        uint32_t s[128];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
          uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
          tmem_ld_32x32b_x32(s_addr + 0, s0);
          tmem_ld_32x32b_x32(s_addr + 32, s1);
          tmem_ld_32x32b_x32(s_addr + 64, s2);
          tmem_ld_32x32b_x32(s_addr + 96, s3);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(g));  // S_g(j+1) may now overwrite the TMEM tile

        const int kv_valid = S - j * kBKV;
        if (kv_valid < kBKV) {
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c >= kv_valid) s[c] = 0xff800000u;
        }
        float mx;
        {
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int c = 0; c < 128; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(s[c]));
          mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        }
        const float m_tile = mx * scale_log2e;
        const bool need = m_tile > m_ref + kLazyTau;  // first tile: m_ref = -inf -> true
        const float m_old = m_ref;
        if (need) m_ref = m_tile;
        const float neg_m = -m_ref;
        if (kPolyN > 0) {
#pragma unroll
          for (int c = kPolyN - 1; c < 128; c += (kPolyN > 0 ? kPolyN : 128))
            s[c] = __float_as_uint(exp2_poly(fmaf(__uint_as_float(s[c]), scale_log2e, neg_m)));
        }
        float neg_m_dep = neg_m;
        if (kStagger && b_active) {
          named_bar_sync(1 + g, 256);
          neg_m_dep = neg_m + __uint_as_float(ld_shared_volatile(token_word));
        }
        uint32_t w[64];
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const bool poly0 = kPolyN > 0 && ((2 * i) % (kPolyN > 0 ? kPolyN : 1)) == kPolyN - 1;
          const bool poly1 = kPolyN > 0 && ((2 * i + 1) % (kPolyN > 0 ? kPolyN : 1)) == kPolyN - 1;
          const float p0 = poly0 ? __uint_as_float(s[2 * i]) : fast_exp2(fmaf(__uint_as_float(s[2 * i]), scale_log2e, neg_m_dep));
          const float p1 = poly1 ? __uint_as_float(s[2 * i + 1])
                                 : fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2e, neg_m_dep));
          sum4[(2 * i) & 3] += p0;
          sum4[(2 * i + 1) & 3] += p1;
          __half2 h = __floats2half2_rn(p0, p1);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (kStagger && b_active) {
          st_shared_volatile(token_word, (w[15] | w[31] | w[47] | w[63]) & 0x80008000u);
          if (!(g == 1 && j + 1 == n_kv)) named_bar_arrive(2 - g, 256);
        }
        const float f_resc = need ? fast_exp2(m_old - m_ref) : 1.0f;  // exp2(-inf) = 0 on the first tile
        l_run = fmaf(l_run, f_resc, (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
        // the P columns and the O accumulator of this group are free once PV_g(j-1) has completed
        if (j > 0) {
          mbar_wait(o_full(g), (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {
#pragma unroll
            for (int part = 0; part < 4; ++part) {
              uint32_t v[16];
              tmem_ld_32x32b_x16(o_addr + part * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f_resc);
              tmem_st_32x32b_x16(o_addr + part * 16, v);
            }
          }
        }
        // P: two probabilities per 32-bit column (even key in the low half), 64 columns
#pragma unroll
        for (int part = 0; part < 4; ++part) {
          uint32_t(&wp)[16] = *reinterpret_cast<uint32_t(*)[16]>(&w[16 * part]);
          tmem_st_32x32b_x16(p_addr + part * 16, wp);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      // final: O / l
      mbar_wait(o_full(g), (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t lo[32], hi[32];
      tmem_ld_32x32b_x32(o_addr, lo);
      tmem_ld_32x32b_x32(o_addr + 32, hi);
      tmem_ld_wait();
      if (qrow < S) {
        const float inv = 1.0f / l_run;
        uint4* op = reinterpret_cast<uint4*>(out + ((long long)frame * S + qrow) * C + head * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t* src = (i < 4) ? &lo[8 * i] : &hi[8 * (i - 4)];
          __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          uint4 v;
          v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
          v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
          op[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ==========================================================================================
// v8: v7 (P in its own tensor-memory columns, CUDA-core row sums) with v6's two threads per query row: 16 softmax
// warps per CTA, four per SM sub-partition, to overlap the per-tile chain that v7 still leaves exposed.
// ==========================================================================================
constexpr int kA8OffToken = kA7OffBar + 256;          // 512 words
constexpr int kA8OffMax = kA8OffToken + 2048;         // [2 parities][2 groups][2 halves][128 rows] fp32
constexpr int kA8Smem = kA8OffMax + 4096 + 1024;

template <int kPolyN, bool kStagger>
__global__ void __launch_bounds__(kA6Threads, 1)
spatial_attn8_kernel(const __grid_constant__ CUtensorMap tmap, __half* __restrict__ out, int S, int C, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, frame = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kBQ;
  const int n_kv = (S + kBKV - 1) / kBKV;
  const bool b_active = q0 + kBQ < S;

  const uint32_t bar = base + kA7OffBar;
  const uint32_t q_full = bar;
  auto kv_full = [&](int s) { return bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar + 8u * (1 + kA7Stages + s); };
  auto s_full = [&](int g) { return bar + 8u * (1 + 2 * kA7Stages + g); };
  auto s_empty = [&](int g) { return bar + 8u * (3 + 2 * kA7Stages + g); };
  auto p_full = [&](int g) { return bar + 8u * (5 + 2 * kA7Stages + g); };
  auto o_full = [&](int g) { return bar + 8u * (7 + 2 * kA7Stages + g); };
  const uint32_t tmem_slot = bar + 8u * (9 + 2 * kA7Stages);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kA7OffBar + 8 * (9 + 2 * kA7Stages));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(q_full, 1);
    for (int s = 0; s < kA7Stages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), b_active ? 2 : 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(s_empty(g), 8);  // one arrival per softmax warp of the group
      mbar_init(p_full(g), 8);
      mbar_init(o_full(g), 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        tma_load_3d(&tmap, base + kA7OffQ, q_full, head * kD, q0, frame);
        tma_load_3d(&tmap, base + kA7OffQ + kTileBytes, q_full, head * kD, q0 + kBQ, frame);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kA7Stages;
        mbar_wait(kv_empty(st), ((j / kA7Stages) & 1) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(kv_full(st), 2 * kTileBytes);
          tma_load_3d(&tmap, base + kA7OffK + st * kTileBytes, kv_full(st), C + head * kD, j * kBKV, frame);
          tma_load_3d(&tmap, base + kA7OffV + st * kTileBytes, kv_full(st), 2 * C + head * kD, j * kBKV, frame);
        }
        __syncwarp();
      }
    } else if (warp == 1 || warp == 2) {
      const int g = warp - 1;
      if (elect_one_sync() && (g == 0 || b_active)) {
        const uint32_t idesc_s = make_idesc_f16(kBQ, kBKV, 0, 0, 0);
        const uint32_t idesc_o = make_idesc_f16(kBQ, kD, 0, 0, 1);
        const uint64_t dq = make_desc_k_sw128(base + kA7OffQ + g * kTileBytes);
        const uint64_t dk0 = make_desc_k_sw128(base + kA7OffK);
        const uint64_t dv0 = make_desc_mn_sw128(base + kA7OffV, 1024);
        const uint32_t t_s = tmem_base + g * kBKV, t_o = tmem_base + kTmemO7 + g * kD, t_p = tmem_base + kTmemP7 + g * 64;
        mbar_wait(q_full, 0);
        int s_next = 0, pv_next = 0;
        long long t0 = clock64();
        while (pv_next < n_kv) {
          bool progress = false;
          if (s_next < n_kv && mbar_test(kv_full(s_next % kA7Stages), (s_next / kA7Stages) & 1) &&
              (s_next == 0 || mbar_test(s_empty(g), (s_next - 1) & 1))) {
            tc_fence_after();
            const uint64_t dk = desc_add(dk0, (s_next % kA7Stages) * (kTileBytes >> 4));
#pragma unroll
            for (int k = 0; k < kD / 16; ++k) umma_f16_ss(t_s, desc_add(dq, 2 * k), desc_add(dk, 2 * k), idesc_s, k != 0);
            tc_commit(s_full(g));
            ++s_next;
            progress = true;
          }
          if (pv_next < s_next && mbar_test(p_full(g), pv_next & 1)) {
            tc_fence_after();
            const int st = pv_next % kA7Stages;
            const uint64_t dv = desc_add(dv0, st * (kTileBytes >> 4));
            const uint32_t acc = pv_next != 0;
#pragma unroll
            for (int ks = 0; ks < kBKV / 16; ++ks)  // O += P V, A operand from tensor memory
              umma_f16_ts(t_o, t_p + ks * 8, desc_add(dv, ks * (2048 >> 4)), idesc_o, acc | (ks != 0));
            tc_commit(o_full(g));
            tc_commit(kv_empty(st));
            ++pv_next;
            progress = true;
          }
          if (progress) t0 = clock64();
          else if (clock64() - t0 > 8000000000ll) __trap();
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================== softmax: 8 warps per 128-row group =====================
    const int sw = warp - 4;
    const int g = sw >> 3;
    const int half = (sw >> 2) & 1;
    const int quarter = warp & 3;  // = sw & 3: the TMEM lane quarter this warp may address
    const int r = quarter * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    if (g == 0 || b_active) {
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const uint32_t s_addr = lane_addr + g * kBKV + half * 64;
      const uint32_t o_addr = lane_addr + kTmemO7 + g * kD + half * 32;
      const uint32_t p_addr = lane_addr + kTmemP7 + g * 64 + half * 32;
      float m_ref = -INFINITY;
      float l_run = 0.f;  // this thread's half of the row sum
      const uint32_t token_word = base + kA8OffToken + 4u * (threadIdx.x - 128);
      if (kStagger) st_shared_volatile(token_word, 0u);
      if (kStagger && b_active && g == 1) named_bar_arrive(1, 512);  // group A goes first
      float* mxbuf = reinterpret_cast<float*>(base_ptr + kA8OffMax);

      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full(g), j & 1);
        tc_fence_after();
        uint32_t s[64];
        {
          uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
          uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
          tmem_ld_32x32b_x32(s_addr + 0, s0);
          tmem_ld_32x32b_x32(s_addr + 32, s1);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_empty(g));  // 8 arrivals: S_g(j+1) may now overwrite the TMEM tile

        const int kv_valid = S - j * kBKV - half * 64;  // valid columns of this half
        if (kv_valid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= kv_valid) s[c] = 0xff800000u;
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 64; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(s[c]));
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // exchange the half-row maxima (double-buffered by tile parity: a slot is rewritten two barriers later)
        {
          float* slot = mxbuf + (((j & 1) * 2 + g) * 2) * 128;
          slot[half * 128 + r] = mx;
          named_bar_sync(3 + g, 256);
          mx = fmaxf(mx, slot[(half ^ 1) * 128 + r]);
        }
        const float m_tile = mx * scale_log2e;
        const bool need = m_tile > m_ref + kLazyTau;  // identical in both threads of the row
        const float m_old = m_ref;
        if (need) m_ref = m_tile;
        const float neg_m = -m_ref;
        if (kPolyN > 0) {
#pragma unroll
          for (int c = kPolyN - 1; c < 64; c += (kPolyN > 0 ? kPolyN : 64))
            s[c] = __float_as_uint(exp2_poly(fmaf(__uint_as_float(s[c]), scale_log2e, neg_m)));
        }
        float neg_m_dep = neg_m;
        if (kStagger && b_active) {
          named_bar_sync(1 + g, 512);
          neg_m_dep = neg_m + __uint_as_float(ld_shared_volatile(token_word));
        }
        uint32_t w[32];
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool poly0 = kPolyN > 0 && ((2 * i) % (kPolyN > 0 ? kPolyN : 1)) == kPolyN - 1;
          const bool poly1 = kPolyN > 0 && ((2 * i + 1) % (kPolyN > 0 ? kPolyN : 1)) == kPolyN - 1;
          const float p0 = poly0 ? __uint_as_float(s[2 * i]) : fast_exp2(fmaf(__uint_as_float(s[2 * i]), scale_log2e, neg_m_dep));
          const float p1 = poly1 ? __uint_as_float(s[2 * i + 1])
                                 : fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), scale_log2e, neg_m_dep));
          sum4[(2 * i) & 3] += p0;
          sum4[(2 * i + 1) & 3] += p1;
          __half2 h = __floats2half2_rn(p0, p1);
          w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (kStagger && b_active) {
          st_shared_volatile(token_word, (w[7] | w[15] | w[23] | w[31]) & 0x80008000u);
          if (!(g == 1 && j + 1 == n_kv)) named_bar_arrive(2 - g, 512);
        }
        const float f_resc = need ? fast_exp2(m_old - m_ref) : 1.0f;  // exp2(-inf) = 0 on the first tile
        l_run = fmaf(l_run, f_resc, (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
        // the P columns and the O accumulator of this group are free once PV_g(j-1) has completed
        if (j > 0) {
          mbar_wait(o_full(g), (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {  // same rows, same decision in both warps of the quarter
#pragma unroll
            for (int part = 0; part < 2; ++part) {  // this warp's 32 of the 64 output columns
              uint32_t v[16];
              tmem_ld_32x32b_x16(o_addr + part * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f_resc);
              tmem_st_32x32b_x16(o_addr + part * 16, v);
            }
          }
        }
        // P: this warp's 64 keys = 32 packed columns
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          uint32_t(&wp)[16] = *reinterpret_cast<uint32_t(*)[16]>(&w[16 * part]);
          tmem_st_32x32b_x16(p_addr + part * 16, wp);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      // final: the two halves add their partial row sums, then each normalises its 32 output columns
      {
        float* slot = mxbuf + (((n_kv & 1) * 2 + g) * 2) * 128;  // the parity the last tile did not use
        slot[half * 128 + r] = l_run;
        named_bar_sync(3 + g, 256);
        l_run += slot[(half ^ 1) * 128 + r];
      }
      mbar_wait(o_full(g), (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t ov[32];
      tmem_ld_32x32b_x32(o_addr, ov);
      tmem_ld_wait();
      if (qrow < S) {
        const float inv = 1.0f / l_run;
        uint4* op = reinterpret_cast<uint4*>(out + ((long long)frame * S + qrow) * C + head * kD + half * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t* src = &ov[8 * i];
          __half2 h0 = __floats2half2_rn(__uint_as_float(src[0]) * inv, __uint_as_float(src[1]) * inv);
          __half2 h1 = __floats2half2_rn(__uint_as_float(src[2]) * inv, __uint_as_float(src[3]) * inv);
          __half2 h2 = __floats2half2_rn(__uint_as_float(src[4]) * inv, __uint_as_float(src[5]) * inv);
          __half2 h3 = __floats2half2_rn(__uint_as_float(src[6]) * inv, __uint_as_float(src[7]) * inv);
          uint4 v;
          v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
          v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
          op[i] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


}  // namespace

// default kernel: v8 (P in tensor memory, two threads per row), all exponentials on MUFU, no stagger
// (tools/attn_bench.py at 28 frames x 9216 tokens x 5 heads: v3 4.99 ms, v5 4.66, v7 3.87, v8 3.69)
constexpr int kDefaultAttnVariant = 14;
static int g_attn_variant_override = -2;
void set_attention_variant(int v) { g_attn_variant_override = v; }

int spatial_attention(const __half* qkv, __half* out, int F, int S, int heads, cudaStream_t st) {
  EVW_CHECK_ARG(qkv && out && F > 0 && S > 0 && heads > 0, "spatial_attention: bad arguments");
  const int C = heads * kD;
  static bool attr_set = false;
  static bool use_v1 = false, use_v2 = false, use_v3 = false, use_v4 = false;
  static int v5_variant = kDefaultAttnVariant;
  typedef void (*attn5_fn)(const CUtensorMap, __half*, int, int, float);
  struct Variant { attn5_fn fn; int threads, smem; };
  // v5 / v6 variants: {polynomial share, stagger}.  EVW_ATTN_V5=<index> or evw_set_attention_variant selects one.
  static const Variant v5_table[] = {
      {spatial_attn5_kernel<4, true>, kA2Threads, kA5Smem},  {spatial_attn5_kernel<8, true>, kA2Threads, kA5Smem},
      {spatial_attn5_kernel<0, true>, kA2Threads, kA5Smem},  {spatial_attn5_kernel<2, true>, kA2Threads, kA5Smem},
      {spatial_attn5_kernel<4, false>, kA2Threads, kA5Smem}, {spatial_attn6_kernel<4, false>, kA6Threads, kA6Smem},
      {spatial_attn6_kernel<4, true>, kA6Threads, kA6Smem},  {spatial_attn6_kernel<0, false>, kA6Threads, kA6Smem},
      {spatial_attn6_kernel<8, false>, kA6Threads, kA6Smem}, {spatial_attn7_kernel<4, true>, kA2Threads, kA7Smem},
      {spatial_attn7_kernel<4, false>, kA2Threads, kA7Smem}, {spatial_attn7_kernel<0, true>, kA2Threads, kA7Smem},
      {spatial_attn7_kernel<8, true>, kA2Threads, kA7Smem},  {spatial_attn7_kernel<0, false>, kA2Threads, kA7Smem},
      {spatial_attn8_kernel<0, false>, kA6Threads, kA8Smem}, {spatial_attn8_kernel<8, false>, kA6Threads, kA8Smem},
      {spatial_attn8_kernel<0, true>, kA6Threads, kA8Smem}};
  if (!attr_set) {
    EVW_CUDA(cudaFuncSetAttribute(spatial_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    EVW_CUDA(cudaFuncSetAttribute(spatial_attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kA2Smem));
    EVW_CUDA(cudaFuncSetAttribute(spatial_attn3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kA3Smem));
    EVW_CUDA(cudaFuncSetAttribute(spatial_attn4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kA4Smem));
    for (const Variant& v : v5_table) EVW_CUDA(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem));
    use_v1 = getenv("EVW_ATTN_V1") != nullptr;
    use_v2 = getenv("EVW_ATTN_V2") != nullptr;
    use_v3 = getenv("EVW_ATTN_V3") != nullptr;
    use_v4 = getenv("EVW_ATTN_V4") != nullptr;
    if (const char* e = getenv("EVW_ATTN_V5")) v5_variant = atoi(e);
    attr_set = true;
  }
  if (g_attn_variant_override >= -1) {  // evw_set_attention_variant (micro-benchmarks): -1 = v3, >= 0 = v5 table index
    use_v3 = g_attn_variant_override == -1;
    if (g_attn_variant_override >= 0) v5_variant = g_attn_variant_override;
    use_v1 = use_v2 = use_v4 = false;
  }
  if (v5_variant < 0 || v5_variant >= (int)(sizeof(v5_table) / sizeof(v5_table[0]))) v5_variant = kDefaultAttnVariant;
  alignas(64) CUtensorMap tmap;
  uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)S, (uint64_t)F};
  uint64_t str[2] = {(uint64_t)3 * C * 2, (uint64_t)3 * C * 2 * S};
  uint32_t box[3] = {(uint32_t)kD, (uint32_t)kBQ, 1};
  int rc = encode_tmap_f16(&tmap, qkv, 3, dims, str, box);
  if (rc) return rc;
  if (use_v1) {
    dim3 grid((S + kBQ - 1) / kBQ, heads, F);
    spatial_attn_kernel<<<grid, kAttnThreads, kAttnSmem, st>>>(tmap, out, S, C, 0.125f * 1.4426950408889634f);
  } else if (use_v2) {
    dim3 grid((S + 2 * kBQ - 1) / (2 * kBQ), heads, F);
    spatial_attn2_kernel<<<grid, kA2Threads, kA2Smem, st>>>(tmap, out, S, C, 0.125f * 1.4426950408889634f);
  } else if (use_v4) {  // P in tensor memory: correct, but the serialised S/PV chain makes it slower than v3 (5.9 vs 5.0 ms)
    dim3 grid((S + 2 * kBQ - 1) / (2 * kBQ), heads, F);
    spatial_attn4_kernel<<<grid, kA2Threads, kA4Smem, st>>>(tmap, out, S, C, 0.125f * 1.4426950408889634f);
  } else if (use_v3) {
    dim3 grid((S + 2 * kBQ - 1) / (2 * kBQ), heads, F);
    spatial_attn3_kernel<<<grid, kA2Threads, kA3Smem, st>>>(tmap, out, S, C, 0.125f * 1.4426950408889634f);
  } else {
    dim3 grid((S + 2 * kBQ - 1) / (2 * kBQ), heads, F);
    const Variant& v = v5_table[v5_variant];
    v.fn<<<grid, v.threads, v.smem, st>>>(tmap, out, S, C, 0.125f * 1.4426950408889634f);
  }
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

}  // namespace evw

// C ABI for tests / micro-benchmarks
extern "C" int evw_spatial_attention_f16(const void* qkv, void* out, int F, int S, int heads, void* stream) {
  return evw::spatial_attention((const __half*)qkv, (__half*)out, F, S, heads, (cudaStream_t)stream);
}
extern "C" void evw_set_attention_variant(int variant) { evw::set_attention_variant(variant); }
extern "C" int evw_temporal_attention_f16(const void* qkv, void* out, int B, int T, int64_t S, int heads, void* stream) {
  return evw::temporal_attention((const __half*)qkv, (__half*)out, B, T, S, heads, (cudaStream_t)stream);
}
extern "C" int evw_group_norm_f16(const void* src0, int src0_fp16, int C0, const float* src1, int C1, int64_t insts,
                                  int64_t rows_per_inst, float eps, const float* gamma, const float* beta, int do_silu,
                                  void* stats_ws, void* out, void* raw_out, void* out_lo, void* stream) {
  return evw::group_norm(src0, src0_fp16, C0, src1, C1, insts, rows_per_inst, eps, gamma, beta, do_silu, (double*)stats_ws,
                         (__half*)out, (__half*)raw_out, (__half*)out_lo, (cudaStream_t)stream);
}
extern "C" int evw_layer_norm_f16(const float* x, const float* rowvec, int64_t rv_div, int64_t rv_mod, int64_t rows, int C,
                                  float eps, const float* gamma, const float* beta, void* out, void* stream) {
  return evw::layer_norm(x, rowvec, rv_div > 0 ? rv_div : 1, rv_mod > 0 ? rv_mod : 1, rows, C, eps, gamma, beta, (__half*)out,
                         (cudaStream_t)stream);
}
