// Spatial self-attention of the UNet transformer blocks (K6 in SURVEY §2.3): flash-attention forward,
// head dim 64, fp16 operands, on tcgen05 tensor cores with TMEM-resident score, probability and output tiles.
//   diffusers attention_processor.py AttnProcessor2_0 as used by BasicTransformerBlock.attn1:
//   softmax(Q K^T / sqrt(64)) V per (frame, head) over S = h*w tokens.
//
// Two kernel generations are kept (evw_set_attention_variant / EVW_ATTN_VARIANT pick one; both are checked against the
// same fp32 reference in tests/test_gpu_unet_ops.py):
//   spatial_attn8_kernel  the DEFAULT: P in its own tensor-memory columns (the PV MMA takes its A operand from TMEM, so
//                         shared memory only carries Q, K, V), CUDA-core row sums, lazy rescale, two threads per query row
//   spatial_attn7_kernel  its predecessor with one thread per row (A/B reference)
// Round 1 went through six earlier generations (P through shared memory, ones-MMA row sums, P aliased onto S, staggered
// MUFU phases, ...); their measurements are in profiles/r01*_attn_bench*.log and the code is in the history (commit
// 36a9904).  At 28 frames x 9216 tokens x 5 heads: v3 4.99 ms, v5 4.66, v6 4.62, v7 3.87, v8 3.69 ms.  What moved the
// needle was taking P out of shared memory; v8 is bound by the MUFU pipe (exp2: 77 % busy).
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
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KiB: one Q, K_j or V_j tile
constexpr int kA2Threads = 384;           // v7: warpgroup 0 = TMA + MMA (+2 idle warps), warpgroups 1, 2 = softmax groups A, B
constexpr int kA6Threads = 640;           // v8: 4 control warps + 16 softmax warps
constexpr float kLazyTau = 8.0f;          // a row's reference exponent moves only when the tile maximum exceeds it by 2^8

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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

__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t units16) {  // advance the 14-bit start address field
  return d + units16;  // never carries out of the address field for in-range tiles
}

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

// kPairBar: the half-row maxima / sums are exchanged under a barrier of the TWO warps that share the rows (one named barrier
// per group and TMEM lane quarter, 64 threads) instead of all eight warps of the group (256 threads): the quarters of a
// group no longer wait for each other's slowest warp at every key tile.
template <int kPolyN, bool kStagger, bool kPairBar = false>
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
          if constexpr (kPairBar) named_bar_sync(3 + g * 4 + quarter, 64);
          else named_bar_sync(3 + g, 256);
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
        if constexpr (kPairBar) named_bar_sync(3 + g * 4 + quarter, 64);
          else named_bar_sync(3 + g, 256);
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
// variants: 0 = v8, 1 = v8 with every 8th exponential on the FMA pipe, 2 = v8 with the two groups staggered on the MUFU
// pipe, 3 = v7 (one thread per row), 4 = v7 with every 4th exponential on the FMA pipe, 5 = v8 with pair barriers (kPairBar)
constexpr int kDefaultAttnVariant = 0;
static int g_attn_variant_override = -2;
void set_attention_variant(int v) { g_attn_variant_override = v; }

int spatial_attention(const __half* qkv, __half* out, int F, int S, int heads, cudaStream_t st) {
  EVW_CHECK_ARG(qkv && out && F > 0 && S > 0 && heads > 0, "spatial_attention: bad arguments");
  const int C = heads * kD;
  static bool attr_set = false;
  static int env_variant = kDefaultAttnVariant;
  typedef void (*attn_fn)(const CUtensorMap, __half*, int, int, float);
  struct Variant { attn_fn fn; int threads, smem; };
  static const Variant table[] = {
      {spatial_attn8_kernel<0, false>, kA6Threads, kA8Smem}, {spatial_attn8_kernel<8, false>, kA6Threads, kA8Smem},
      {spatial_attn8_kernel<0, true>, kA6Threads, kA8Smem},  {spatial_attn7_kernel<0, false>, kA2Threads, kA7Smem},
      {spatial_attn7_kernel<4, false>, kA2Threads, kA7Smem}, {spatial_attn8_kernel<0, false, true>, kA6Threads, kA8Smem}};
  constexpr int kVariants = (int)(sizeof(table) / sizeof(table[0]));
  if (!attr_set) {
    for (const Variant& v : table) EVW_CUDA(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem));
    if (const char* e = getenv("EVW_ATTN_VARIANT")) env_variant = atoi(e);
    attr_set = true;
  }
  int variant = g_attn_variant_override >= 0 ? g_attn_variant_override : env_variant;  // evw_set_attention_variant wins
  if (variant < 0 || variant >= kVariants) variant = kDefaultAttnVariant;
  alignas(64) CUtensorMap tmap;
  uint64_t dims[3] = {(uint64_t)3 * C, (uint64_t)S, (uint64_t)F};
  uint64_t str[2] = {(uint64_t)3 * C * 2, (uint64_t)3 * C * 2 * S};
  uint32_t box[3] = {(uint32_t)kD, (uint32_t)kBQ, 1};
  int rc = encode_tmap_f16(&tmap, qkv, 3, dims, str, box);
  if (rc) return rc;
  dim3 grid((S + 2 * kBQ - 1) / (2 * kBQ), heads, F);
  const Variant& v = table[variant];
  v.fn<<<grid, v.threads, v.smem, st>>>(tmap, out, S, C, 0.125f * 1.4426950408889634f);
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
extern "C" int evw_layer_norm_f32(const float* x, int64_t rows, int C, float eps, const float* gamma, const float* beta, float* out,
                                  void* stream) {
  EVW_CHECK_ARG(x && out && gamma && beta, "evw_layer_norm_f32: null pointer");
  return evw::layer_norm(x, nullptr, 1, 1, rows, C, eps, gamma, beta, nullptr, (cudaStream_t)stream, out);
}

extern "C" int evw_layer_norm_f16(const float* x, const float* rowvec, int64_t rv_div, int64_t rv_mod, int64_t rows, int C,
                                  float eps, const float* gamma, const float* beta, void* out, void* stream) {
  return evw::layer_norm(x, rowvec, rv_div > 0 ? rv_div : 1, rv_mod > 0 ? rv_mod : 1, rows, C, eps, gamma, beta, (__half*)out,
                         (cudaStream_t)stream);
}
