// tcgen05 implicit-GEMM for the UNet denoise step (hot path 1): one persistent, warp-specialised
// kernel serves every contraction of the network —
//   nn.Linear                       (1 tap)                   diffusers attention.py / embeddings.py
//   Conv2d 3x3 pad 1 (+1x1 shortcut) (9 taps [+1 from a second tensor])  diffusers resnet.py ResnetBlock2D
//   Conv3d (3,1,1) pad (1,0,0)       (3 taps along T)          diffusers resnet.py TemporalResnetBlock
// as D[m,n] = sum_{tap,k} A[m shifted by tap, k] * W[n, tap*K + k] with fp16 operands and fp32
// accumulation in tensor memory.
//
//   warp 0      TMA producer: 5-D tiled loads of the activation tile (zero fill outside the image =
//               the convolution padding) + 2-D loads of the weight tile, 128B swizzle, mbarrier ring;
//               issued under elect.sync
//   warp 1      MMA issuer: tcgen05.mma kind::f16, N=BLOCK_N (<=256), K=16 per instr.; cta_group::2 M=256 issued by the
//               leader CTA of each pair (cta_group::1 M=128 in the single-CTA mode), under elect.sync (no uniform-operand
//               waterfall: 65 SASS instructions per 64-wide k-block)
//   warp 2      TMEM allocator (512 columns = two accumulator stages)
//   warps 4..11 epilogue (8 warps; a 16-warp GEGLU-only instantiation exists, off by default):
//               tcgen05.ld -> bias / broadcast row vector / GEGLU / scaled residuals -> global
// The accumulator is double-buffered so the epilogue of tile i overlaps the main loop of tile i+1.
// Default launch mode (k2Cta): CTA PAIRS (clusters of two, one TPC) run tcgen05.mma.cta_group::2 with M = 256 — the two
// CTAs work on m-adjacent 128-row tiles of the SAME n-tile, each loads its own activation tile and HALF of the weight
// tile, and the leader CTA's single MMA stream drives both tensor cores.  Per CTA and k-block that is 16 KB + BLOCK_N*64 B
// of operand traffic instead of 16 KB + BLOCK_N*128 B: the shared-memory data pipe (TMA writes + tensor-core operand
// reads), which capped the 160-wide tiles of every N = 320 / 640 layer at ~1.05 PFLOP/s, drops from ~1.8x to ~1.25x of
// its per-MMA budget.  EVW_GEMM_CLUSTER=0 / evw_set_gemm_cluster(0) selects the independent-CTA (cta_group::1) launch.
#include "common.h"
#include "tc_common.cuh"
#include "tc_gemm.h"
#include <cstdlib>

namespace evw {

using namespace tc;

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // fp16 elements = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kAccStride = 256;  // TMEM columns between the two accumulator stages
constexpr int kCtrlThreads = 128;    // 4 control warps (TMA, MMA, TMEM allocator, spare)
// epilogue warps: 8; a 16-warp instantiation exists for the GEGLU layers (exact erf GELU, ~26 instructions per output:
// issue slots 59 % busy, tensor pipe 45 % at K = 320) but measured slower and is off by default
constexpr int kEpiWarpsDefault = 8, kEpiWarpsGeglu = 16;
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB

// Division by a run-time constant d (1 <= d, dividend < 2^31) as umulhi + add + shift (Granlund-Montgomery round-up):
// the tile-index decomposition runs per thread per tile and cost ~100 instructions per warp and tile with IDIV sequences
// (13 % of all instructions issued by the K = 320 GEMMs, profiles/r02c_ncu_gemm_epilogue.txt).
struct FastDiv {
  uint32_t d = 1, mul = 1, shr = 0;
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d ? d : 1;
  uint32_t shr = 0;
  while ((1ull << shr) < f.d) ++shr;
  f.shr = shr;
  f.mul = (uint32_t)((((1ull << 32) * ((1ull << shr) - f.d)) / f.d) + 1);
  return f;
}
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv& f) { return (__umulhi(n, f.mul) + n) >> f.shr; }
__device__ __forceinline__ void fd_divmod(uint32_t n, const FastDiv& f, uint32_t& q, uint32_t& rem) {
  q = fd_div(n, f);
  rem = n - q * f.d;
}

struct KernelParams {
  GemmEpilogue ep;
  FastDiv fd_n_tiles, fd_tiles_x, fd_tiles_y, fd_T, fd_rv_div, fd_rv_mod, fd_gn_cg, fd_gn_rpi;
  long long gn_rows_per_inst;
  int bx_shift;
  int tiles_x, tiles_y, T, B, X, Y, bx, by;
  int n_tiles, block_n, N;
  int num_taps;
  int chunks[2];
  int8_t tap_dx[kMaxTaps], tap_dy[kMaxTaps], tap_dt[kMaxTaps], tap_src[kMaxTaps];
  int num_stages;
  int res_slots;  // > 0 (a power of two): res1 comes through a ring of res_slots 16 KiB boxes (128 rows x 32 fp32 columns)
  int res_shift;  // log2(res_slots)
  int total_tiles;
  int total_pairs;  // cluster mode: pairs of m-adjacent tiles sharing one weight tile (B multicast)
};

// exact-GELU 0.5 x (1 + erf(x / sqrt 2)) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2), branch free, with erfc from
// Abramowitz-Stegun 7.1.26: erfc(z) = (a1 t + .. + a5 t^5) exp(-z^2), t = 1 / (1 + p z).  Since |x| = (1/t - 1) sqrt2 / p,
// 0.5 |x| (a1 t + .. + a5 t^5) is itself a degree-5 polynomial B(t) = (sqrt2 / 2p) (1 - t) (a1 + a2 t + .. + a5 t^4), whose
// coefficients are folded here: 2 MUFU + 10 FMA-pipe instructions per element (13 before the folding; erff, and even an IEEE
// 1/x, made the GEGLU epilogue slower than its main loop).  |abs err| < 5e-7, relative L2 3e-8 against erf in float64
// (tools/gelu_check.py) — far below the fp16 rounding of the result.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(ax, 0.3275911f * 0.70710678118654752f, 1.0f)));
  float b = fmaf(-2.2910481280905266f, t, 5.427682952378112f);
  b = fmaf(b, t, -6.204762423397693f);
  b = fmaf(b, t, 3.6822150125075983f);
  b = fmaf(b, t, -1.1641381704241662f);
  b = fmaf(b, t, 0.5500507570266749f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));  // exp(-x^2 / 2)
  return fmaf(-b, e, fmaxf(x, 0.0f));
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
// residuals may alias the output buffer (in-place update): coherent loads
__device__ __forceinline__ void load8_plain(const float* p, float (&v)[8]) {
  float4 a = reinterpret_cast<const float4*>(p)[0];
  float4 b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8_plain(const __half* p, float (&v)[8]) {
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  uint4 raw;
  __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}

// 16 consecutive values of a residual / row-vector operand (fp32 or fp16) at element offset `off`;
// when only 8 remain (N tail) the upper half is zero.  Coherent loads: residuals may alias the output.
__device__ __forceinline__ void load16(const void* base, int is_fp16, long long off, bool full, float (&v)[16]) {
  float lo[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) hi[i] = 0.f;
  if (is_fp16) {
    const __half* p = reinterpret_cast<const __half*>(base) + off;
    load8_plain(p, lo);
    if (full) load8_plain(p + 8, hi);
  } else {
    const float* p = reinterpret_cast<const float*>(base) + off;
    load8_plain(p, lo);
    if (full) load8_plain(p + 8, hi);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = lo[i]; v[8 + i] = hi[i]; }
}

__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&w)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]),
               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

// store 16 consecutive outputs at element offset `off`: one 32-byte sector (fp16) or two (fp32) per row
__device__ __forceinline__ void store16(const GemmEpilogue& ep, const float (&v)[16], long long off, bool wide_ok) {
  if (ep.out_fp16) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    __half* p = reinterpret_cast<__half*>(ep.out) + off;
#ifdef EVW_DEBUG_NO_STORE  // A/B experiment (tools/gpu_r2_call23.sh): the kernel with its arithmetic but without its output stores
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("" ::"r"(w[i]));
    if (ep.s2 != 12345.678f) return;
#endif
    if (wide_ok) {
      st_global_256(p, w);
    } else {
      *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(p + 8) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    if (ep.out_lo) {  // split-precision operand: the fp16 tail carries the rounding error of the head
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 hd = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        __half2 t = __floats2half2_rn(v[2 * i] - hd.x, v[2 * i + 1] - hd.y);
        w[i] = *reinterpret_cast<uint32_t*>(&t);
      }
      __half* pl = reinterpret_cast<__half*>(ep.out_lo) + off;
      if (wide_ok && ((reinterpret_cast<uintptr_t>(ep.out_lo) & 31) == 0)) {
        st_global_256(pl, w);
      } else {
        *reinterpret_cast<uint4*>(pl) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(pl + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
  } else {
    float* p = reinterpret_cast<float*>(ep.out) + off;
#ifdef EVW_DEBUG_NO_STORE
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("" ::"f"(v[i]));
    if (ep.s2 != 12345.678f) return;
#endif
    if (wide_ok) {
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[i]);
      st_global_256(p, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = __float_as_uint(v[8 + i]);
      st_global_256(p + 8, w);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    }
  }
}

// k2Cta: CTAs are launched as clusters of two (a CTA pair on one TPC) that work on m-adjacent tiles of the SAME n-tile.
// Protocol (CTA rank 0 = leader):
//   full[s]    lives in the leader: its producer arrives once with expect_tx = both CTAs' bytes, BOTH producers' TMA loads
//              (cp.async.bulk.tensor ... cta_group::2) complete_tx on it; only the leader's MMA warp waits on it
//   empty[s]   one per CTA, released by the leader's tcgen05.commit.cta_group::2 multicast to both CTAs
//   tfull[a]   one per CTA (multicast commit): each CTA's epilogue drains its own 128 TMEM lanes
//   tempty[a]  lives in the leader, count = 2 x epilogue warps: the peer's epilogue warps arrive remotely
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA loads of a CTA pair: destination in the executing CTA, completion on `cluster_bar` (the leader's barrier)
__device__ __forceinline__ void tma_load_5d_2cta(const CUtensorMap* m, uint32_t dst, uint32_t cluster_bar, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap* m, uint32_t dst, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Non-GEGLU epilogue of one 128 x BLOCK_N tile for this thread's row: columns [16 k, 16 k + 16) for k = cgrp, cgrp + NG, ...
//   out = s0 * acc + bias' [+ rowvec] [+ s1 * res1] [+ s2 * res2]        (bias' = s0 * bias, staged in shared memory)
// RV / R1 / R2 select the operands at compile time.  res1 (the HBM-resident residual stream) is software-pipelined over
// two register sets: the loads of step k + NG are issued before the TMEM wait of step k.  The broadcast row vector (one
// row per frame or batch element: L1-resident) and res2 (one layer) are loaded at the point of use.
// out_off / rv_off are element offsets of column n0 of this row (of the broadcast row); ncols = N - n0.
// (must be inlined: as a real call the tcgen05.ld results were consumed before they arrived — 40 of 80 GEMM tests failed)
// RT: res1 does not come from global memory through registers but from a shared-memory ring of 128-row x 32-column fp32
// boxes that warp 3 fills with TMA (128B swizzle) several boxes ahead: the residual stream is then fetched in full lines
// by the copy engine with 64 KB in flight per SM, instead of 32 uncoalesced 16-byte requests per warp instruction whose
// latency two epilogue warps per scheduler cannot hide (to_out at K = 320 ran at 2.2x its HBM bound).  Box j of a tile
// serves steps 2j (column group 0) and 2j + 1 (column group 1); res_box0 = index of the tile's first box in the ring.
struct ResRing {
  uint32_t base, full_bar, empty_bar;  // shared-memory addresses: boxes, full[slots], empty[slots]
  uint32_t slots_mask, shift, box0;
  int row, lane;
};
// ST: GroupNorm statistics of the output.  Each step's 16 columns x 32 rows of the warp are summed over the rows by
// recursive halving (lanes exchange half of their column sums at each of four stages: 15 shuffles per quantity instead of
// 80), lanes 2c / 2c + 1 then hold column c's sum / sum of squares; a segmented reduction over the columns (4 shuffles per
// quantity) leaves each group's part of the step in the lane of its first column, which stores it to a slot of its own
// (warp quarter, step, group within the step) in shared memory.  No atomics, fixed summation order: the per-tile sums are
// reproducible.  The caller adds a tile's slots up (double) and sends them to global memory once per tile and group.
constexpr int kStatsSub = 3;  // groups a 16-column step can touch with >= 8 channels per group
struct StatsCtx {
  float2* slots;  // shared memory: [16 steps][kStatsSub] of this warp quarter and accumulator stage
  FastDiv cg;     // channels per group
  int n0;         // first column of the tile
  int lane;
};
__device__ __forceinline__ void stats_halve(float (&s)[16], int w, int lane_bit, int lane) {
  // keep the half of the 2w live values selected by this lane's bit, add the partner's copy of the same half
  const bool hi = (lane & lane_bit) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < w) {
      const float keep = hi ? s[w + i] : s[i];
      const float send = hi ? s[i] : s[w + i];
      s[i] = keep + __shfl_xor_sync(0xffffffffu, send, lane_bit);
    }
  }
}
// TS: TMA-store epilogue.  A warp's 32 rows x 16 columns of one step are staged in the warp's own shared-memory slab
// (fp32: 64-byte rows, SWIZZLE_64B; fp16: 32-byte rows, SWIZZLE_32B — conflict-free 16-byte stores) and leave through one
// cp.async.bulk.tensor store issued by lane 0: the LSU / L1 path handled one scattered 32-byte sector per lane and store
// instruction and cost 8-46 % of the level-0 linears (profiles/r02x_gemm_bench_nostore.log); the copy engine drains the
// slab asynchronously and clips rows / columns outside the tensor.  Two slabs per warp: before a slab is rewritten lane 0
// waits until at most one of its store groups is still reading shared memory (every step commits a group, empty or not,
// so that "one pending" always means "the other slab").
struct StoreCtx {
  const CUtensorMap* map;
  uint32_t slab;   // shared-memory address of this warp's two slabs (2 x 2 KiB)
  uint32_t count;  // steps staged so far by this warp (slab parity)
  int col0;        // output column of the tile's first column
  int x, y, t, b;  // coordinates of the warp's first row
  int lane;
  bool valid;      // tile inside the problem (CTA pairs run a dummy tile when the m-tile count is odd)
};
template <bool kFp16>
__device__ __forceinline__ void stage_store16(StoreCtx& sc, const float (&v)[16], int col) {
  const uint32_t buf = sc.slab + (sc.count & 1u) * 2048u;
  if (sc.lane == 0) tc::bulk_wait_read<1>();
  __syncwarp();
  if constexpr (kFp16) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    const uint32_t row = buf + (uint32_t)sc.lane * 32u, sw = ((uint32_t)sc.lane >> 2) & 1u;  // SWIZZLE_32B: chunk ^= (row / 4) % 2
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((0u ^ sw) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((1u ^ sw) << 4)), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
  } else {
    const uint32_t row = buf + (uint32_t)sc.lane * 64u, sw = ((uint32_t)sc.lane >> 1) & 3u;  // SWIZZLE_64B: chunk ^= (row / 2) % 4
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((c ^ sw) << 4)), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                   "f"(v[4 * c + 2]), "f"(v[4 * c + 3]) : "memory");
  }
  tc::fence_proxy_async_smem();
  __syncwarp();
  if (sc.lane == 0) {
    if (sc.valid) tc::tma_store_5d(sc.map, buf, sc.col0 + col, sc.x, sc.y, sc.t, sc.b);
    tc::bulk_commit();
  }
  ++sc.count;
}

template <bool RV, bool R1, bool R2, bool RT, bool ST, bool TS, bool GL = false>
__device__ __forceinline__ void epi_plain(const GemmEpilogue& ep, uint32_t t_addr, const float* bs, int cgrp, int NG, int nsteps,
                                          bool row_ok, int ncols, long long out_off, long long rv_off, bool wide_ok,
                                          uint32_t bar_full, uint32_t phase, const ResRing& rr, const StatsCtx& sx, StoreCtx& sc) {
  auto load_r1 = [&](float (&x)[16], int k) {
    if constexpr (!R1 || RT) return;
    const int c = k * 16;
    if (!(row_ok && k < nsteps && c < ncols)) return;
    load16(ep.res1, ep.res1_fp16, out_off + c, c + 16 <= ncols, x);
  };
  auto step = [&](int k, const float (&cur)[16], float (&nxt)[16], int k_next) {
    uint32_t raw[16];
    __syncwarp();
    tmem_ld_32x32b_x16(t_addr + k * 16, raw);
    load_r1(nxt, k_next);
    float rs[16];
    if constexpr (RT) {  // this step's 16 residual columns of this thread's row from the staged box
      const uint32_t cnt = rr.box0 + (uint32_t)(k >> 1);
      const uint32_t slot = cnt & rr.slots_mask;
      mbar_wait(rr.full_bar + 8u * slot, (cnt >> rr.shift) & 1u);
      const uint32_t rowp = rr.base + slot * 16384u + (uint32_t)rr.row * 128u;
      const uint32_t sw = (uint32_t)(rr.row & 7);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const uint32_t chunk = ((uint32_t)((k & 1) * 4 + qd)) ^ sw;  // 128B swizzle: 16-byte chunk index XOR (row % 8)
        float4 t4;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t4.x), "=f"(t4.y), "=f"(t4.z), "=f"(t4.w) : "r"(rowp + chunk * 16u));
        rs[4 * qd] = t4.x; rs[4 * qd + 1] = t4.y; rs[4 * qd + 2] = t4.z; rs[4 * qd + 3] = t4.w;
      }
      // the box is released at the END of the step, after the loaded values have been consumed: arriving right after
      // issuing the loads let the next TMA write overtake them (a 16-byte chunk of a row, a few thousand times per 80 M
      // outputs — tools/rt_probe.py); either this data dependency or a fence.proxy.async before the arrive removes it
    }
    tmem_ld_wait();
    const int c = k * 16;
    float sv[16];
    if constexpr (ST) {
#pragma unroll
      for (int i = 0; i < 16; ++i) sv[i] = 0.f;
    }
    float v[16];
    if constexpr (TS) {  // rows outside the tensor are staged too (and clipped by the store): keep them defined
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    if (row_ok && c < ncols) {
      const bool full = c + 16 <= ncols;
      const float4* b4 = reinterpret_cast<const float4*>(bs + k * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t4 = b4[i];
        v[4 * i] = fmaf(__uint_as_float(raw[4 * i]), ep.s0, t4.x);
        v[4 * i + 1] = fmaf(__uint_as_float(raw[4 * i + 1]), ep.s0, t4.y);
        v[4 * i + 2] = fmaf(__uint_as_float(raw[4 * i + 2]), ep.s0, t4.z);
        v[4 * i + 3] = fmaf(__uint_as_float(raw[4 * i + 3]), ep.s0, t4.w);
      }
      if constexpr (R1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(ep.s1, RT ? rs[i] : cur[i], v[i]);
      }
      // row vector / second residual: loaded at the point of use, eight columns at a time (register pressure)
      if constexpr (RV) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (hf == 0 || full) {
            float rv[8];
            load8_plain(ep.rowvec + rv_off + c + 8 * hf, rv);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * hf + i] += rv[i];
          }
        }
      }
      if constexpr (R2) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (hf == 0 || full) {
            float r2[8];
            load8_plain(ep.res2 + out_off + c + 8 * hf, r2);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * hf + i] = fmaf(ep.s2, r2[i], v[8 * hf + i]);
          }
        }
      }
      if constexpr (GL) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
      }
      if constexpr (ST) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sv[i] = (full || i < 8) ? v[i] : 0.f;
      }
      if constexpr (!TS) {
        if (full) {
          store16(ep, v, out_off + c, wide_ok);
        } else {  // N tail: only the first 8 columns exist (N is a multiple of 8)
          float v8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v8[i] = v[i];
          if (ep.out_fp16) store8(reinterpret_cast<__half*>(ep.out) + out_off + c, v8);
          else store8(reinterpret_cast<float*>(ep.out) + out_off + c, v8);
        }
      }
    }
    if constexpr (TS) {
      if (c < ncols) {  // warp-uniform; columns past N are clipped by the store
        if (ep.out_fp16) stage_store16<true>(sc, v, c);
        else stage_store16<false>(sc, v, c);
      }
    }
    if constexpr (ST) {  // executed by every lane (rows outside the tensor contribute zeros)
      if (c < ncols) {   // warp-uniform
        float sq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) sq[i] = sv[i] * sv[i];
        stats_halve(sv, 8, 16, sx.lane); stats_halve(sq, 8, 16, sx.lane);
        stats_halve(sv, 4, 8, sx.lane);  stats_halve(sq, 4, 8, sx.lane);
        stats_halve(sv, 2, 4, sx.lane);  stats_halve(sq, 2, 4, sx.lane);
        stats_halve(sv, 1, 2, sx.lane);  stats_halve(sq, 1, 2, sx.lane);
        float ts = sv[0] + __shfl_xor_sync(0xffffffffu, sv[0], 1);
        float tq = sq[0] + __shfl_xor_sync(0xffffffffu, sq[0], 1);
        const int ci = sx.lane >> 1;  // lanes 2 ci and 2 ci + 1 hold column ci of this step
        uint32_t g, rem;              // group of the column and its position inside the group
        fd_divmod((uint32_t)(sx.n0 + c + ci), sx.cg, g, rem);
#pragma unroll
        for (int d = 1; d < 16; d *= 2) {  // add column ci + d while it belongs to the same group (and to this step)
          const float os = __shfl_down_sync(0xffffffffu, ts, 2 * d);
          const float oq = __shfl_down_sync(0xffffffffu, tq, 2 * d);
          const bool same = ci + d < 16 && rem + (uint32_t)d < sx.cg.d;
          ts += same ? os : 0.f;
          tq += same ? oq : 0.f;
        }
        if ((sx.lane & 1) == 0 && (ci == 0 || rem == 0)) {
          const uint32_t g0 = fd_div((uint32_t)(sx.n0 + c), sx.cg);
          sx.slots[k * kStatsSub + (int)(g - g0)] = make_float2(ts, tq);
        }
      }
    }
    if constexpr (RT) {  // 8 arrivals (one per epilogue warp) free the box for the residual producer
      const uint32_t cnt2 = rr.box0 + (uint32_t)(k >> 1);
      __syncwarp();
      if (rr.lane == 0) mbar_arrive(rr.empty_bar + 8u * (cnt2 & rr.slots_mask));
    }
  };
  // res1 runs TWO steps ahead over three register sets: with one step of look-ahead the epilogue of a K = 320 GEMM sat
  // in long-scoreboard stalls (30 % of its samples, profiles/r02c_ncu_gemm_epilogue.txt) — 8 warps x 2 KB in flight per SM
  // cannot cover the latency of the residual stream
  float A[16], B[16], C[16];
  load_r1(A, cgrp);
  load_r1(B, cgrp + NG);
  mbar_wait_relaxed(bar_full, phase);
  tc_fence_after();
  for (int k = cgrp; k < nsteps; k += 3 * NG) {
    step(k, A, C, k + 2 * NG);
    if (k + NG < nsteps) step(k + NG, B, A, k + 3 * NG);
    if (k + 2 * NG < nsteps) step(k + 2 * NG, C, B, k + 4 * NG);
  }
}

// kEpi selects the epilogue at compile time — 8 = GEGLU, else a bit set (1 = row vector, 2 = res1, 4 = res2) of the plain
// epilogue's operands: one kernel body per combination.  Compiled into one body they pushed each other into spills (every
// variant needs 66-144 registers on its own, all six together hit the 168-register ceiling with 150 B of spill traffic,
// and GEGLU beside them went from 0.52 to 0.69 ms).
constexpr int kEpiGeglu = 8;
constexpr int kStatsBytes = 2 * 4 * 16 * kStatsSub * 8;  // shared memory of the GroupNorm partial sums (kStats)
template <bool k2Cta, int kEpiWarps, int kEpi>
__global__ void __launch_bounds__(kCtrlThreads + 32 * kEpiWarps, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
               const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_bh,
               const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ CUtensorMap tmap_o,
               const __grid_constant__ KernelParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = P.num_stages;
  // weight bytes per stage in THIS CTA: the whole BLOCK_N x 64 tile, or half of it in a CTA pair
  const uint32_t b_tile_bytes = (uint32_t)P.block_n * kBlockK * (k2Cta ? 1 : 2);
  const uint32_t stage_bytes = kATileBytes + b_tile_bytes;
  constexpr bool kResTma = (kEpi & 16) != 0;  // res1 staged through shared memory by TMA (warp 3)
  constexpr bool kStats = (kEpi & 32) != 0;   // GroupNorm statistics of the output accumulated by the epilogue
  constexpr bool kStoreTma = (kEpi & 64) != 0;  // outputs leave through per-warp shared-memory slabs and TMA stores
  const uint32_t res_base = smem_base + stages * stage_bytes;  // kResTma: P.res_slots boxes of 16 KiB
  const uint32_t st_base = res_base + (kResTma ? (uint32_t)P.res_slots * 16384u : 0u);  // kStoreTma: 2 x 2 KiB per epilogue warp
  const uint32_t bar_base = st_base + (kStoreTma ? (uint32_t)kEpiWarps * 4096u : 0u);  // 8-byte barriers
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * stages + 4);
  auto res_full_bar = [&](int s) { return bar_base + 8u * (2 * stages + 4) + 16u + 8u * s; };
  auto res_empty_bar = [&](int s) { return bar_base + 8u * (2 * stages + 4) + 16u + 64u + 8u * s; };
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a0);
    tma_prefetch_desc(&tmap_a1);
    tma_prefetch_desc(&tmap_b);
    if (k2Cta) tma_prefetch_desc(&tmap_bh);
  }
  const uint32_t crank = k2Cta ? cluster_ctarank() : 0u;
  // virtual tile index v -> tile: plain mode walks tiles, cluster mode walks pairs (rank picks the m-tile of the pair;
  // an odd m-tile count leaves one phantom tile whose loads fall outside the tensor and whose rows are never stored)
  const int v_first = k2Cta ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int v_step = k2Cta ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int v_limit = k2Cta ? P.total_pairs : P.total_tiles;
  auto to_tile = [&](int v) {
    if (!k2Cta) return v;
    const int mp = v / P.n_tiles, nt = v - mp * P.n_tiles;
    return (2 * mp + (int)crank) * P.n_tiles + nt;
  };
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), k2Cta ? 2 * kEpiWarps : kEpiWarps);  // pair: both CTAs' epilogues release the leader's MMA
    }
    if (kResTma) {
      for (int r = 0; r < P.res_slots; ++r) {
        mbar_init(res_full_bar(r), 1);
        mbar_init(res_empty_bar(r), kEpiWarps);  // every epilogue warp reads each box once
      }
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    if (k2Cta) tmem_alloc_2cta(tmem_slot, 512);  // issued by the same warp of both CTAs of the pair
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  if (k2Cta) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int total_chunks = 0;
  for (int t = 0; t < P.num_taps; ++t) total_chunks += P.chunks[P.tap_src[t]];
  // tile index -> (n-tile, x-tile, y-tile, frame, batch)
  auto decode = [&](int tile, int& n_tile, int& tx, int& ty, int& tt, int& tb) {
    uint32_t m, a, b, c, d;
    fd_divmod((uint32_t)tile, P.fd_n_tiles, m, a);
    n_tile = (int)a;
    fd_divmod(m, P.fd_tiles_x, m, b);
    tx = (int)b;
    fd_divmod(m, P.fd_tiles_y, m, c);
    ty = (int)c;
    fd_divmod(m, P.fd_T, m, d);
    tt = (int)d;
    tb = (int)m;
  };

  if (kEpiWarps > 8) {
    // 640 threads launch with <= 96 registers each; the control warpgroup hands registers to the epilogue warpgroups
    // (4 x 32 x 56 + 512 x 104 = 60 416 <= 640 x 96)
    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
  }
  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
#ifdef EVW_DEBUG_SKIP_W
    int dbg_loads = 0;
#endif
    for (int v = v_first; v < v_limit; v += v_step) {
      const int tile = to_tile(v);
      int n_tile, tx, ty, tt, tb;
      decode(tile, n_tile, tx, ty, tt, tb);
      const int x0 = tx * P.bx, y0 = ty * P.by, n0 = n_tile * P.block_n;
      int kglob = 0;
      for (int tap = 0; tap < P.num_taps; ++tap) {
        const int src = P.tap_src[tap];
        const CUtensorMap* ma = src ? &tmap_a1 : &tmap_a0;
        const int cx = x0 + P.tap_dx[tap], cy = y0 + P.tap_dy[tap], ct = tt + P.tap_dt[tap];
        for (int kc = 0; kc < P.chunks[src]; ++kc, ++kglob) {
          mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);
          if (elect_one_sync()) {
            const uint32_t sa = smem_base + stage * stage_bytes;
            if (k2Cta) {
              // own activation tile + own half of the weight rows; both CTAs' bytes complete on the leader's barrier
              const uint32_t lead_full = mapa_cluster(full_bar(stage), 0);
              if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * stage_bytes);
              tma_load_5d_2cta(ma, sa, lead_full, kc * kBlockK, cx, cy, ct, tb);
              tma_load_2d_2cta(&tmap_bh, sa + kATileBytes, lead_full, kglob * kBlockK, n0 + (int)crank * (P.block_n >> 1));
            } else {
#ifdef EVW_DEBUG_SKIP_W  // timing experiment only (wrong results): after the ring's first pass the weight tile is not re-loaded
              if (dbg_loads >= stages) {
                mbar_arrive_expect_tx(full_bar(stage), kATileBytes);
                tma_load_5d(ma, sa, full_bar(stage), kc * kBlockK, cx, cy, ct, tb);
              } else
#endif
              {
                mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
                tma_load_5d(ma, sa, full_bar(stage), kc * kBlockK, cx, cy, ct, tb);
                tma_load_2d(&tmap_b, sa + kATileBytes, full_bar(stage), kglob * kBlockK, n0);
              }
            }
          }
          __syncwarp();
#ifdef EVW_DEBUG_SKIP_W
          ++dbg_loads;
#endif
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && (!k2Cta || crank == 0)) {
    // ===================== MMA issuer (the leader CTA of a pair) =====================
    const uint32_t idesc = make_idesc_f16(k2Cta ? 2 * kBlockM : kBlockM, P.block_n);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int v = v_first; v < v_limit; v += v_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kAccStride;
      for (int kc = 0; kc < total_chunks; ++kc) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint64_t da = make_desc_k_sw128(sa);
        const uint64_t db = make_desc_k_sw128(sa + kATileBytes);
        if (elect_one_sync()) {
          if (k2Cta) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_f16_ss_2cta(d_tmem, da + 2ull * k, db + 2ull * k, idesc, (kc | k) != 0);
            tc_commit_2cta(empty_bar(stage), (uint16_t)3);
            if (kc == total_chunks - 1) tc_commit_2cta(tfull_bar(acc), (uint16_t)3);
          } else {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_f16_ss(d_tmem, da + 2ull * k, db + 2ull * k, idesc, (kc | k) != 0);
            tc_commit(empty_bar(stage));
            if (kc == total_chunks - 1) tc_commit(tfull_bar(acc));
          }
        }
        __syncwarp();
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 3 && kResTma) {
    // ===================== residual producer: res1 boxes (128 rows x 32 fp32 columns) of every tile, in order =====================
    if (lane == 0) tma_prefetch_desc(&tmap_r);
    const int boxes = P.block_n >> 5;
    uint32_t cnt = 0;
    for (int v = v_first; v < v_limit; v += v_step) {
      const int tile = to_tile(v);
      int n_tile, tx, ty, tt, tb;
      decode(tile, n_tile, tx, ty, tt, tb);
      const int x0 = tx * P.bx, y0 = ty * P.by, n0 = n_tile * P.block_n;
      for (int j = 0; j < boxes; ++j, ++cnt) {
        const uint32_t slot = cnt & (uint32_t)(P.res_slots - 1);
        mbar_wait_relaxed(res_empty_bar(slot), ((cnt >> P.res_shift) & 1u) ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(res_full_bar(slot), 16384u);
          tma_load_5d(&tmap_r, res_base + slot * 16384u, res_full_bar(slot), n0 + 32 * j, x0, y0, tt, tb);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // Two warps share each TMEM lane quarter and alternate 16-column steps.  Per step: issue the TMEM load, issue the
    // global loads of the NEXT step's residual / row-vector operands (software pipeline over two register sets: their
    // latency overlaps this step's work), wait for TMEM, one FFMA per output against the pre-scaled bias staged in shared
    // memory (128-bit reads), store one full 32-byte sector per row where alignment allows.  The hot loop is compiled
    // per operand combination (epi_plain<RV, R1, R2>): with two epilogue warps per scheduler the loop is bound by issue
    // slots, and the generic version spent 4/5 of them on moves, predicates and index arithmetic.
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;         // row of the 128-row tile
    const int cgrp = (warp - 4) >> 2;    // column group of this warp
    constexpr int NG = kEpiWarps / 4;
    const GemmEpilogue& ep = P.ep;
    const int ldo = ep.geglu ? P.N / 2 : P.N;
    const int rv_ld = ep.rv_ld ? ep.rv_ld : ldo;
    const int esz = ep.out_fp16 ? 2 : 4;
    const bool wide_ok = (((long long)ldo * esz) % 32 == 0) && ((reinterpret_cast<uintptr_t>(ep.out) & 31) == 0);
    float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + 8 * (2 * stages + 4) + 16 + 128);
    // kStats: [2 accumulator stages][4 warp quarters][16 steps][kStatsSub] partial group sums of a tile
    float2* stats_s = reinterpret_cast<float2*>(bias_s + 2 * 256);
    const int et = threadIdx.x - 128;    // 0 .. 32*kEpiWarps-1
    int prev_inst = -1, prev_gfirst = 0, prev_ng = 0, prev_n0 = 0, prev_cols = 0;  // kStats: the tile still in shared memory
    auto flush_stats = [&](int acc_prev) {  // after a barrier that follows the tile's last step; one thread per group
      if (prev_inst >= 0 && et < prev_ng) {
        const int g = prev_gfirst + et, cg = (int)P.fd_gn_cg.d;
        const int lo = max(g * cg, prev_n0) - prev_n0, hi = min((g + 1) * cg, prev_n0 + prev_cols) - 1 - prev_n0;
        double s = 0.0, sq = 0.0;
        for (int k = lo >> 4; k <= hi >> 4; ++k) {
          const int sub = g - (int)fd_div((uint32_t)(prev_n0 + 16 * k), P.fd_gn_cg);
          for (int qq = 0; qq < 4; ++qq) {
            const float2 v = stats_s[((acc_prev * 4 + qq) * 16 + k) * kStatsSub + sub];
            s += (double)v.x;
            sq += (double)v.y;
          }
        }
        double* dst = ep.gn_stats + ((long long)prev_inst * 32 + g) * 2;
        atomicAdd(dst, s);
        atomicAdd(dst + 1, sq);
      }
    };
    const int mx = r & (P.bx - 1), my = r >> P.bx_shift;  // bx is a power of two
    constexpr bool is_geglu = (kEpi & ~64) == kEpiGeglu;
    // residual operands stream from DRAM exactly once: pull the row segment of the NEXT tile into L2 one tile ahead
    // (prefetch.global.L2, no registers / shared memory), so the epilogue's loads find it there
    auto prefetch_residuals = [&](int v_p) {
      if (v_p >= v_limit || ((kResTma || !ep.res1) && !ep.res2)) return;
      const int tile_p = to_tile(v_p);
      if (tile_p >= P.total_tiles) return;
      int n_tile_p, tx_p, ty_p, tt_p, tb_p;
      decode(tile_p, n_tile_p, tx_p, ty_p, tt_p, tb_p);
      const int gx_p = tx_p * P.bx + mx, gy_p = ty_p * P.by + my;
      if (gx_p >= P.X || gy_p >= P.Y) return;
      const long long row_p = (((long long)tb_p * P.T + tt_p) * P.Y + gy_p) * P.X + gx_p;
      const long long off = row_p * ldo + (long long)n_tile_p * P.block_n;
      const int cols = min(P.block_n, P.N - n_tile_p * P.block_n);
      if (ep.res1 && !kResTma) {
        const int e1 = ep.res1_fp16 ? 2 : 4;
        const char* b = reinterpret_cast<const char*>(ep.res1) + off * e1;
        for (int o = cgrp * 128; o < cols * e1; o += 128 * NG) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
      }
      if (ep.res2) {
        const char* b = reinterpret_cast<const char*>(ep.res2) + off * 4;
        for (int o = cgrp * 128; o < cols * 4; o += 128 * NG) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
      }
    };
    prefetch_residuals(v_first);
    StoreCtx sc;
    sc.map = &tmap_o; sc.slab = st_base + (uint32_t)(warp - 4) * 4096u; sc.count = 0; sc.lane = lane;
    sc.col0 = 0; sc.x = sc.y = sc.t = sc.b = 0; sc.valid = false;
    if (kStoreTma && warp == 4 && lane == 0) tma_prefetch_desc(&tmap_o);
    // first row of this warp's 32 rows inside the tile: (x, y) offset of row 32 q
    const int wx = (q * 32) & (P.bx - 1), wy = (q * 32) >> P.bx_shift;
    int it = 0;
    for (int v = v_first; v < v_limit; v += v_step, ++it) {
      const int tile = to_tile(v);
      prefetch_residuals(v + v_step);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int n_tile, tx, ty, tt, tb;
      decode(tile, n_tile, tx, ty, tt, tb);
      const int gx = tx * P.bx + mx, gy = ty * P.by + my;
      const bool row_ok = gx < P.X && gy < P.Y && tile < P.total_tiles;
      const long long row = (((long long)tb * P.T + tt) * P.Y + gy) * P.X + gx;
      const int n0 = n_tile * P.block_n;
      float* bs = bias_s + acc * 256;
      if (et < P.block_n) {
        float bv = (ep.bias && n0 + et < P.N) ? __ldg(ep.bias + n0 + et) : 0.f;
        // s0 is folded into the staged bias: out = s0 * acc + s0 * bias (GEGLU: value columns only)
        if (!is_geglu || (et & 16) == 0) bv *= ep.s0;
        bs[et] = bv;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      StatsCtx sx;
      if (kStats) {
        flush_stats(acc ^ 1);  // every warp has finished the previous tile (barrier above)
        const int ncols_t = min(P.block_n, P.N - n0);
        sx.cg = P.fd_gn_cg; sx.n0 = n0; sx.lane = lane;
        sx.slots = stats_s + (acc * 4 + q) * 16 * kStatsSub;
        // all rows of a tile belong to one GroupNorm instance (checked by gemm_enable_gn_stats)
        const uint32_t row0 = (uint32_t)((((long long)tb * P.T + tt) * P.Y + ty * P.by) * P.X + tx * P.bx);
        prev_inst = (tile < P.total_tiles) ? (int)fd_div(row0, P.fd_gn_rpi) : -1;
        prev_gfirst = (int)fd_div((uint32_t)n0, P.fd_gn_cg);
        prev_ng = (int)fd_div((uint32_t)(n0 + ncols_t - 1), P.fd_gn_cg) - prev_gfirst + 1;
        prev_n0 = n0;
        prev_cols = ncols_t;
      }

      if (kStoreTma) {
        sc.col0 = is_geglu ? n0 / 2 : n0;
        sc.x = tx * P.bx + wx; sc.y = ty * P.by + wy; sc.t = tt; sc.b = tb;
        sc.valid = tile < P.total_tiles;
      }
      const uint32_t t_addr = tmem_base + acc * kAccStride + ((uint32_t)(q * 32) << 16);
      if constexpr (is_geglu) {
        mbar_wait_relaxed(tfull_bar(acc), acc_phase);
        tc_fence_after();
        // 32 accumulator columns = [16 value | 16 gate] -> 16 outputs.  (Issuing the tensor-memory load of step k + 1 before
        // the arithmetic of step k changed nothing — 0.493 vs 0.490 ms at 258 048 x 2560 x 320, profiles/r02t_geglu_prefetch.log:
        // the epilogue is bound by its ~20 instructions per output, not by the load latency.)
        for (int k = cgrp; k * 32 < P.block_n; k += NG) {
          uint32_t raw[32];
          __syncwarp();
          tmem_ld_32x32b_x32(t_addr + k * 32, raw);
          tmem_ld_wait();
          const int nh = n0 + k * 32;
          if ((kStoreTma || row_ok) && nh < P.N) {  // TMA store: every lane stages its row, the store clips
            float vv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float hv = fmaf(__uint_as_float(raw[i]), ep.s0, bs[k * 32 + i]);  // s0 (value + bias)
              const float gv = __uint_as_float(raw[16 + i]) + bs[k * 32 + 16 + i];
              vv[i] = hv * gelu_erf(gv);
            }
            if constexpr (kStoreTma) stage_store16<true>(sc, vv, k * 16);
            else store16(ep, vv, row * ldo + nh / 2, wide_ok);
          }
        }
      } else {
        const long long out_off = row * ldo + n0;
        long long rv_off = 0;
        if (ep.rowvec) {
          const uint32_t rq = fd_div((uint32_t)(row + ep.rv_row0), P.fd_rv_div);
          uint32_t qq, rem;
          fd_divmod(rq, P.fd_rv_mod, qq, rem);
          rv_off = (long long)rem * rv_ld + n0;
        }
        const int nsteps = P.block_n / 16;
        const int ncols = P.N - n0;  // valid columns of this tile (may exceed block_n)
        const uint32_t bar_full = tfull_bar(acc);
        ResRing rr;
        rr.base = res_base; rr.full_bar = res_full_bar(0); rr.empty_bar = res_empty_bar(0);
        rr.slots_mask = (uint32_t)(P.res_slots - 1); rr.shift = (uint32_t)P.res_shift;
        rr.box0 = (uint32_t)it * (uint32_t)(P.block_n >> 5); rr.row = r; rr.lane = lane;
        epi_plain<(kEpi & 1) != 0, (kEpi & 2) != 0, (kEpi & 4) != 0, kResTma, kStats, kStoreTma, (kEpi & 128) != 0>(
            ep, t_addr, bs, cgrp, NG, nsteps, row_ok, ncols, out_off, rv_off, wide_ok, bar_full, acc_phase, rr, sx, sc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (k2Cta) mbar_arrive_cluster(mapa_cluster(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
      }
    }
    if (kStats) {  // the last tile's sums
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      flush_stats((it - 1) & 1);
    }
    if (kStoreTma && lane == 0) tc::bulk_wait<0>();  // this warp's stores are complete before the CTA's shared memory goes away
  }

  tc_fence_before();
  __syncthreads();
  if (k2Cta) cluster_sync_all();  // no CTA leaves while its peer may still arrive on its barriers or read its operands
  if (warp == 2) {
    tc_fence_after();
    if (k2Cta) tmem_dealloc_2cta(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

}  // namespace

static int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B);
int encode_tmap_swz(CUtensorMap* out, int fp16, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes) {
  const CUtensorMapSwizzle swz = swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                               : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  return encode_tmap(out, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes,
                     box, swz);
}
int encode_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box);
}
int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}
static int encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return EVW_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu box %u %u %u base %p", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, base);
    return EVW_ERR_CUDA;
  }
  return EVW_OK;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

// Shape of a 128-row tile on the image: bx x by pixels, powers of two.  Token rows (Y == 1): one 128-row strip.  Otherwise
// the shape that wastes the fewest tile rows on partial tiles, the widest on ties — unchanged (128, or the width itself) for
// the power-of-two widths of the UNet / VAE, but e.g. 32 x 4 instead of 128 x 1 for VGGT's 148-pixel-wide DPT maps (92 %
// instead of 58 % of the MMA rows valid).  Tiles narrower than 32 pixels lose the TMA-store epilogue: they must win by > 15 %;
// no tile is wider than the next power of two above the image width.
static int choose_bx(int X, int Y) {
  if (Y == 1) return kBlockM;
  int top = 8;
  while (top < X && top < kBlockM) top <<= 1;
  int best = top;
  double best_score = -1.0;
  for (int bx = top; bx >= 8; bx >>= 1) {
    const int by = kBlockM / bx;
    const long long tx = (X + bx - 1) / bx, ty = (Y + by - 1) / by;
    const double util = (double)X * Y / ((double)tx * bx * ty * by);
    const double score = util * (bx >= 32 ? 1.0 : 0.85);
    if (score > best_score + 1e-9) {
      best_score = score;
      best = bx;
    }
  }
  return best;
}

int gemm_cluster_mode();
int gemm_pair_min_k();
int gemm_store_tma_mode();

int gemm_plan(GemmOp* op, const GemmProblem& pr) {
  EVW_CHECK_ARG(pr.C0 > 0 && pr.C0 % kBlockK == 0, "gemm: C0=%d must be a positive multiple of 64", pr.C0);
  EVW_CHECK_ARG(pr.C1 % kBlockK == 0, "gemm: C1=%d must be a multiple of 64", pr.C1);
  EVW_CHECK_ARG(pr.N > 0 && pr.N % 8 == 0, "gemm: N=%d must be a multiple of 8", pr.N);
  EVW_CHECK_ARG(pr.num_taps >= 1 && pr.num_taps <= kMaxTaps, "gemm: bad tap count %d", pr.num_taps);
  EVW_CHECK_ARG(pr.X > 0 && pr.Y > 0 && pr.T > 0 && pr.B > 0, "gemm: bad extents");
  EVW_CHECK_ARG(pr.a0 && pr.w && pr.ep.out, "gemm: null operand");
  EVW_CHECK_ARG(((uintptr_t)pr.a0 & 15) == 0 && ((uintptr_t)pr.w & 15) == 0 && ((uintptr_t)pr.ep.out & 15) == 0,
                "gemm: operands must be 16-byte aligned");
  KernelParams P{};
  P.ep = pr.ep;
  P.X = pr.X; P.Y = pr.Y; P.T = pr.T; P.B = pr.B; P.N = pr.N;
  P.bx = choose_bx(pr.X, pr.Y);  // token rows: a single 128-row strip (rows past X are zero-filled)
  P.by = kBlockM / P.bx;
  P.tiles_x = (pr.X + P.bx - 1) / P.bx;
  P.tiles_y = (pr.Y + P.by - 1) / P.by;
  int bn = pr.block_n;
  if (bn <= 0) {
    // largest tile width that divides N evenly among {256,160,128,...}; GEGLU needs a multiple of 32
    const int cands[] = {256, 160, 128, 192, 96, 64, 32, 16};
    bn = 0;
    for (int c : cands) {
      if (pr.N % c == 0 && (!pr.ep.geglu || c % 32 == 0)) { bn = c; break; }
    }
    if (bn == 0) bn = pr.ep.geglu ? 128 : (pr.N < 128 ? ((pr.N + 15) / 16) * 16 : 128);
  }
  EVW_CHECK_ARG(bn % 16 == 0 && bn >= 16 && bn <= 256, "gemm: BLOCK_N=%d invalid", bn);
  EVW_CHECK_ARG(!pr.ep.geglu || (bn % 32 == 0 && pr.N % 32 == 0), "gemm: GEGLU needs N and BLOCK_N multiples of 32");
  EVW_CHECK_ARG(!pr.ep.res2 || pr.ep.res1, "gemm: res2 needs res1 (the two-residual epilogue loads both)");
  EVW_CHECK_ARG(!pr.ep.act || (pr.ep.act == 1 && !pr.ep.geglu && !pr.ep.res1 && !pr.ep.rowvec && !pr.ep.out_lo && !pr.ep.gn_stats),
                "gemm: the GELU epilogue takes no row vector, residual, split output or GroupNorm statistics");
  EVW_CHECK_ARG(!pr.ep.out_lo || (pr.ep.out_fp16 && !pr.ep.geglu && pr.N % 16 == 0 && ((uintptr_t)pr.ep.out_lo & 15) == 0),
                "gemm: out_lo needs an fp16, non-GEGLU output with N a multiple of 16");
  P.block_n = bn;
  P.n_tiles = (pr.N + bn - 1) / bn;
  {
    const long long rows_total = (long long)pr.B * pr.T * pr.Y * pr.X;
    EVW_CHECK_ARG(pr.ep.rv_row0 >= 0 && rows_total + pr.ep.rv_row0 < (1ll << 31), "gemm: rv_row0 out of range");
    EVW_CHECK_ARG(rows_total < (1ll << 31) && pr.ep.rv_div < (1ll << 31) && pr.ep.rv_mod < (1ll << 31) && pr.ep.rv_div > 0 &&
                      pr.ep.rv_mod > 0,
                  "gemm: more than 2^31 rows (or a broadcast period out of range)");
    P.fd_n_tiles = make_fastdiv((uint32_t)P.n_tiles);
    P.fd_tiles_x = make_fastdiv((uint32_t)P.tiles_x);
    P.fd_tiles_y = make_fastdiv((uint32_t)P.tiles_y);
    P.fd_T = make_fastdiv((uint32_t)pr.T);
    P.fd_rv_div = make_fastdiv((uint32_t)pr.ep.rv_div);
    P.fd_rv_mod = make_fastdiv((uint32_t)pr.ep.rv_mod);
    P.bx_shift = 0;
    while ((1 << P.bx_shift) < P.bx) ++P.bx_shift;
  }
  P.num_taps = pr.num_taps;
  P.chunks[0] = pr.C0 / kBlockK;
  P.chunks[1] = pr.C1 / kBlockK;
  long long ktot = 0;
  for (int t = 0; t < pr.num_taps; ++t) {
    P.tap_dx[t] = pr.tap_dx[t]; P.tap_dy[t] = pr.tap_dy[t]; P.tap_dt[t] = pr.tap_dt[t]; P.tap_src[t] = pr.tap_src[t];
    EVW_CHECK_ARG(pr.tap_src[t] == 0 || (pr.tap_src[t] == 1 && pr.a1 && pr.C1 > 0), "gemm: tap %d uses a missing source", t);
    ktot += pr.tap_src[t] ? pr.C1 : pr.C0;
  }
  EVW_CHECK_ARG(ktot == pr.K_total, "gemm: K_total=%lld but taps cover %lld", (long long)pr.K_total, ktot);
  P.total_tiles = P.tiles_x * P.tiles_y * P.T * P.B * P.n_tiles;
  const int m_tiles = P.tiles_x * P.tiles_y * P.T * P.B;
  P.total_pairs = ((m_tiles + 1) / 2) * P.n_tiles;
  int sms = sm_count();
  // CTA-pair (cta_group::2) mode needs at least one full pair and a 1 KiB-aligned half weight tile (bn % 16 == 0 holds)
  // Auto mode pairs CTAs only where the main loop dominates: with K_total < gemm_pair_min_k() (K = 320 / 640 linears, whose
  // tiles are epilogue-bound) the pair's lock-step accumulator hand-over costs more than the halved weight traffic saves
  // (profiles/r02b_gemm_bench.log: 320->320 linear 596 -> 457 TFLOP/s, 3x3 convolutions 935 -> 1172 / 956 -> 1269).
  const int mode = gemm_cluster_mode();
  op->cluster = ((mode == 1 || (mode == 2 && pr.K_total >= gemm_pair_min_k())) && m_tiles >= 2 && sms >= 2) ? 1 : 0;
  const uint32_t stage_bytes = kATileBytes + bn * kBlockK * (op->cluster ? 1 : 2);  // a pair member holds half the weight tile
  // TMA-staged residual: fp32 res1, whole 32-column boxes, a row pitch the tensor map accepts; EVW_GEMM_RES_TMA=0 disables
  static const bool res_tma_enabled = [] { const char* e = getenv("EVW_GEMM_RES_TMA"); return !(e && atoi(e) == 0); }();
  op->res_tma = (res_tma_enabled && pr.ep.res1 && !pr.ep.res1_fp16 && !pr.ep.geglu && bn % 32 == 0 && pr.N % 4 == 0 &&
                 ((uintptr_t)pr.ep.res1 & 15) == 0) ? 1 : 0;
  P.res_slots = op->res_tma ? 4 : 0;  // 64 KiB of residual in flight per SM
  P.res_shift = 2;
  // operand stages + residual ring + barriers, TMEM slot, residual barriers + staged bias + GroupNorm partial sums + 1 KiB
  // of alignment slack: as many stages (<= 8) as fit into the 227 KiB a CTA may use
  // TMA-store epilogue (EVW_GEMM_STORE_TMA=0 disables): one output, 16-byte aligned rows, and every warp's 32 tile rows must be
  // a box of the output's (x, y) grid — 32 consecutive x positions, or whole image rows when the tile spans the image width
  // EVW_GEMM_STORE_TMA: 0 = never, 1 = wherever possible, 2 = only the single-CTA (K_total < pair threshold) GEMMs,
  // 3 = only fp32 outputs (A/B runs)
  const int store_tma_mode = gemm_store_tma_mode();
  {
    const int ldo = pr.ep.geglu ? pr.N / 2 : pr.N;
    const int esz = pr.ep.out_fp16 ? 2 : 4;
    const bool rows_ok = P.bx >= 32 || (P.tiles_x == 1 && P.bx == pr.X && 32 % P.bx == 0);
    const bool strided_req = pr.out_sx != 1 || pr.out_sy != 1 || pr.out_ox != 0 || pr.out_oy != 0;
    bool want = store_tma_mode != 0 || strided_req;  // a strided output view exists only on the TMA-store path
    if (store_tma_mode == 2 && op->cluster && !strided_req) want = false;
    if (store_tma_mode == 3 && pr.ep.out_fp16 && !strided_req) want = false;
    op->store_tma = (want && !pr.ep.out_lo && rows_ok && ((long long)ldo * esz) % 16 == 0 && ldo % 16 == 0 &&
                     (!pr.ep.geglu || bn % 32 == 0)) ? 1 : 0;
  }
  auto smem_for = [&](int s) {
    return (long long)s * stage_bytes + P.res_slots * 16384 + (op->store_tma ? kEpiWarpsDefault * 4096 : 0) + 8 * (2 * s + 4) + 16 +
           128 + 2 * 256 * 4 + kStatsBytes + 1024;
  };
  int stages = 8;
  while (stages > 1 && smem_for(stages) > 227 * 1024) --stages;
  EVW_CHECK_ARG(stages >= 2, "gemm: not enough shared memory for BLOCK_N=%d", bn);
  P.num_stages = stages;
  op->smem_bytes = (int)smem_for(stages);
  if (op->cluster) {
    const int want = 2 * P.total_pairs;
    op->grid = want < (sms & ~1) ? want : (sms & ~1);
  } else {
    op->grid = P.total_tiles < sms ? P.total_tiles : sms;
  }
  static_assert(sizeof(KernelParams) <= sizeof(op->params), "GemmOp::params too small");
  memcpy(op->params, &P, sizeof(P));

  const int Tm = pr.Tmap > 0 ? pr.Tmap : pr.T;
  uint64_t dims[5] = {(uint64_t)pr.C0, (uint64_t)pr.X, (uint64_t)pr.Y, (uint64_t)Tm, (uint64_t)pr.B};
  uint64_t str[4] = {(uint64_t)pr.C0 * 2, (uint64_t)pr.C0 * 2 * pr.X, (uint64_t)pr.C0 * 2 * pr.X * pr.Y,
                     (uint64_t)pr.C0 * 2 * pr.X * pr.Y * Tm};
  uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)P.bx, (uint32_t)P.by, 1, 1};
  int rc = encode_tmap_f16(reinterpret_cast<CUtensorMap*>(op->tmap_a0), pr.a0, 5, dims, str, box);
  if (rc) return rc;
  if (pr.a1 && pr.C1 > 0) {
    EVW_CHECK_ARG(((uintptr_t)pr.a1 & 15) == 0, "gemm: a1 must be 16-byte aligned");
    dims[0] = pr.C1;
    str[0] = (uint64_t)pr.C1 * 2; str[1] = str[0] * pr.X; str[2] = str[1] * pr.Y; str[3] = str[2] * Tm;
    rc = encode_tmap_f16(reinterpret_cast<CUtensorMap*>(op->tmap_a1), pr.a1, 5, dims, str, box);
    if (rc) return rc;
  } else {
    memcpy(op->tmap_a1, op->tmap_a0, sizeof(op->tmap_a0));
  }
  uint64_t bdims[2] = {(uint64_t)pr.K_total, (uint64_t)pr.N};
  uint64_t bstr[1] = {(uint64_t)pr.K_total * 2};
  uint32_t bbox[2] = {(uint32_t)kBlockK, (uint32_t)bn};
  rc = encode_tmap_f16(reinterpret_cast<CUtensorMap*>(op->tmap_b), pr.w, 2, bdims, bstr, bbox);
  if (rc) return rc;
  uint32_t hbox[2] = {(uint32_t)kBlockK, (uint32_t)(bn / 2)};
  rc = encode_tmap_f16(reinterpret_cast<CUtensorMap*>(op->tmap_bh), pr.w, 2, bdims, bstr, hbox);
  if (rc) return rc;
  if (op->res_tma) {
    uint64_t rdims[5] = {(uint64_t)pr.N, (uint64_t)pr.X, (uint64_t)pr.Y, (uint64_t)pr.T, (uint64_t)pr.B};
    uint64_t rstr[4] = {(uint64_t)pr.N * 4, (uint64_t)pr.N * 4 * pr.X, (uint64_t)pr.N * 4 * pr.X * pr.Y,
                        (uint64_t)pr.N * 4 * pr.X * pr.Y * pr.T};
    uint32_t rbox[5] = {32u, (uint32_t)P.bx, (uint32_t)P.by, 1, 1};
    rc = encode_tmap_f32(reinterpret_cast<CUtensorMap*>(op->tmap_r), pr.ep.res1, 5, rdims, rstr, rbox);
    if (rc) return rc;
  } else {
    memcpy(op->tmap_r, op->tmap_a0, sizeof(op->tmap_a0));
  }
  const bool strided_out = pr.out_sx != 1 || pr.out_sy != 1 || pr.out_ox != 0 || pr.out_oy != 0;
  op->strided_out = strided_out ? 1 : 0;
  EVW_CHECK_ARG(!strided_out || (op->store_tma && pr.out_sx >= 1 && pr.out_sy >= 1 && pr.out_ox >= 0 && pr.out_ox < pr.out_sx &&
                                 pr.out_oy >= 0 && pr.out_oy < pr.out_sy && !pr.ep.res1 && !pr.ep.res2 && !pr.ep.geglu),
                "gemm: a strided output view needs the TMA-store epilogue (no residuals, no GEGLU)");
  if (op->store_tma) {
    const uint64_t ldo = pr.ep.geglu ? pr.N / 2 : pr.N, esz = pr.ep.out_fp16 ? 2 : 4;
    const uint64_t sx = (uint64_t)pr.out_sx, sy = (uint64_t)pr.out_sy;
    const uint64_t px = ldo * esz;                // bytes per output pixel
    const uint64_t line = px * sx * pr.X;         // bytes per output image row
    uint64_t odims[5] = {ldo, (uint64_t)pr.X, (uint64_t)pr.Y, (uint64_t)pr.T, (uint64_t)pr.B};
    uint64_t ostr[4] = {px * sx, line * sy, line * sy * pr.Y, line * sy * pr.Y * pr.T};
    const uint32_t bxw = P.bx >= 32 ? 32u : (uint32_t)P.bx;
    uint32_t obox[5] = {16u, bxw, 32u / bxw, 1, 1};
    const char* obase = reinterpret_cast<const char*>(pr.ep.out) + (uint64_t)pr.out_oy * line + (uint64_t)pr.out_ox * px;
    rc = encode_tmap_swz(reinterpret_cast<CUtensorMap*>(op->tmap_o), pr.ep.out_fp16, obase, 5, odims, ostr, obox,
                         pr.ep.out_fp16 ? 32 : 64);
    if (rc) return rc;
  } else {
    memcpy(op->tmap_o, op->tmap_a0, sizeof(op->tmap_a0));
  }
  op->flops = 2.0 * pr.X * pr.Y * pr.T * pr.B * (double)pr.N * (double)pr.K_total;
  return EVW_OK;
}

bool gemm_strided_out_ok(int X, int Y, int N) {
  const int bx = choose_bx(X, Y);
  const int tiles_x = (X + bx - 1) / bx;
  const bool rows_ok = bx >= 32 || (tiles_x == 1 && bx == X && 32 % bx == 0);
  return rows_ok && N % 16 == 0;
}

void upconv2x_phase(GemmProblem& pr, int phase) {
  const int py = phase >> 1, px = phase & 1;
  pr.num_taps = 4;
  int t = 0;
  for (int iy = 0; iy < 2; ++iy)
    for (int ix = 0; ix < 2; ++ix, ++t) {
      pr.tap_dy[t] = (int8_t)(py == 0 ? iy - 1 : iy);
      pr.tap_dx[t] = (int8_t)(px == 0 ? ix - 1 : ix);
      pr.tap_dt[t] = 0;
      pr.tap_src[t] = 0;
    }
  pr.K_total = 4LL * pr.C0;
  pr.out_sx = 2; pr.out_sy = 2; pr.out_ox = px; pr.out_oy = py;
}

// -1: EVW_GEMM_GN_STATS (default on), 0 / 1: forced (tests, A/B runs); read when a GroupNorm asks its producer for statistics
static int g_gemm_gn_stats = -1;
static bool gemm_gn_stats_enabled() {
  if (g_gemm_gn_stats < 0) {
    const char* e = getenv("EVW_GEMM_GN_STATS");
    g_gemm_gn_stats = (e && atoi(e) == 0) ? 0 : 1;
  }
  return g_gemm_gn_stats != 0;
}
void set_gemm_gn_stats(int on) { g_gemm_gn_stats = on < 0 ? -1 : (on ? 1 : 0); }

// Ask a planned GEMM to accumulate the GroupNorm(32) statistics of its output (see GemmEpilogue::gn_stats).
// Returns 0 when enabled, 1 when this GEMM cannot provide them (the caller keeps its own statistics pass).
int gemm_enable_gn_stats(GemmOp* op, double* stats, long long rows_per_inst) {
  if (!gemm_gn_stats_enabled()) return 1;
  KernelParams P;
  memcpy(&P, op->params, sizeof(P));
  const long long rows = (long long)P.B * P.T * P.Y * P.X;
  if (P.ep.geglu || P.ep.res2 || (P.ep.res1 && !op->res_tma) || P.N % 32 != 0 || !stats || rows_per_inst <= 0 || op->strided_out) return 1;
  const int cg = P.N / 32;
  if (cg < 8) return 1;  // a 16-column step may touch at most kStatsSub groups
  if (rows % rows_per_inst != 0) return 1;
  const bool flat = P.Y == 1 && P.T == 1 && P.B == 1;  // token rows: a tile is 128 consecutive rows
  if (flat ? (rows_per_inst % kBlockM != 0) : (rows_per_inst % ((long long)P.X * P.Y) != 0)) return 1;
  static const int plain_wide = [] { const char* e = getenv("EVW_GEMM_PLAIN_WARPS"); return (e && atoi(e) == 16) ? 1 : 0; }();
  if (plain_wide) return 1;
  P.ep.gn_stats = stats; P.ep.gn_cg = cg;
  P.gn_rows_per_inst = rows_per_inst;
  P.fd_gn_cg = make_fastdiv((uint32_t)cg);
  P.fd_gn_rpi = make_fastdiv((uint32_t)rows_per_inst);
  memcpy(op->params, &P, sizeof(P));
  return 0;
}

// -1: EVW_GEMM_CLUSTER, default ON = CTA pairs with tcgen05.mma.cta_group::2 (M = 256).  Round 1's cluster mode only
// multicast the weight tile between two cta_group::1 CTAs and measured neutral (profiles/r01f_gemm_bench.log): the
// 160-wide tiles are bound by the shared-memory data pipe (tensor-core operand reads + TMA writes), which multicast does
// not relieve; halving the weight operand per CTA does (profiles/r02*_gemm_bench.log).  0 = independent CTAs.
// returns 0 = never pair, 1 = always pair (forced: tests, A/B runs), 2 = auto (pair when K_total >= gemm_pair_min_k())
static int g_gemm_cluster = -1;
int gemm_cluster_mode() {
  if (g_gemm_cluster < 0) {
    const char* e = getenv("EVW_GEMM_CLUSTER");
    g_gemm_cluster = !e ? 2 : (atoi(e) == 0 ? 0 : (atoi(e) == 1 ? 1 : 2));
  }
  return g_gemm_cluster;
}
int gemm_pair_min_k() {
  static const int k = [] { const char* e = getenv("EVW_GEMM_PAIR_MIN_K"); return e ? atoi(e) : 1024; }();
  return k;
}
static int g_gemm_store_tma = -1;
int gemm_store_tma_mode() {
  if (g_gemm_store_tma < 0) {
    const char* e = getenv("EVW_GEMM_STORE_TMA");
    g_gemm_store_tma = e ? atoi(e) : 1;
  }
  return g_gemm_store_tma;
}
void set_gemm_store_tma(int mode) { g_gemm_store_tma = mode < 0 ? -1 : mode; }
void set_gemm_cluster_mode(int on) { g_gemm_cluster = on < 0 ? -1 : (on ? 1 : 0); }
// EVW_GEMM_GEGLU_WARPS=16 selects the 16-warp GEGLU epilogue.  Measured slower than 8 warps (0.554 vs 0.514 ms at
// 258048 x 2560 x 320, profiles/r01h_gemm_geglu_warps.log): the epilogue is bound by instruction count, not by latency.
static int gemm_geglu_wide_epilogue() {
  static const int wide = [] { const char* e = getenv("EVW_GEMM_GEGLU_WARPS"); return (e && atoi(e) == 16) ? 1 : 0; }();
  return wide;
}

int gemm_launch(const GemmOp& op, cudaStream_t stream) {
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                           const CUtensorMap, const KernelParams);
  // [pair mode][epilogue]: 0..5 = plain epilogue by operand set {none, rv, r1, rv+r1, r1+r2, rv+r1+r2}, 6 = GEGLU with
  // 8 epilogue warps, 7 = GEGLU with 16, 8..11 = the four residual sets {r1, rv+r1, r1+r2, rv+r1+r2} with res1 staged by TMA
#define EVW_GEMM_ROW(C)                                                                                                          \
  {tc_gemm_kernel<C, kEpiWarpsDefault, 0>, tc_gemm_kernel<C, kEpiWarpsDefault, 1>, tc_gemm_kernel<C, kEpiWarpsDefault, 2>,       \
   tc_gemm_kernel<C, kEpiWarpsDefault, 3>, tc_gemm_kernel<C, kEpiWarpsDefault, 6>, tc_gemm_kernel<C, kEpiWarpsDefault, 7>,       \
   tc_gemm_kernel<C, kEpiWarpsDefault, kEpiGeglu>, tc_gemm_kernel<C, kEpiWarpsGeglu, kEpiGeglu>,                                 \
   tc_gemm_kernel<C, kEpiWarpsDefault, 16 + 2>, tc_gemm_kernel<C, kEpiWarpsDefault, 16 + 3>,                                     \
   tc_gemm_kernel<C, kEpiWarpsDefault, 16 + 6>, tc_gemm_kernel<C, kEpiWarpsDefault, 16 + 7>,                                     \
   tc_gemm_kernel<C, kEpiWarpsGeglu, 0>, tc_gemm_kernel<C, kEpiWarpsGeglu, 1>, /* 12, 13: no-residual epilogues, 16 warps */ \
   tc_gemm_kernel<C, kEpiWarpsDefault, 32 + 0>, tc_gemm_kernel<C, kEpiWarpsDefault, 32 + 1>,                                     \
   tc_gemm_kernel<C, kEpiWarpsDefault, 32 + 16 + 2>, /* 14..16: GroupNorm statistics of the output {none, rv, r1 by TMA} */   \
   /* 17..27: the same epilogues with the TMA-store path (index 17 + i = variant i of {0,1,2,3,6: GEGLU,8,9,10,11,14,15,16}) */   \
   tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 0>, tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 1>,                                     \
   tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 2>, tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 3>,                                     \
   tc_gemm_kernel<C, kEpiWarpsDefault, 64 + kEpiGeglu>,                                                                          \
   tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 16 + 2>, tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 16 + 3>,                           \
   tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 16 + 6>, tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 16 + 7>,                           \
   tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 32 + 0>, tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 32 + 1>,                           \
   tc_gemm_kernel<C, kEpiWarpsDefault, 64 + 32 + 16 + 2>,                                                                        \
   /* 29, 30: GELU of (acc + bias), no other operand: direct stores / TMA stores */                                             \
   tc_gemm_kernel<C, kEpiWarpsDefault, 128 + 0>, tc_gemm_kernel<C, kEpiWarpsDefault, 128 + 64 + 0>}
  static const KernelFn fns[2][31] = {EVW_GEMM_ROW(false), EVW_GEMM_ROW(true)};
#undef EVW_GEMM_ROW
  static bool attr_set = false;
  if (!attr_set) {
    for (int c = 0; c < 2; ++c)
      for (int w = 0; w < 31; ++w) {
        cudaError_t e = cudaFuncSetAttribute(fns[c][w], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
          set_error("cudaFuncSetAttribute(tc_gemm_kernel): %s", cudaGetErrorString(e));
          return EVW_ERR_CUDA;
        }
      }
    attr_set = true;
  }
  KernelParams P;
  memcpy(&P, op.params, sizeof(P));
  const CUtensorMap& ta0 = *reinterpret_cast<const CUtensorMap*>(op.tmap_a0);
  const CUtensorMap& ta1 = *reinterpret_cast<const CUtensorMap*>(op.tmap_a1);
  const CUtensorMap& tb = *reinterpret_cast<const CUtensorMap*>(op.tmap_b);
  const CUtensorMap& tbh = *reinterpret_cast<const CUtensorMap*>(op.tmap_bh);
  // EVW_GEMM_GEGLU_WARPS=16 / EVW_GEMM_PLAIN_WARPS=16: sixteen epilogue warps for the GEGLU / the residual-free epilogues
  static const int plain_wide = [] { const char* e = getenv("EVW_GEMM_PLAIN_WARPS"); return (e && atoi(e) == 16) ? 1 : 0; }();
  int wide = (P.ep.geglu && gemm_geglu_wide_epilogue()) ? 1 : 0;
  if (!P.ep.geglu && !P.ep.res1 && !P.ep.res2 && plain_wide) wide = 1;
  const int threads = kCtrlThreads + 32 * (wide ? kEpiWarpsGeglu : kEpiWarpsDefault);
  int epi;
  if (P.ep.geglu) epi = 6 + wide;
  else if (P.ep.res2) epi = P.ep.rowvec ? 5 : 4;  // res2 is only ever used together with res1
  else epi = (P.ep.rowvec ? 1 : 0) + (P.ep.res1 ? 2 : 0);
  if (op.res_tma) epi = 8 + (epi - 2);  // epi in {2,3,4,5} -> {8,9,10,11}
  if (!P.ep.geglu && wide) epi = 12 + epi;  // epi in {0,1} -> {12,13}
  if (P.ep.gn_stats) {
    epi = epi == 0 ? 14 : (epi == 1 ? 15 : (epi == 8 ? 16 : -1));
    if (epi < 0) {
      set_error("gemm: GroupNorm statistics are not built for this epilogue");
      return EVW_ERR_INVALID;
    }
    // [instances, 32, 2] doubles, instances = rows / rows per instance
    const long long rows = (long long)P.B * P.T * P.Y * P.X;
    cudaError_t em = cudaMemsetAsync(P.ep.gn_stats, 0, sizeof(double) * 64 * (rows / P.gn_rows_per_inst), stream);
    if (em != cudaSuccess) {
      set_error("gemm: clearing the GroupNorm statistics: %s", cudaGetErrorString(em));
      return EVW_ERR_CUDA;
    }
  }
  if (op.store_tma) {  // twin of the selected epilogue with the TMA-store path, where one is built
    static const int twin[17] = {17, 18, 19, 20, -1, -1, 21, -1, 22, 23, 24, 25, -1, -1, 26, 27, 28};
    if (twin[epi] >= 0) epi = twin[epi];
  }
  if (P.ep.act) epi = op.store_tma ? 30 : 29;  // gemm_plan admits it only without row vector / residual / statistics / GEGLU
  if (op.strided_out && epi < 17) {
    set_error("gemm: strided output view without a TMA-store kernel for this epilogue");
    return EVW_ERR_INVALID;
  }
  KernelFn fn = fns[op.cluster ? 1 : 0][epi];
  const CUtensorMap& tr = *reinterpret_cast<const CUtensorMap*>(op.tmap_r);
  const CUtensorMap& to = *reinterpret_cast<const CUtensorMap*>(op.tmap_o);
  cudaError_t e;
  if (op.cluster) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)op.grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)op.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, fn, ta0, ta1, tb, tbh, tr, to, P);
  } else {
    fn<<<op.grid, threads, op.smem_bytes, stream>>>(ta0, ta1, tb, tbh, tr, to, P);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    set_error("tc_gemm_kernel launch: %s", cudaGetErrorString(e));
    return EVW_ERR_CUDA;
  }
  return EVW_OK;
}

}  // namespace evw

// ------------------------------------------------------------------------------------------
// C ABI: one generic entry (used by the parity tests and by Python-side micro-benchmarks)
// ------------------------------------------------------------------------------------------
extern "C" void evw_set_gemm_cluster(int on) { evw::set_gemm_cluster_mode(on); }
extern "C" void evw_set_gemm_gn_stats(int on) { evw::set_gemm_gn_stats(on); }
extern "C" void evw_set_gemm_store_tma(int mode) { evw::set_gemm_store_tma(mode); }

extern "C" int evw_gemm_f16_gn(const void* a0, const void* a1, const void* w, int B, int T, int Y, int X, int C0, int C1,
                               int N, int num_taps, const int8_t* h_taps /*[num_taps,4] dx,dy,dt,src*/, void* out,
                               int out_fp16, const float* bias, const float* rowvec, int64_t rv_div, int64_t rv_mod,
                               const void* res1, int res1_fp16, float s1, const float* res2, float s2, float s0, int geglu,
                               int block_n, void* out_lo, double* gn_stats, int64_t gn_rows_per_inst, void* stream) {
  evw::GemmProblem pr{};
  pr.a0 = a0; pr.a1 = a1; pr.w = w;
  pr.B = B; pr.T = T; pr.Y = Y; pr.X = X; pr.C0 = C0; pr.C1 = C1; pr.N = N;
  pr.num_taps = num_taps;
  EVW_CHECK_ARG(num_taps >= 1 && num_taps <= evw::kMaxTaps && h_taps, "evw_gemm_f16: bad taps");
  long long ktot = 0;
  for (int t = 0; t < num_taps; ++t) {
    pr.tap_dx[t] = h_taps[4 * t]; pr.tap_dy[t] = h_taps[4 * t + 1]; pr.tap_dt[t] = h_taps[4 * t + 2];
    pr.tap_src[t] = h_taps[4 * t + 3];
    ktot += pr.tap_src[t] ? C1 : C0;
  }
  pr.K_total = ktot;
  pr.block_n = block_n;
  pr.ep.out = out; pr.ep.out_fp16 = out_fp16; pr.ep.out_lo = out_lo; pr.ep.bias = bias; pr.ep.rowvec = rowvec;
  pr.ep.rv_div = rv_div > 0 ? rv_div : 1; pr.ep.rv_mod = rv_mod > 0 ? rv_mod : 1;
  pr.ep.res1 = res1; pr.ep.res1_fp16 = res1_fp16; pr.ep.s1 = s1; pr.ep.res2 = res2; pr.ep.s2 = s2; pr.ep.s0 = s0;
  EVW_CHECK_ARG(geglu >= 0 && geglu <= 2, "evw_gemm_f16: geglu must be 0, 1 (GEGLU) or 2 (GELU of acc + bias)");
  pr.ep.geglu = geglu == 1 ? 1 : 0;
  pr.ep.act = geglu == 2 ? 1 : 0;
  EVW_CHECK_ARG(!(pr.ep.act && gn_stats), "evw_gemm_f16_gn: no GroupNorm statistics with the GELU epilogue");
  evw::GemmOp op;
  int rc = evw::gemm_plan(&op, pr);
  if (rc) return rc;
  if (gn_stats) {
    EVW_CHECK_ARG(evw::gemm_enable_gn_stats(&op, gn_stats, gn_rows_per_inst) == 0,
                  "evw_gemm_f16_gn: this GEMM cannot accumulate GroupNorm statistics (epilogue / geometry)");
  }
  return evw::gemm_launch(op, (cudaStream_t)stream);
}

// nearest-x2 up-sampling + 3x3 convolution (diffusers Upsample2D): a fp16 [F, h, w, C] -> out fp32 [F, 2h, 2w, N];
// w4 = the four phase weight matrices [4][N, 4C] (evoworld_b200/ops.py::upconv_weights), bias fp32 [N] or null
extern "C" int evw_upconv2x_f16(const void* a, const void* w4, const float* bias, void* out, int F, int h, int w, int C, int N,
                                void* stream) {
  EVW_CHECK_ARG(a && w4 && out, "evw_upconv2x_f16: null pointer");
  for (int phase = 0; phase < 4; ++phase) {
    evw::GemmProblem pr{};
    pr.a0 = a; pr.w = reinterpret_cast<const __half*>(w4) + (size_t)phase * N * 4 * C;
    pr.B = 1; pr.T = F; pr.Y = h; pr.X = w; pr.C0 = C; pr.N = N;
    evw::upconv2x_phase(pr, phase);
    pr.ep.out = out; pr.ep.out_fp16 = 0; pr.ep.bias = bias;
    evw::GemmOp op;
    int rc = evw::gemm_plan(&op, pr);
    if (rc) return rc;
    rc = evw::gemm_launch(op, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return EVW_OK;
}

extern "C" int evw_gemm_f16(const void* a0, const void* a1, const void* w, int B, int T, int Y, int X, int C0, int C1,
                            int N, int num_taps, const int8_t* h_taps, void* out, int out_fp16, const float* bias,
                            const float* rowvec, int64_t rv_div, int64_t rv_mod, const void* res1, int res1_fp16, float s1,
                            const float* res2, float s2, float s0, int geglu, int block_n, void* out_lo, void* stream) {
  return evw_gemm_f16_gn(a0, a1, w, B, T, Y, X, C0, C1, N, num_taps, h_taps, out, out_fp16, bias, rowvec, rv_div, rv_mod, res1,
                         res1_fp16, s1, res2, s2, s0, geglu, block_n, out_lo, nullptr, 0, stream);
}
