// Memory-bound kernels of the UNet denoise step (hot path 1).  Activations are channels-last
// ([rows, C], rows = (b, t, y, x) flattened); the residual stream is fp32, GEMM operands fp16.
//
//   group_norm (stats + apply[+SiLU])   nn.GroupNorm(32) in ResnetBlock2D / TemporalResnetBlock /
//                                       TransformerSpatioTemporalModel.norm / conv_norm_out  (diffusers resnet.py)
//   layer_norm                          BasicTransformerBlock / TemporalBasicTransformerBlock norms (attention.py)
//   temporal_attention                  TemporalBasicTransformerBlock.attn1 over T frames (K9 in SURVEY §2.3)
//   upsample2x / downsplit              Upsample2D nearest x2, Downsample2D stride-2 phase split
//   pre / post                          pipeline_evoworld.py:691-695 (scale + concat) and :709-714 (CFG + Euler)
//   timestep embedding                  diffusers embeddings.py Timesteps(flip_sin_to_cos=True, shift 0)
#include "common.h"
#include "unet_elem.h"
#include <cuda_fp16.h>
#include <cstdlib>

namespace evw {
namespace {

__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// ------------------------------------------------------------------------------------------
// GroupNorm: (1) per-channel partial sums in registers (each thread owns fixed channel quads, rows are
// strided over thread rows -> coalesced 16-byte loads), reduced per group in shared memory and added
// to double accumulators — or the sums arrive from the epilogue of the GEMM that produced the tensor (tc_gemm.cu);
// (2) apply: (mean, rstd) per (instance, group) from the sums, y = a x + b [+SiLU] -> fp16.
// ------------------------------------------------------------------------------------------
constexpr int kGnMaxNQ = 2;  // channel quads per thread: C <= 4 * 512 * kGnMaxNQ

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __half* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  float2 a = __half22float2(*reinterpret_cast<__half2*>(&r.x)), b = __half22float2(*reinterpret_cast<__half2*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename T0, int NQ, int U>  // NQ channel quads per thread, U rows in flight per thread
__global__ void __launch_bounds__(512, 2)
gn_stats_kernel(const T0* __restrict__ src0, int C0, const float* __restrict__ src1, int C1, long long rows_per_inst,
                int rows_per_block, int groups, int TQ, double* __restrict__ stats) {
  const int inst = blockIdx.y;
  const int C = C0 + C1;
  const int cg = C / groups;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows_per_inst, r0 + rows_per_block);
  __shared__ double s_sum[64], s_sq[64];
  for (int i = threadIdx.x; i < groups; i += blockDim.x) { s_sum[i] = 0; s_sq[i] = 0; }
  __syncthreads();
  const int tq = threadIdx.x % TQ, tr = threadIdx.x / TQ, R = blockDim.x / TQ;
  float s[NQ][4], q[NQ][4];
#pragma unroll
  for (int j = 0; j < NQ; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[j][i] = 0.f; q[j][i] = 0.f; }
  // U rows in flight per thread (independent 16-byte loads) to cover DRAM latency
  for (long long rb = r0 + tr; rb < r1; rb += (long long)U * R) {
    float4 v[U][NQ];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * R;
      const long long row = (long long)inst * rows_per_inst + r;
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        v[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < r1) {
          const int c = 4 * (tq + j * TQ);
          v[u][j] = (c < C0) ? ld4(src0 + row * C0 + c) : ld4(src1 + row * C1 + (c - C0));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        s[j][0] += v[u][j].x; s[j][1] += v[u][j].y; s[j][2] += v[u][j].z; s[j][3] += v[u][j].w;
        q[j][0] = fmaf(v[u][j].x, v[u][j].x, q[j][0]); q[j][1] = fmaf(v[u][j].y, v[u][j].y, q[j][1]);
        q[j][2] = fmaf(v[u][j].z, v[u][j].z, q[j][2]); q[j][3] = fmaf(v[u][j].w, v[u][j].w, q[j][3]);
      }
  }
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const int c = 4 * (tq + j * TQ);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(&s_sum[(c + i) / cg], (double)s[j][i]);
      atomicAdd(&s_sq[(c + i) / cg], (double)q[j][i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < groups; i += blockDim.x) {
    atomicAdd(&stats[((long long)inst * groups + i) * 2 + 0], s_sum[i]);
    atomicAdd(&stats[((long long)inst * groups + i) * 2 + 1], s_sq[i]);
  }
}

__device__ __forceinline__ float silu_fast(float y) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return y * r;
}

// apply: each thread owns one channel octet of one instance — its 8 (scale, shift) pairs stay in registers — and walks
// rows with U loads in flight; y = a x + b [-> SiLU] -> fp16 with a = rstd * gamma, b = beta - mean * a computed by the
// thread itself from the (instance, group) sums (a separate finalize kernel and its [instances, C] table until round 2).
// (v1 re-read the 64-byte table slice per 32 input bytes and was L1-bound at 88 %.)
template <typename T0, int U>
// (320 threads x 94 registers: two blocks per SM.  Asking for three — 64 registers, spills — gained 4 %; a minimum of ONE made
// ptxas take 100 registers, which rounds to 104 per thread and leaves room for a single block: +3 ms per step.)
__global__ void __launch_bounds__(320, 2)
gn_apply_kernel(const T0* __restrict__ src0, int C0, const float* __restrict__ src1, int C1, long long rows_per_inst,
                int rows_per_block, const double* __restrict__ stats, int groups, double cnt, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu, __half* __restrict__ out,
                __half* __restrict__ raw_out, __half* __restrict__ out_lo) {
  const int C = C0 + C1;
  const int cv = C / 8;
  const int tc = threadIdx.x % cv, tr = threadIdx.x / cv, R = blockDim.x / cv;
  const int c = tc * 8;
  const long long inst = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows_per_inst, r0 + rows_per_block);
  float sa[8], sb[8];
  {
    const int cg = C / groups;
    int g_have = -1;
    float mean = 0.f, rstd = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c + i) / cg;
      if (g != g_have) {  // an octet touches one or two groups (more only for toy widths)
        const double m = stats[(inst * groups + g) * 2] / cnt;
        double var = stats[(inst * groups + g) * 2 + 1] / cnt - m * m;
        if (var < 0) var = 0;
        rstd = (float)(1.0 / sqrt(var + (double)eps));
        mean = (float)m;
        g_have = g;
      }
      sa[i] = rstd * __ldg(gamma + c + i);
      sb[i] = __ldg(beta + c + i) - mean * sa[i];
    }
  }
  const bool from0 = c < C0;
  for (long long rb = r0 + tr; rb < r1; rb += (long long)U * R) {
    float4 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * R;
      if (r < r1) {
        const long long row = inst * rows_per_inst + r;
        if (from0) {
          va[u] = ld4(src0 + row * C0 + c);
          vb[u] = ld4(src0 + row * C0 + c + 4);
        } else {
          va[u] = ld4(src1 + row * C1 + (c - C0));
          vb[u] = ld4(src1 + row * C1 + (c - C0) + 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = rb + (long long)u * R;
      if (r >= r1) break;
      const long long row = inst * rows_per_inst + r;
      const float v[8] = {va[u].x, va[u].y, va[u].z, va[u].w, vb[u].x, vb[u].y, vb[u].z, vb[u].w};
      if (raw_out) {
        uint4 raw;
        __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(raw_out + row * C + c) = raw;
      }
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y = fmaf(v[i], sa[i], sb[i]);
        o[i] = do_silu ? silu_fast(y) : y;
      }
      uint4 raw;
      __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
      *reinterpret_cast<uint4*>(out + row * C + c) = raw;
      if (out_lo) {  // split-precision consumer: fp16 tail = rounding error of the head
        uint4 rl;
        __half2* hl = reinterpret_cast<__half2*>(&rl);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 hd = __half22float2(h[i]);
          hl[i] = __floats2half2_rn(o[2 * i] - hd.x, o[2 * i + 1] - hd.y);
        }
        *reinterpret_cast<uint4*>(out_lo + row * C + c) = rl;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, optional broadcast row vector added first, fp16 output
// ------------------------------------------------------------------------------------------
template <int MAXV, int ROWS>  // MAXV float4 per lane: C <= 128 * MAXV; ROWS rows per warp, loaded together (DRAM latency)
__global__ void layer_norm_kernel(const float* __restrict__ x, const float* __restrict__ rowvec, long long rv_div,
                                  long long rv_mod, long long rows, int C, float eps, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, __half* __restrict__ out, float* __restrict__ out32) {
  // out32 (optional, instead of out): fp32 result — the CLIP encoder's pre-LayerNorm output is its fp32 residual stream
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = C / 4;
  float4 v[ROWS][MAXV];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const long long row = min(row0 + r, rows - 1);
    const float4* xp = reinterpret_cast<const float4*>(x + row * C);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int j = lane + i * 32;
      if (j < nv) v[r][i] = xp[j];
    }
  }
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const long long row = min(row0 + r, rows - 1);
    const float4* rp = rowvec ? reinterpret_cast<const float4*>(rowvec + ((row / rv_div) % rv_mod) * C) : nullptr;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int j = lane + i * 32;
      if (j < nv) {
        if (rp) {
          float4 t = __ldg(rp + j);
          v[r][i].x += t.x; v[r][i].y += t.y; v[r][i].z += t.z; v[r][i].w += t.w;
        }
        s += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
      }
    }
    mean[r] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < ROWS; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    mean[r] = mean[r] / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int j = lane + i * 32;
      if (j < nv) {
        float a = v[r][i].x - mean[r], b = v[r][i].y - mean[r], c = v[r][i].z - mean[r], d = v[r][i].w - mean[r];
        q += (a * a + b * b) + (c * c + d * d);
      }
    }
    rstd[r] = q;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < ROWS; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
  const float4* gp = reinterpret_cast<const float4*>(gamma);
  const float4* bp = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (row0 + r >= rows) break;
    const float rs = rsqrtf(rstd[r] / (float)C + eps), mu = mean[r];
    uint2* op = reinterpret_cast<uint2*>(out + (row0 + r) * C);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int j = lane + i * 32;
      if (j < nv) {
        float4 g = __ldg(gp + j), b = __ldg(bp + j);
        if (out32) {
          reinterpret_cast<float4*>(out32 + (row0 + r) * C)[j] =
              make_float4((v[r][i].x - mu) * rs * g.x + b.x, (v[r][i].y - mu) * rs * g.y + b.y, (v[r][i].z - mu) * rs * g.z + b.z,
                          (v[r][i].w - mu) * rs * g.w + b.w);
          continue;
        }
        __half2 h0 = __floats2half2_rn((v[r][i].x - mu) * rs * g.x + b.x, (v[r][i].y - mu) * rs * g.y + b.y);
        __half2 h1 = __floats2half2_rn((v[r][i].z - mu) * rs * g.z + b.z, (v[r][i].w - mu) * rs * g.w + b.w);
        uint2 o2;
        o2.x = *reinterpret_cast<uint32_t*>(&h0);
        o2.y = *reinterpret_cast<uint32_t*>(&h1);
        op[j] = o2;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Temporal self-attention over T <= 32 frames, head dim 64.  qkv fp16 [B*T*S, 3C] rows (b,t,s),
// columns [q | k | v], each [heads, 64].  One (b, s, head) problem per half-warp (T <= 16) or warp.
// Lane i owns query frame i: q in registers, K/V staged in shared memory and read as broadcasts.
// ------------------------------------------------------------------------------------------
constexpr int kTaWarps = 4;

template <int TP>  // lanes per problem: 16 or 32
__global__ void __launch_bounds__(kTaWarps * 32)
temporal_attn_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int Bn, int T, long long S, int heads,
                     float scale_log2e) {
  constexpr int PPW = 32 / TP;  // problems per warp
  __shared__ __align__(16) __half s_k[kTaWarps][PPW][TP][64];
  __shared__ __align__(16) __half s_v[kTaWarps][PPW][TP][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / TP, li = lane % TP;
  const long long nprob = (long long)Bn * S * heads;
  const long long prob0 = ((long long)blockIdx.x * kTaWarps + warp) * PPW;
  const int C = heads * 64;
  const long long ld = 3LL * C;
  // cooperative K/V load: for each problem of this warp, T rows x 128 B; 8 lanes per row
  for (int pp = 0; pp < PPW; ++pp) {
    const long long prob = prob0 + pp;
    if (prob >= nprob) break;
    const int h = (int)(prob % heads);
    const long long bs = prob / heads;
    const long long s = bs % S, b = bs / S;
    for (int t = lane >> 3; t < T; t += 4) {
      const long long row = ((long long)b * T + t) * S + s;
      const uint4* kp = reinterpret_cast<const uint4*>(qkv + row * ld + C + h * 64) + (lane & 7);
      const uint4* vp = reinterpret_cast<const uint4*>(qkv + row * ld + 2 * C + h * 64) + (lane & 7);
      reinterpret_cast<uint4*>(&s_k[warp][pp][t][0])[lane & 7] = __ldg(kp);
      reinterpret_cast<uint4*>(&s_v[warp][pp][t][0])[lane & 7] = __ldg(vp);
    }
  }
  __syncwarp();
  const long long prob = prob0 + sub;
  if (prob >= nprob || li >= T) return;
  const int h = (int)(prob % heads);
  const long long bs = prob / heads;
  const long long s = bs % S, b = bs / S;
  const long long qrow = ((long long)b * T + li) * S + s;
  __half2 q[32];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(qkv + qrow * ld + h * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 r = __ldg(qp + i);
      q[4 * i + 0] = *reinterpret_cast<__half2*>(&r.x);
      q[4 * i + 1] = *reinterpret_cast<__half2*>(&r.y);
      q[4 * i + 2] = *reinterpret_cast<__half2*>(&r.z);
      q[4 * i + 3] = *reinterpret_cast<__half2*>(&r.w);
    }
  }
  float sc[TP];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < TP; ++j) {
    sc[j] = -INFINITY;
    if (j < T) {
      const __half2* kj = reinterpret_cast<const __half2*>(&s_k[warp][sub][j][0]);
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        float2 a = __half22float2(q[d]), kk = __half22float2(kj[d]);
        acc = fmaf(a.x, kk.x, acc);
        acc = fmaf(a.y, kk.y, acc);
      }
      sc[j] = acc * scale_log2e;
      mx = fmaxf(mx, sc[j]);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < TP; ++j) {
    sc[j] = (j < T) ? exp2f(sc[j] - mx) : 0.f;
    sum += sc[j];
  }
  const float inv = 1.0f / sum;
  float o[64];
#pragma unroll
  for (int d = 0; d < 64; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < TP; ++j) {
    if (j < T) {
      const __half2* vj = reinterpret_cast<const __half2*>(&s_v[warp][sub][j][0]);
      const float p = sc[j] * inv;
#pragma unroll
      for (int d = 0; d < 32; ++d) {
        float2 vv = __half22float2(vj[d]);
        o[2 * d] = fmaf(p, vv.x, o[2 * d]);
        o[2 * d + 1] = fmaf(p, vv.y, o[2 * d + 1]);
      }
    }
  }
  uint4* op = reinterpret_cast<uint4*>(out + qrow * C + h * 64);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 r;
    __half2 h0 = __floats2half2_rn(o[8 * i + 0], o[8 * i + 1]), h1 = __floats2half2_rn(o[8 * i + 2], o[8 * i + 3]);
    __half2 h2 = __floats2half2_rn(o[8 * i + 4], o[8 * i + 5]), h3 = __floats2half2_rn(o[8 * i + 6], o[8 * i + 7]);
    r.x = *reinterpret_cast<uint32_t*>(&h0);
    r.y = *reinterpret_cast<uint32_t*>(&h1);
    r.z = *reinterpret_cast<uint32_t*>(&h2);
    r.w = *reinterpret_cast<uint32_t*>(&h3);
    op[i] = r;
  }
}

// ------------------------------------------------------------------------------------------
// nearest x2 upsample (fp32 [n,h,w,C] -> fp16 [n,2h,2w,C]) and stride-2 phase split
// (fp32 [n,h,w,C] -> fp16 [n*4, h/2, w/2, C], image index n*4 + (y&1)*2 + (x&1))
// ------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const float* __restrict__ x, __half* __restrict__ out, long long n, int h, int w, int C) {
  const int cv = C / 8;
  const long long total = n * (2LL * h) * (2LL * w) * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % cv) * 8;
  long long p = idx / cv;
  const int ox = (int)(p % (2 * w)); p /= 2 * w;
  const int oy = (int)(p % (2 * h));
  const long long img = p / (2 * h);
  const float* src = x + ((img * h + oy / 2) * w + ox / 2) * C + c;
  float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  uint4 r;
  __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2 = __floats2half2_rn(b.x, b.y),
          h3 = __floats2half2_rn(b.z, b.w);
  r.x = *reinterpret_cast<uint32_t*>(&h0); r.y = *reinterpret_cast<uint32_t*>(&h1);
  r.z = *reinterpret_cast<uint32_t*>(&h2); r.w = *reinterpret_cast<uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(out + ((img * 2 * h + oy) * 2 * w + ox) * C + c) = r;
}

__global__ void downsplit_kernel(const float* __restrict__ x, __half* __restrict__ out, long long n, int h, int w, int C) {
  const int cv = C / 8;
  const int h2 = (h + 1) / 2, w2 = (w + 1) / 2;
  const long long total = n * 4 * (long long)h2 * w2 * cv;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % cv) * 8;
  long long p = idx / cv;
  const int x2 = (int)(p % w2); p /= w2;
  const int y2 = (int)(p % h2); p /= h2;
  const int ph = (int)(p % 4);
  const long long img = p / 4;
  const int iy = 2 * y2 + (ph >> 1), ix = 2 * x2 + (ph & 1);
  uint4 r = make_uint4(0, 0, 0, 0);
  if (iy < h && ix < w) {
    const float* src = x + ((img * h + iy) * w + ix) * C + c;
    float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2_ = __floats2half2_rn(b.x, b.y),
            h3 = __floats2half2_rn(b.z, b.w);
    r.x = *reinterpret_cast<uint32_t*>(&h0); r.y = *reinterpret_cast<uint32_t*>(&h1);
    r.z = *reinterpret_cast<uint32_t*>(&h2_); r.w = *reinterpret_cast<uint32_t*>(&h3);
  }
  *reinterpret_cast<uint4*>(out + idx * 8) = r;
}

// ------------------------------------------------------------------------------------------
// pre: latents [1,T,4,h,w] fp32 (NCHW per frame) / sqrt(sigma^2+1), duplicated for CFG, concatenated
// with cond [2,T,Cc,h,w] -> fp16 channels-last [2*T, h, w, Cpad] (zero padded to Cpad)
// ------------------------------------------------------------------------------------------
// split != 0 (needs 3 Cin <= Cpad): the zero padding of conv_in's 64-channel operand carries the split-precision form of
// the input for free — channels [0,Cin) = fp16 head, [Cin,2Cin) = fp16 tail (x - head), [2Cin,3Cin) = head again, against
// conv_in weights packed as [W_hi | W_hi | W_lo] (evoworld_b200/unet.py::pack_parameters).
__device__ __forceinline__ void put_split(__half* o, int c, int Cin, int split, float v) {
  const __half hd = __float2half_rn(v);
  o[c] = hd;
  if (split) {
    o[Cin + c] = __float2half_rn(v - __half2float(hd));
    o[2 * Cin + c] = hd;
  }
}
// The per-step scalars (sigma, guidance, ...) live in a small DEVICE array written by set_step_args_kernel right before
// the plan runs, so that the captured CUDA graph of the plan does not bake them in:
//   dargs[0] = 1/sqrt(sigma^2+1), [1] = sigma, [2] = sigma_next, [3] = g_min, [4] = g_max
__global__ void set_step_args_kernel(float* __restrict__ tsteps, int B, float* __restrict__ dargs, float timestep, float sigma,
                                     float sigma_next, float g_min, float g_max) {
  const int i = threadIdx.x;
  if (i < B) tsteps[i] = timestep;
  if (i == 0) {
    dargs[0] = 1.0f / sqrtf(sigma * sigma + 1.0f);
    dargs[1] = sigma; dargs[2] = sigma_next; dargs[3] = g_min; dargs[4] = g_max;
  }
}

__global__ void pre_kernel(const float* __restrict__ latents, const float* __restrict__ cond, int Bc, int T, int Cl,
                           int Cc, long long HW, const float* __restrict__ dargs, int Cpad, int split, __half* __restrict__ out) {
  const float inv_scale = dargs[0];
  const long long total = (long long)Bc * T * HW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long p = idx % HW;
  const long long f = idx / HW;  // frame in the CFG batch: b*T + t
  const int t = (int)(f % T);
  const int Cin = Cl + Cc;
  __half* o = out + idx * Cpad;
  for (int c = 0; c < Cl; ++c) put_split(o, c, Cin, split, latents[((long long)t * Cl + c) * HW + p] * inv_scale);
  for (int c = 0; c < Cc; ++c) put_split(o, Cl + c, Cin, split, cond[(f * Cc + c) * HW + p]);
  for (int c = (split ? 3 : 1) * Cin; c < Cpad; ++c) o[c] = __float2half_rn(0.f);
}

// raw sample [B,T,Cin,h,w] fp32 -> fp16 channels-last padded (UNet.forward called on its own)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, long long frames, int Cin, long long HW, int Cpad,
                                    int split, __half* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= frames * HW) return;
  const long long p = idx % HW, f = idx / HW;
  __half* o = out + idx * Cpad;
  for (int c = 0; c < Cin; ++c) put_split(o, c, Cin, split, x[(f * Cin + c) * HW + p]);
  for (int c = (split ? 3 : 1) * Cin; c < Cpad; ++c) o[c] = __float2half_rn(0.f);
}

// model output fp32 channels-last [2*T, h, w, Npad] -> NCHW [B,T,Co,h,w]
// fold != 0: conv_out ran in split precision with the weight tail packed into the padded output columns
// (rows [Co, 2Co) of the weight matrix = W_lo), so the result is y[c] + y[Co + c]
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ y, long long frames, int Co, long long HW, int Npad, int fold,
                                    float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= frames * HW) return;
  const long long p = idx % HW, f = idx / HW;
  for (int c = 0; c < Co; ++c) out[(f * Co + c) * HW + p] = y[idx * Npad + c] + (fold ? y[idx * Npad + Co + c] : 0.f);
}

// post: CFG combine + Euler (v-prediction) step on NCHW latents, reading the channels-last model output
//   v = vu + g_t (vc - vu);  x0 = v * (-sigma / sqrt(sigma^2+1)) + x / (sigma^2+1);  d = (x - x0)/sigma;
//   x' = x + d (sigma_next - sigma)
__global__ void post_kernel(const float* __restrict__ y, int T, int Cl, long long HW, int Npad, int fold,
                            const float* __restrict__ dargs, float* __restrict__ latents) {
  const float sigma = dargs[1], sigma_next = dargs[2], g_min = dargs[3], g_max = dargs[4];
  const long long total = (long long)T * HW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long p = idx % HW;
  const int t = (int)(idx / HW);
  const float g = (T > 1) ? g_min + (g_max - g_min) * ((float)t / (float)(T - 1)) : g_min;
  const float c_out = -sigma / sqrtf(sigma * sigma + 1.0f);
  const float c_skip = 1.0f / (sigma * sigma + 1.0f);
  const float* yu = y + idx * Npad;
  const float* yc = y + ((long long)T * HW + idx) * Npad;
  for (int c = 0; c < Cl; ++c) {
    const float vu = yu[c] + (fold ? yu[Cl + c] : 0.f), vc = yc[c] + (fold ? yc[Cl + c] : 0.f);
    const float v = vu + g * (vc - vu);
    float* lp = latents + ((long long)t * Cl + c) * HW + p;
    const float x = *lp;
    const float x0 = v * c_out + x * c_skip;
    const float d = (x - x0) / sigma;
    *lp = x + d * (sigma_next - sigma);
  }
}

// sinusoidal timestep embedding (flip_sin_to_cos=True, downscale_freq_shift=0): out [n, dim] fp16 = [cos | sin]
__global__ void timestep_embed_kernel(const float* __restrict__ t, int n, int dim, __half* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (idx >= n * half) return;
  const int i = idx / half, k = idx % half;
  const float f = expf(-logf(10000.0f) * (float)k / (float)half);
  const float a = t[i] * f;
  out[(long long)i * dim + k] = __float2half_rn(cosf(a));
  out[(long long)i * dim + half + k] = __float2half_rn(sinf(a));
}

// y = silu(x) (fp32 -> fp16), used on the time embedding before every time_emb_proj
__global__ void silu_f16_kernel(const float* __restrict__ x, __half* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(silu(x[i]));
}
__global__ void cast_f16_kernel(const float* __restrict__ x, __half* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(x[i]);
}
__global__ void fill_f32_kernel(float* __restrict__ out, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// ------------------------------------------------------------------------------------------
// Temporal self-attention, tensor-core version: one warp per (b, s, head) problem.  The T x 64 q/k/v tiles are staged
// in shared memory with coalesced 16-byte loads (144-byte row pitch: conflict-free fragment reads), S = Q K^T and
// O = P V run on mma.sync.m16n8k16 (fp16 in, fp32 accumulate) with the score fragments re-used in registers as the
// A operand of P V, the softmax reduces over the 4 lanes that share a row.  T = 14 frames cannot fill a 128-row tcgen05
// tile (8 problems would have to be packed block-diagonally, 8x wasted MMA work), and at ~100 instructions per problem the
// kernel is bound by its 4 x T x 128 bytes of HBM traffic instead of by 3500 scalar instructions per lane (v1: 0.44 ms at
// 2 x 14 x 9216 x 5 heads = 1.5 TB/s).
// ------------------------------------------------------------------------------------------
constexpr int kTmWarps = 4;
constexpr int kTmPitch = 72;  // halves per shared-memory row (64 + 8 pad)

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int MT>  // 16-frame tiles: 1 (T <= 16) or 2 (T <= 32)
__global__ void __launch_bounds__(kTmWarps * 32)
temporal_attn_mma_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int Bn, int T, long long S, int heads,
                         float scale_log2e) {
  constexpr int R = 16 * MT;
  extern __shared__ __align__(16) __half ta_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long nprob = (long long)Bn * S * heads;
  const long long prob = (long long)blockIdx.x * kTmWarps + warp;
  if (prob >= nprob) return;  // warps are independent: no block-level barrier below
  __half* sq = ta_smem + (size_t)warp * 3 * R * kTmPitch;
  __half* sk = sq + R * kTmPitch;
  __half* sv = sk + R * kTmPitch;
  const int h = (int)(prob % heads);
  const long long bs = prob / heads;
  const long long s = bs % S, b = bs / S;
  const int C = heads * 64;
  const long long ld = 3LL * C;
  // stage q, k, v: 8 lanes x 16 bytes per row, 4 rows per trip; rows >= T are zero (P = 0 times garbage must stay 0)
  for (int t = lane >> 3; t < R; t += 4) {
    uint4 q4 = make_uint4(0, 0, 0, 0), k4 = q4, v4 = q4;
    if (t < T) {
      const __half* base = qkv + (((long long)b * T + t) * S + s) * ld + h * 64 + (lane & 7) * 8;
      q4 = __ldg(reinterpret_cast<const uint4*>(base));
      k4 = __ldg(reinterpret_cast<const uint4*>(base + C));
      v4 = __ldg(reinterpret_cast<const uint4*>(base + 2 * C));
    }
    const int o = t * kTmPitch + (lane & 7) * 8;
    *reinterpret_cast<uint4*>(sq + o) = q4;
    *reinterpret_cast<uint4*>(sk + o) = k4;
    *reinterpret_cast<uint4*>(sv + o) = v4;
  }
  __syncwarp();
  const int g = lane >> 2, tig = lane & 3;
  // ---- S = Q K^T
  float sacc[MT][2 * MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2 * MT; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) sacc[mt][nt][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int d0 = 16 * ks + 2 * tig;
    uint32_t a[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const __half* r0 = sq + (16 * mt + g) * kTmPitch + d0;
      const __half* r1 = r0 + 8 * kTmPitch;
      a[mt][0] = *reinterpret_cast<const uint32_t*>(r0);
      a[mt][1] = *reinterpret_cast<const uint32_t*>(r1);
      a[mt][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
      a[mt][3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
    }
#pragma unroll
    for (int nt = 0; nt < 2 * MT; ++nt) {
      const __half* kr = sk + (8 * nt + g) * kTmPitch + d0;
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(kr), b1 = *reinterpret_cast<const uint32_t*>(kr + 8);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_16816(sacc[mt][nt], a[mt], b0, b1);
    }
  }
  // ---- softmax over the key frames of each row (rows 16 mt + g and + 8; a row lives in the 4 lanes of a quad)
  float inv[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 8 * nt + 2 * tig + e;
          float v = sacc[mt][nt][2 * hf + e];
          v = j < T ? v : -INFINITY;
          sacc[mt][nt][2 * hf + e] = v;
          mx = fmaxf(mx, v);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2 * MT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float pv = exp2f((sacc[mt][nt][2 * hf + e] - mx) * scale_log2e);  // exp2(-inf) = 0 for masked frames
          sacc[mt][nt][2 * hf + e] = pv;
          sum += pv;
        }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      inv[mt][hf] = 1.0f / sum;
    }
  // ---- O = P V: the score fragments are the A fragments of the second product
  float oacc[MT][8][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) oacc[mt][nt][i] = 0.f;
#pragma unroll
  for (int kt = 0; kt < MT; ++kt) {
    uint32_t pa[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      __half2 h0 = __floats2half2_rn(sacc[mt][2 * kt][0], sacc[mt][2 * kt][1]);
      __half2 h1 = __floats2half2_rn(sacc[mt][2 * kt][2], sacc[mt][2 * kt][3]);
      __half2 h2 = __floats2half2_rn(sacc[mt][2 * kt + 1][0], sacc[mt][2 * kt + 1][1]);
      __half2 h3 = __floats2half2_rn(sacc[mt][2 * kt + 1][2], sacc[mt][2 * kt + 1][3]);
      pa[mt][0] = *reinterpret_cast<uint32_t*>(&h0);
      pa[mt][1] = *reinterpret_cast<uint32_t*>(&h1);
      pa[mt][2] = *reinterpret_cast<uint32_t*>(&h2);
      pa[mt][3] = *reinterpret_cast<uint32_t*>(&h3);
    }
    const uint32_t vrow = (uint32_t)__cvta_generic_to_shared(sv + (16 * kt + (lane & 15)) * kTmPitch);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      uint32_t b0, b1;
      asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b0), "=r"(b1) : "r"(vrow + nt * 16));
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) mma_16816(oacc[mt][nt], pa[mt], b0, b1);
    }
  }
  // ---- normalise, stage through the (now free) q tile, store 16 bytes per lane
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      __half2 lo = __floats2half2_rn(oacc[mt][nt][0] * inv[mt][0], oacc[mt][nt][1] * inv[mt][0]);
      __half2 hi = __floats2half2_rn(oacc[mt][nt][2] * inv[mt][1], oacc[mt][nt][3] * inv[mt][1]);
      *reinterpret_cast<__half2*>(sq + (16 * mt + g) * kTmPitch + 8 * nt + 2 * tig) = lo;
      *reinterpret_cast<__half2*>(sq + (16 * mt + g + 8) * kTmPitch + 8 * nt + 2 * tig) = hi;
    }
  __syncwarp();
  for (int t = lane >> 3; t < T; t += 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(sq + t * kTmPitch + (lane & 7) * 8);
    *reinterpret_cast<uint4*>(out + (((long long)b * T + t) * S + s) * C + h * 64 + (lane & 7) * 8) = v;
  }
}

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

// ==========================================================================================
// host launchers
// ==========================================================================================
int group_norm(const void* src0, int src0_fp16, int C0, const float* src1, int C1, long long insts,
               long long rows_per_inst, float eps, const float* gamma, const float* beta, int do_silu, double* stats,
               __half* out, __half* raw_out, __half* out_lo, cudaStream_t st, int have_stats) {
  // have_stats: `stats` already holds the sums (accumulated by the epilogue of the GEMM that produced src0)
  const int C = C0 + C1, groups = 32;
  EVW_CHECK_ARG(C % groups == 0 && C % 8 == 0 && C0 % 8 == 0, "group_norm: C=%d (C0=%d) not supported", C, C0);
  const int Q = C / 4;
  int NQ = (Q + 511) / 512;
  while (NQ <= kGnMaxNQ && Q % NQ != 0) ++NQ;
  EVW_CHECK_ARG(NQ <= kGnMaxNQ, "group_norm: C=%d too wide", C);
  const int TQ = Q / NQ;
  const int R = 512 / TQ > 0 ? 512 / TQ : 1;
  const int threads = TQ * R;
  // stats scratch: [insts, 32, 2] doubles
  if (!have_stats) EVW_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * groups * insts, st));
  long long want_blocks = (long long)sm_count() * 4 / (insts > 0 ? insts : 1) + 1;
  constexpr int U1 = 8, U2 = 2;  // rows in flight per thread for NQ = 1 / 2
  const int U = NQ == 1 ? U1 : U2;
  int rpb = (int)((rows_per_inst + want_blocks - 1) / want_blocks);
  if (rpb < U * R) rpb = U * R;
  dim3 grid((unsigned)((rows_per_inst + rpb - 1) / rpb), (unsigned)insts);
#define EVW_GN_STATS(T0, NQv, Uv) \
  gn_stats_kernel<T0, NQv, Uv><<<grid, threads, 0, st>>>((const T0*)src0, C0, src1, C1, rows_per_inst, rpb, groups, TQ, stats)
  if (have_stats) {}
  else if (src0_fp16) { if (NQ == 1) EVW_GN_STATS(__half, 1, U1); else EVW_GN_STATS(__half, 2, U2); }
  else                { if (NQ == 1) EVW_GN_STATS(float, 1, U1); else EVW_GN_STATS(float, 2, U2); }
#undef EVW_GN_STATS
  {
    const double cnt = (double)rows_per_inst * (C / groups);
    const int cv = C / 8;
    EVW_CHECK_ARG(cv <= 320, "group_norm: C=%d too wide for the apply kernel", C);
    const int Ra = 320 / cv > 0 ? 320 / cv : 1;
    const int athreads = cv * Ra;
    constexpr int UA = 4;
    long long awant = (long long)sm_count() * 8 / (insts > 0 ? insts : 1) + 1;
    int arpb = (int)((rows_per_inst + awant - 1) / awant);
    if (arpb < UA * Ra) arpb = UA * Ra;
    dim3 agrid((unsigned)((rows_per_inst + arpb - 1) / arpb), (unsigned)insts);
    if (src0_fp16)
      gn_apply_kernel<__half, UA><<<agrid, athreads, 0, st>>>((const __half*)src0, C0, src1, C1, rows_per_inst, arpb, stats, groups,
                                                               cnt, eps, gamma, beta, do_silu, out, raw_out, out_lo);
    else
      gn_apply_kernel<float, UA><<<agrid, athreads, 0, st>>>((const float*)src0, C0, src1, C1, rows_per_inst, arpb, stats, groups,
                                                              cnt, eps, gamma, beta, do_silu, out, raw_out, out_lo);
  }
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

int layer_norm(const float* x, const float* rowvec, long long rv_div, long long rv_mod, long long rows, int C, float eps,
               const float* gamma, const float* beta, __half* out, cudaStream_t st, float* out32) {
  EVW_CHECK_ARG(C % 4 == 0 && C <= 128 * 20, "layer_norm: C=%d not supported", C);
  const int warps = 8;
  auto grid = [&](int rows_per_warp) { return (unsigned)((rows + (long long)warps * rows_per_warp - 1) / ((long long)warps * rows_per_warp)); };
  if (C <= 128 * 3)
    layer_norm_kernel<3, 4><<<grid(4), warps * 32, 0, st>>>(x, rowvec, rv_div, rv_mod, rows, C, eps, gamma, beta, out, out32);
  else if (C <= 128 * 5)
    layer_norm_kernel<5, 4><<<grid(4), warps * 32, 0, st>>>(x, rowvec, rv_div, rv_mod, rows, C, eps, gamma, beta, out, out32);
  else if (C <= 128 * 10)
    layer_norm_kernel<10, 2><<<grid(2), warps * 32, 0, st>>>(x, rowvec, rv_div, rv_mod, rows, C, eps, gamma, beta, out, out32);
  else
    layer_norm_kernel<20, 1><<<grid(1), warps * 32, 0, st>>>(x, rowvec, rv_div, rv_mod, rows, C, eps, gamma, beta, out, out32);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

int temporal_attention(const __half* qkv, __half* out, int B, int T, long long S, int heads, cudaStream_t st) {
  EVW_CHECK_ARG(T >= 1 && T <= 32, "temporal_attention: T=%d must be in [1,32]", T);
  const float scale_log2e = 0.125f * 1.4426950408889634f;  // head dim 64
  const long long nprob = (long long)B * S * heads;
  static const bool use_v1 = getenv("EVW_TEMPORAL_ATTN_V1") != nullptr;  // scalar kernel, A/B timing only
  if (!use_v1) {
    static bool attr_set = false;
    constexpr int smem1 = kTmWarps * 3 * 16 * kTmPitch * 2, smem2 = kTmWarps * 3 * 32 * kTmPitch * 2;
    if (!attr_set) {
      EVW_CUDA(cudaFuncSetAttribute(temporal_attn_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      attr_set = true;
    }
    const unsigned grid = (unsigned)((nprob + kTmWarps - 1) / kTmWarps);
    if (T <= 16)
      temporal_attn_mma_kernel<1><<<grid, kTmWarps * 32, smem1, st>>>(qkv, out, B, T, S, heads, scale_log2e);
    else
      temporal_attn_mma_kernel<2><<<grid, kTmWarps * 32, smem2, st>>>(qkv, out, B, T, S, heads, scale_log2e);
    EVW_LAUNCH_CHECK();
    return EVW_OK;
  }
  if (T <= 16) {
    const long long per_block = kTaWarps * 2;
    temporal_attn_kernel<16><<<(unsigned)((nprob + per_block - 1) / per_block), kTaWarps * 32, 0, st>>>(
        qkv, out, B, T, S, heads, scale_log2e);
  } else {
    const long long per_block = kTaWarps;
    temporal_attn_kernel<32><<<(unsigned)((nprob + per_block - 1) / per_block), kTaWarps * 32, 0, st>>>(
        qkv, out, B, T, S, heads, scale_log2e);
  }
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

int upsample2x(const float* x, __half* out, long long n, int h, int w, int C, cudaStream_t st) {
  EVW_CHECK_ARG(C % 8 == 0, "upsample2x: C %% 8");
  upsample2x_kernel<<<blocks_for(n * 4LL * h * w * (C / 8), 256), 256, 0, st>>>(x, out, n, h, w, C);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int downsplit(const float* x, __half* out, long long n, int h, int w, int C, cudaStream_t st) {
  EVW_CHECK_ARG(C % 8 == 0, "downsplit: C %% 8");
  downsplit_kernel<<<blocks_for(n * 4LL * ((h + 1) / 2) * ((w + 1) / 2) * (C / 8), 256), 256, 0, st>>>(x, out, n, h, w, C);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int set_step_args(float* tsteps, int B, float* dargs, float timestep, float sigma, float sigma_next, float g_min, float g_max,
                  cudaStream_t st) {
  EVW_CHECK_ARG(B >= 1 && B <= 32, "set_step_args: B=%d", B);
  set_step_args_kernel<<<1, 32, 0, st>>>(tsteps, B, dargs, timestep, sigma, sigma_next, g_min, g_max);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int pre_concat(const float* latents, const float* cond, int Bc, int T, int Cl, int Cc, long long HW, const float* dargs, int Cpad,
               int split, __half* out, cudaStream_t st) {
  EVW_CHECK_ARG(!split || 3 * (Cl + Cc) <= Cpad, "pre_concat: split needs 3 Cin <= Cpad");
  pre_kernel<<<blocks_for((long long)Bc * T * HW, 256), 256, 0, st>>>(latents, cond, Bc, T, Cl, Cc, HW, dargs, Cpad, split, out);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int nchw_to_nhwc_f16(const float* x, long long frames, int Cin, long long HW, int Cpad, int split, __half* out, cudaStream_t st) {
  EVW_CHECK_ARG(!split || 3 * Cin <= Cpad, "nchw_to_nhwc_f16: split needs 3 Cin <= Cpad");
  nchw_to_nhwc_kernel<<<blocks_for(frames * HW, 256), 256, 0, st>>>(x, frames, Cin, HW, Cpad, split, out);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int nhwc_to_nchw_f32(const float* y, long long frames, int Co, long long HW, int Npad, int fold, float* out, cudaStream_t st) {
  EVW_CHECK_ARG(!fold || 2 * Co <= Npad, "nhwc_to_nchw_f32: fold needs 2 Co <= Npad");
  nhwc_to_nchw_kernel<<<blocks_for(frames * HW, 256), 256, 0, st>>>(y, frames, Co, HW, Npad, fold, out);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int post_cfg_euler(const float* y, int T, int Cl, long long HW, int Npad, int fold, const float* dargs, float* latents,
                   cudaStream_t st) {
  EVW_CHECK_ARG(!fold || 2 * Cl <= Npad, "post_cfg_euler: fold needs 2 Cl <= Npad");
  post_kernel<<<blocks_for((long long)T * HW, 256), 256, 0, st>>>(y, T, Cl, HW, Npad, fold, dargs, latents);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int timestep_embed(const float* t, int n, int dim, __half* out, cudaStream_t st) {
  timestep_embed_kernel<<<blocks_for((long long)n * dim / 2, 128), 128, 0, st>>>(t, n, dim, out);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int silu_f16(const float* x, __half* out, long long n, cudaStream_t st) {
  silu_f16_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, out, n);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int cast_f16(const float* x, __half* out, long long n, cudaStream_t st) {
  cast_f16_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, out, n);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int fill_f32(float* out, float v, int n, cudaStream_t st) {
  fill_f32_kernel<<<blocks_for(n, 128), 128, 0, st>>>(out, v, n);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
int add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t st) {
  add_f32_kernel<<<blocks_for(n, 256), 256, 0, st>>>(a, b, out, n);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

}  // namespace evw
