// Small kernels of the VGGT forward pass (reference third_party/vggt/vggt, called at
// unified_loop_consistency.py:114-136).  The linears and convolutions run on the tcgen05 implicit GEMM (tc_gemm.cu), the
// frame / global attention (head width 64) on the tcgen05 flash kernel (tc_attention.cu), the LayerNorms on
// layer_norm_kernel (unet_elem.cu), the camera trunk's attention (head width 128 over <= 1024 frames) on
// small_attention_kernel (clip_elem.cu); evoworld_b200/vggt.py strings them together.  What is left is HBM-bound
// elementwise work:
//   qknorm_rope_kernel   layers/attention.py:54-58 — LayerNorm(64) of every q / k head followed by the 2-D rotary embedding
//                        (layers/rope.py:116-188), in place on the fused qkv activation
//   bilinear_ac_kernel   heads/dpt_head.py:463-484 custom_interpolate = F.interpolate(bilinear, align_corners=True), channels
//                        last, optional per-pixel addend (the uv positional embedding of dpt_head.py:258-259), fp16 or fp32 out
//   adaln_modulate_kernel heads/camera_head.py:118-122  gate * (LN(x) * (1 + scale) + shift) + x
//   dpt_activate_kernel  heads/head_act.py:62-112 (exp / inv_log points, 1 + exp confidence)
#include "common.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace evw {
namespace {

// Eight threads per (token row, q|k, head): thread `sub` holds features 8 sub .. 8 sub + 7 of the 64 (one 16-byte load; a
// warp covers 4 consecutive heads = 512 contiguous bytes).  LayerNorm over the 64 features (3 shuffle stages inside the
// 8-lane group), then per 32-feature half (sub < 4: vertical, position y; else horizontal, position x):
// out[j] = t[j] cos[p][j % 16] + rot[j] sin[p][j % 16] with rot = (-t[16:], t[:16]) — the partner feature j +- 16 sits in
// the thread two lanes away (sub ^ 2), same slot.  cos / sin: fp32 [max_pos, 16].
__global__ void __launch_bounds__(256)
qknorm_rope_kernel(__half* __restrict__ qkv, long long items, int heads, int tokens_per_frame, const int* __restrict__ pos_yx,
                   const float* __restrict__ q_gamma, const float* __restrict__ q_beta, const float* __restrict__ k_gamma,
                   const float* __restrict__ k_beta, const float* __restrict__ cos_t, const float* __restrict__ sin_t, float eps) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = (t >> 3) < items;
  const long long item = live ? (t >> 3) : items - 1;  // tail lanes shadow the last item (they take part in the shuffles)
  const int sub = (int)(t & 7);
  const int per_row = 2 * heads;
  const long long row = item / per_row;
  const int rem = (int)(item - row * per_row);
  const int which = rem >= heads ? 1 : 0;
  const int C = heads * 64;
  __half* p = qkv + row * (3ll * C) + (long long)rem * 64 + sub * 8;  // q heads then k heads: rem * 64 = which * C + head * 64
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  float x[8];
  {
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h2[i]);
      x[2 * i] = f.x; x[2 * i + 1] = f.y;
    }
  }
  float s = ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
#pragma unroll
  for (int o = 4; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 64.0f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] -= mean;
    q = fmaf(x[i], x[i], q);
  }
#pragma unroll
  for (int o = 4; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / 64.0f) + eps);
  const float4* g4 = reinterpret_cast<const float4*>((which ? k_gamma : q_gamma) + sub * 8);
  const float4* b4 = reinterpret_cast<const float4*>((which ? k_beta : q_beta) + sub * 8);
  const int tok = (int)(row % tokens_per_frame);
  const int pos = pos_yx[2 * tok + (sub < 4 ? 0 : 1)];
  const float4* c4 = reinterpret_cast<const float4*>(cos_t + pos * 16 + (sub & 1) * 8);
  const float4* s4 = reinterpret_cast<const float4*>(sin_t + pos * 16 + (sub & 1) * 8);
  const float sgn = (sub & 2) ? 1.0f : -1.0f;
  float y[8];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const float4 g = __ldg(g4 + hf), b = __ldg(b4 + hf);
    y[4 * hf] = fmaf(x[4 * hf] * rstd, g.x, b.x);
    y[4 * hf + 1] = fmaf(x[4 * hf + 1] * rstd, g.y, b.y);
    y[4 * hf + 2] = fmaf(x[4 * hf + 2] * rstd, g.z, b.z);
    y[4 * hf + 3] = fmaf(x[4 * hf + 3] * rstd, g.w, b.w);
  }
  float o[8];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const float4 c = __ldg(c4 + hf), sn = __ldg(s4 + hf);
    const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float partner = __shfl_xor_sync(0xffffffffu, y[4 * hf + i], 2);
      o[4 * hf + i] = fmaf(y[4 * hf + i], cc[i], sgn * partner * ss[i]);
    }
  }
  if (live) {
    uint4 w;
    __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]);
    __half2 h2 = __floats2half2_rn(o[4], o[5]), h3 = __floats2half2_rn(o[6], o[7]);
    w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
    w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(p) = w;
  }
}

// src fp32 [F, h, w, C] -> dst [F, H, W, C] (fp16 or fp32), 4 channels per thread; index arithmetic as ATen's
// upsample_bilinear2d with align_corners: scale = (in - 1) / (out - 1), i0 = (int)(scale * o), lambda1 = scale * o - i0.
template <bool OUT_HALF>
__global__ void __launch_bounds__(256)
bilinear_ac_kernel(const float* __restrict__ src, void* __restrict__ dst, const float* __restrict__ addend, int F, int h, int w,
                   int H, int W, int C4, float sy, float sx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)F * H * W * C4;
  if (i >= total) return;
  const int c = (int)(i % C4);
  long long t = i / C4;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H);
  const int f = (int)(t / H);
  const float fy = sy * y, fx = sx * x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = fy - y0, lx1 = fx - x0, ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
  const float4* s4 = reinterpret_cast<const float4*>(src) + (long long)f * h * w * C4;
  const float4 v00 = __ldg(s4 + ((long long)y0 * w + x0) * C4 + c), v01 = __ldg(s4 + ((long long)y0 * w + x1) * C4 + c);
  const float4 v10 = __ldg(s4 + ((long long)y1 * w + x0) * C4 + c), v11 = __ldg(s4 + ((long long)y1 * w + x1) * C4 + c);
  float4 r;
  r.x = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
  r.y = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
  r.z = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
  r.w = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
  if (addend) {
    const float4 ad = __ldg(reinterpret_cast<const float4*>(addend) + ((long long)y * W + x) * C4 + c);
    r.x += ad.x; r.y += ad.y; r.z += ad.z; r.w += ad.w;
  }
  if (OUT_HALF) {
    __half2 h0 = __floats2half2_rn(r.x, r.y), h1 = __floats2half2_rn(r.z, r.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    reinterpret_cast<uint2*>(dst)[i] = u;
  } else {
    reinterpret_cast<float4*>(dst)[i] = r;
  }
}

// x <- relu(x) in place (fp32) and its fp16 copy: ResidualConvUnit's nn.ReLU(inplace=True) (heads/dpt_head.py:333,397) — the
// convolution reads the fp16 copy, the skip connection (:410) the overwritten fp32 tensor.
__global__ void relu_inplace_kernel(float* __restrict__ x, __half* __restrict__ out, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<float4*>(x)[i];
  v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  reinterpret_cast<float4*>(x)[i] = v;
  __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
  reinterpret_cast<uint2*>(out)[i] = u;
}

// out = gate * (xn * (1 + scale) + shift) + x;  mod fp32 [rows, 3 C] = (shift | scale | gate)
__global__ void adaln_modulate_kernel(const float* __restrict__ xn, const float* __restrict__ mod, const float* __restrict__ x,
                                      float* __restrict__ out, long long rows, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  const float* m = mod + r * 3ll * C;
  out[i] = m[2 * C + c] * (xn[i] * (1.0f + m[C + c]) + m[c]) + x[i];
}

// x fp32 [rows, ld]: channels 0 .. n_ch-2 -> points (mode 0: exp, 1: sign(v) expm1(|v|)), channel n_ch-1 -> 1 + exp
__global__ void dpt_activate_kernel(const float* __restrict__ x, long long rows, int ld, int n_ch, int mode, float* __restrict__ pts,
                                    float* __restrict__ conf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float* r = x + i * ld;
  for (int c = 0; c < n_ch - 1; ++c) {
    const float v = r[c];
    pts[i * (n_ch - 1) + c] = mode == 0 ? expf(v) : copysignf(expm1f(fabsf(v)), v);
  }
  conf[i] = 1.0f + expf(r[n_ch - 1]);
}

}  // namespace
}  // namespace evw

extern "C" int evw_qknorm_rope_f16(void* qkv, int64_t rows, int heads, int tokens_per_frame, const int* pos_yx, const float* q_gamma,
                                   const float* q_beta, const float* k_gamma, const float* k_beta, const float* cos_t,
                                   const float* sin_t, float eps, void* stream) {
  EVW_CHECK_ARG(qkv && pos_yx && q_gamma && q_beta && k_gamma && k_beta && cos_t && sin_t, "evw_qknorm_rope_f16: null pointer");
  EVW_CHECK_ARG(rows >= 0 && heads >= 1 && tokens_per_frame >= 1, "evw_qknorm_rope_f16: bad extents");
  if (rows == 0) return EVW_OK;
  const long long items = rows * heads * 2;  // (row, q|k, head): 8 threads each
  const long long blocks = (items * 8 + 255) / 256;
  EVW_CHECK_ARG(blocks < (1ll << 31), "evw_qknorm_rope_f16: too many rows");
  EVW_CHECK_ARG(((uintptr_t)qkv & 15) == 0 && ((uintptr_t)q_gamma & 15) == 0 && ((uintptr_t)q_beta & 15) == 0 && ((uintptr_t)k_gamma & 15) == 0 &&
                    ((uintptr_t)k_beta & 15) == 0 && ((uintptr_t)cos_t & 15) == 0 && ((uintptr_t)sin_t & 15) == 0,
                "evw_qknorm_rope_f16: pointers must be 16-byte aligned");
  evw::qknorm_rope_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((__half*)qkv, items, heads, tokens_per_frame, pos_yx,
                                                                               q_gamma, q_beta, k_gamma, k_beta, cos_t, sin_t, eps);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_bilinear_ac_f32(const float* src, void* dst, int out_fp16, const float* addend, int F, int h, int w, int H, int W,
                                   int C, void* stream) {
  EVW_CHECK_ARG(src && dst && F >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1 && C >= 4 && C % 4 == 0,
                "evw_bilinear_ac_f32: bad arguments (C must be a multiple of 4)");
  const long long total = (long long)F * H * W * (C / 4);
  const long long blocks = (total + 255) / 256;
  EVW_CHECK_ARG(blocks < (1ll << 31), "evw_bilinear_ac_f32: output too large");
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f;
  const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
  if (out_fp16)
    evw::bilinear_ac_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, addend, F, h, w, H, W, C / 4, sy, sx);
  else
    evw::bilinear_ac_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, addend, F, h, w, H, W, C / 4, sy, sx);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_relu_inplace_f16(float* x, void* out, int64_t n, void* stream) {
  EVW_CHECK_ARG(x && out && n >= 0 && n % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 7) == 0,
                "evw_relu_inplace_f16: n must be a multiple of 4 and the pointers 16- / 8-byte aligned");
  if (n == 0) return EVW_OK;
  evw::relu_inplace_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (__half*)out, n / 4);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_adaln_modulate_f32(const float* xn, const float* mod, const float* x, float* out, int64_t rows, int C, void* stream) {
  EVW_CHECK_ARG(xn && mod && x && out && rows >= 0 && C >= 1, "evw_adaln_modulate_f32: bad arguments");
  if (rows == 0) return EVW_OK;
  const long long n = rows * C;
  evw::adaln_modulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xn, mod, x, out, rows, C);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_dpt_activate_f32(const float* x, int64_t rows, int ld, int n_ch, int mode, float* pts, float* conf, void* stream) {
  EVW_CHECK_ARG(x && pts && conf && rows >= 0 && n_ch >= 2 && n_ch <= ld && (mode == 0 || mode == 1), "evw_dpt_activate_f32: bad arguments");
  if (rows == 0) return EVW_OK;
  evw::dpt_activate_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, rows, ld, n_ch, mode, pts, conf);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
