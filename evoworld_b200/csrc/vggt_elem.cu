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
//   patchify_kernel      models/aggregator.py:201 + layers/patch_embed.py:66-77: normalised 14 x 14 patches as GEMM rows
#include "common.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace evw {
namespace {

// Eight threads per (token row, q|k): thread `sub` owns features 8 sub .. 8 sub + 7 of EVERY head of its row (one 16-byte
// load per head; the 8 threads cover a head's 128 contiguous bytes) and keeps the LayerNorm weights and the row's cos / sin
// in registers across the heads — read per head they made the kernel L1-bound (l1tex 81 %, 128 B of table reads per 16 B of
// data: profiles/r02aq_ncu_full_vggt_elem.txt).  LayerNorm over the 64 features (3 shuffle stages inside the 8-lane group),
// then per 32-feature half (sub < 4: vertical, position y; else horizontal, position x):
// out[j] = t[j] cos[p][j % 16] + rot[j] sin[p][j % 16] with rot = (-t[16:], t[:16]) — the partner feature j +- 16 sits in
// the thread two lanes away (sub ^ 2), same slot.  cos / sin: fp32 [max_pos, 16].
__global__ void __launch_bounds__(256)
qknorm_rope_kernel(__half* __restrict__ qkv, long long items /* rows * 2 */, int heads, int tokens_per_frame,
                   const int* __restrict__ pos_yx, const float* __restrict__ q_gamma, const float* __restrict__ q_beta,
                   const float* __restrict__ k_gamma, const float* __restrict__ k_beta, const float* __restrict__ cos_t,
                   const float* __restrict__ sin_t, float eps) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = (t >> 3) < items;
  const long long item = live ? (t >> 3) : items - 1;  // tail lanes shadow the last item (they take part in the shuffles)
  const int sub = (int)(t & 7);
  const long long row = item >> 1;
  const int which = (int)(item & 1);
  const int C = heads * 64;
  __half* p = qkv + row * (3ll * C) + (long long)which * C + sub * 8;
  float g[8], b[8], cs[8], sn[8];
  {
    const float4* g4 = reinterpret_cast<const float4*>((which ? k_gamma : q_gamma) + sub * 8);
    const float4* b4 = reinterpret_cast<const float4*>((which ? k_beta : q_beta) + sub * 8);
    const int tok = (int)(row % tokens_per_frame);
    const int pos = pos_yx[2 * tok + (sub < 4 ? 0 : 1)];
    const float4* c4 = reinterpret_cast<const float4*>(cos_t + pos * 16 + (sub & 1) * 8);
    const float4* s4 = reinterpret_cast<const float4*>(sin_t + pos * 16 + (sub & 1) * 8);
    const float sgn = (sub & 2) ? 1.0f : -1.0f;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float4 gv = __ldg(g4 + hf), bv = __ldg(b4 + hf), cv = __ldg(c4 + hf), sv = __ldg(s4 + hf);
      g[4 * hf] = gv.x; g[4 * hf + 1] = gv.y; g[4 * hf + 2] = gv.z; g[4 * hf + 3] = gv.w;
      b[4 * hf] = bv.x; b[4 * hf + 1] = bv.y; b[4 * hf + 2] = bv.z; b[4 * hf + 3] = bv.w;
      cs[4 * hf] = cv.x; cs[4 * hf + 1] = cv.y; cs[4 * hf + 2] = cv.z; cs[4 * hf + 3] = cv.w;
      sn[4 * hf] = sgn * sv.x; sn[4 * hf + 1] = sgn * sv.y; sn[4 * hf + 2] = sgn * sv.z; sn[4 * hf + 3] = sgn * sv.w;
    }
  }
  uint4 raw = *reinterpret_cast<const uint4*>(p);
  for (int h = 0; h < heads; ++h) {
    uint4 nxt = raw;
    if (h + 1 < heads) nxt = *reinterpret_cast<const uint4*>(p + (h + 1) * 64);  // next head in flight during this one's math
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
    float o8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float y = fmaf(x[i] * rstd, g[i], b[i]);
      const float partner = __shfl_xor_sync(0xffffffffu, y, 2);
      o8[i] = fmaf(y, cs[i], partner * sn[i]);
    }
    if (live) {
      uint4 w;
      __half2 h0 = __floats2half2_rn(o8[0], o8[1]), h1 = __floats2half2_rn(o8[2], o8[3]);
      __half2 h2 = __floats2half2_rn(o8[4], o8[5]), h3 = __floats2half2_rn(o8[6], o8[7]);
      w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
      w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(p + h * 64) = w;
    }
    raw = nxt;
  }
}

// src fp32 [F, h, w, C] -> dst [F, H, W, C] (fp16 or fp32), 4 channels per thread; index arithmetic as ATen's
// upsample_bilinear2d with align_corners: scale = (in - 1) / (out - 1), i0 = (int)(scale * o), lambda1 = scale * o - i0.
template <bool OUT_HALF>
__global__ void __launch_bounds__(256)
bilinear_ac_kernel(const float* __restrict__ src, void* __restrict__ dst, const float* __restrict__ addend, int F, int h, int w,
                   int H, int W, int C4, int c4_shift, float sy, float sx) {
  // grid (ceil(W C4 / 256), H, F): the row / frame come from the block index, one division (a shift when C4 is a power of
  // two) splits the rest into (x, channel quad) — three 64-bit divisions per thread made the first version issue-bound
  // (80 % issue, 5 % DRAM: profiles/r02aq_ncu_full_vggt_elem.txt)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= W * C4) return;
  const int x = c4_shift >= 0 ? (idx >> c4_shift) : idx / C4;
  const int c = idx - x * C4;
  const int y = blockIdx.y, f = blockIdx.z;
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
  const long long pix = (long long)y * W + x;
  if (addend) {
    const float4 ad = __ldg(reinterpret_cast<const float4*>(addend) + pix * C4 + c);
    r.x += ad.x; r.y += ad.y; r.z += ad.z; r.w += ad.w;
  }
  const long long o = ((long long)f * H * W + pix) * C4 + c;
  if (OUT_HALF) {
    __half2 h0 = __floats2half2_rn(r.x, r.y), h1 = __floats2half2_rn(r.z, r.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    reinterpret_cast<uint2*>(dst)[o] = u;
  } else {
    reinterpret_cast<float4*>(dst)[o] = r;
  }
}

// DINOv2 patch embedding operand (layers/patch_embed.py:66-77 + the image normalisation of models/aggregator.py:201): images
// fp32 [F, 3, H, W] in [0, 1] -> fp16 [F * (H/p) * (W/p), Kp] rows = patches, columns = (channel, ky, kx) of the normalised
// pixel (x - mean[c]) / std[c], zero-padded from 3 p^2 to Kp (a whole number of 64-wide GEMM K blocks).
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __half* __restrict__ out, long long rows, int H, int W, int p, int Kp, float m0, float m1,
                float m2, float s0, float s1, float s2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Kp) return;
  const long long row = i / Kp;
  const int k = (int)(i - row * Kp);
  const int pp = p * p;
  float v = 0.f;
  if (k < 3 * pp) {
    const int c = k / pp, r = k - c * pp;
    const int ky = r / p, kx = r - ky * p;
    const int w0 = W / p, h0 = H / p;
    const long long f = row / ((long long)h0 * w0);
    const int pr = (int)(row - f * h0 * w0);
    const int py = pr / w0, px = pr - py * w0;
    const float x = img[((f * 3 + c) * H + (py * p + ky)) * (long long)W + px * p + kx];
    v = __fdiv_rn(x - (c == 0 ? m0 : (c == 1 ? m1 : m2)), c == 0 ? s0 : (c == 1 ? s1 : s2));
  }
  out[i] = __float2half_rn(v);
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
  const long long items = rows * 2;  // (row, q|k): 8 threads each, looping over the heads
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
  EVW_CHECK_ARG(H <= 65535 && F <= 65535 && (long long)W * (C / 4) < (1ll << 31), "evw_bilinear_ac_f32: output too large");
  const int C4 = C / 4;
  int shift = -1;
  if ((C4 & (C4 - 1)) == 0) {
    shift = 0;
    while ((1 << shift) < C4) ++shift;
  }
  const dim3 grid((unsigned)((W * C4 + 255) / 256), (unsigned)H, (unsigned)F);
  const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f;
  const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
  if (out_fp16)
    evw::bilinear_ac_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, addend, F, h, w, H, W, C4, shift, sy, sx);
  else
    evw::bilinear_ac_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, addend, F, h, w, H, W, C4, shift, sy, sx);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_patchify_f16(const float* images, void* out, int F, int H, int W, int patch, int Kp, const float* mean3,
                                const float* std3, void* stream) {
  EVW_CHECK_ARG(images && out && mean3 && std3 && F >= 1 && patch >= 1 && H >= patch && W >= patch && H % patch == 0 && W % patch == 0,
                "evw_patchify_f16: bad arguments (H and W must be multiples of the patch size)");
  EVW_CHECK_ARG(Kp >= 3 * patch * patch && Kp % 8 == 0, "evw_patchify_f16: Kp=%d must cover 3 p^2 and be a multiple of 8", Kp);
  const long long rows = (long long)F * (H / patch) * (W / patch);
  const long long blocks = (rows * Kp + 255) / 256;
  EVW_CHECK_ARG(blocks < (1ll << 31), "evw_patchify_f16: too many patches");
  evw::patchify_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(images, (__half*)out, rows, H, W, patch, Kp, mean3[0], mean3[1],
                                                                           mean3[2], std3[0], std3[1], std3[2]);
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
