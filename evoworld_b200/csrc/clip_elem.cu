// Small kernels of the CLIP ViT image encoder (transformers CLIPVisionModelWithProjection as the
// pipeline calls it, evoworld/pipeline/pipeline_evoworld.py:289): self-attention over a few hundred tokens with an
// arbitrary head width (ViT-H: 257 tokens, 16 heads of 80) and the MLP activation.  The linears run on the tcgen05 GEMM
// (tc_gemm.cu), the LayerNorms on layer_norm_kernel (unet_elem.cu); evoworld_b200/clip.py strings them together.
//
// Why not the tcgen05 flash kernel: it is built for head dim 64 and 128-query tiles; one image is 257 x 257 scores per
// head, 0.34 GFLOP per layer — CUDA cores, K / V of one (image, head) resident in shared memory.
#include "common.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace evw {
namespace {

constexpr int kSaWarps = 8;
constexpr int kSaMaxKeysPerLane = 32;  // S <= 1024

// qkv fp16 [B*S, 3*H*D] (q | k | v, heads contiguous inside each) -> out fp16 [B*S, H*D]; one CTA per (head, image,
// query slice).  K and V of the (image, head) sit in shared memory as fp32 [S][D+1] (odd pitch: conflict-free column
// walks); each warp owns one query at a time: lanes split the keys for the scores and the head dimensions for P V.
__global__ void __launch_bounds__(kSaWarps * 32)
small_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int S, int H, int D, float scale, int q_slices) {
  extern __shared__ float smem[];
  const int pitch = D + 1;
  float* ks = smem;                          // [S][pitch]
  float* vs = ks + (size_t)S * pitch;        // [S][pitch]
  float* qs = vs + (size_t)S * pitch;        // [warps][D]
  float* ps = qs + kSaWarps * D;             // [warps][S]
  const int h = blockIdx.x, b = blockIdx.y, slice = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long row0 = (long long)b * S;
  const int ld = 3 * H * D;
  for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
    const int j = i / D, d = i - j * D;
    const __half* r = qkv + (row0 + j) * ld + h * D + d;
    ks[j * pitch + d] = __half2float(r[H * D]);
    vs[j * pitch + d] = __half2float(r[2 * H * D]);
  }
  __syncthreads();
  float* q = qs + warp * D;
  float* p = ps + (size_t)warp * S;
  const int per = (S + q_slices - 1) / q_slices;
  const int q_end = min(S, (slice + 1) * per);
  for (int i = slice * per + warp; i < q_end; i += kSaWarps) {
    for (int d = lane; d < D; d += 32) q[d] = __half2float(qkv[(row0 + i) * ld + h * D + d]) * scale;
    __syncwarp();
    float sc[kSaMaxKeysPerLane];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < kSaMaxKeysPerLane; ++t) {
      const int j = lane + 32 * t;
      if (j < S) {
        const float* kr = ks + j * pitch;
        float a = 0.f;
        for (int d = 0; d < D; ++d) a = fmaf(q[d], kr[d], a);
        sc[t] = a;
        m = fmaxf(m, a);
      }
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < kSaMaxKeysPerLane; ++t) {
      const int j = lane + 32 * t;
      if (j < S) {
        const float e = __expf(sc[t] - m);
        p[j] = e;
        sum += e;
      }
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.0f / sum;
    for (int d = lane; d < D; d += 32) {
      float a = 0.f;
      for (int j = 0; j < S; ++j) a = fmaf(p[j], vs[j * pitch + d], a);
      out[(row0 + i) * (long long)(H * D) + h * D + d] = __float2half_rn(a * inv);
    }
    __syncwarp();
  }
}

// Activation + cast: x fp32 -> fp16.  mode 0 = GELU (erf form, nn.GELU / transformers "gelu"), 1 = quick_gelu x sigmoid(1.702 x),
// 2 = ReLU, 3 = identity (cast only), 4 = SiLU (the last three: VGGT's DPT / camera heads, evoworld_b200/vggt.py)
__global__ void act_f16_kernel(const float* __restrict__ x, __half* __restrict__ out, long long n, int mode) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  float r;
  switch (mode) {
    case 0: r = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); break;
    case 1: r = v / (1.0f + __expf(-1.702f * v)); break;
    case 2: r = fmaxf(v, 0.0f); break;
    case 3: r = v; break;
    default: r = v / (1.0f + expf(-v)); break;
  }
  out[i] = __float2half_rn(r);
}

}  // namespace
}  // namespace evw

extern "C" int evw_small_attention_f16(const void* qkv, void* out, int B, int S, int heads, int head_dim, float scale, void* stream) {
  EVW_CHECK_ARG(qkv && out && B >= 1 && S >= 1 && S <= 32 * evw::kSaMaxKeysPerLane && heads >= 1 && head_dim >= 1 && head_dim <= 256,
                "evw_small_attention_f16: B=%d S=%d heads=%d head_dim=%d not supported (S <= 1024, head_dim <= 256)", B, S, heads,
                head_dim);
  const size_t smem = ((size_t)2 * S * (head_dim + 1) + evw::kSaWarps * (size_t)head_dim + evw::kSaWarps * (size_t)S) * sizeof(float);
  EVW_CHECK_ARG(smem <= 227 * 1024, "evw_small_attention_f16: S=%d x head_dim=%d does not fit into shared memory", S, head_dim);
  static bool attr_set = false;
  if (!attr_set) {
    EVW_CUDA(cudaFuncSetAttribute(evw::small_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  // enough CTAs to cover the SMs: split the queries of an (image, head) over up to 8 CTAs (each re-stages K / V)
  int slices = 1;
  while (slices < 8 && B * heads * slices < 148 && S / (slices * 2) >= evw::kSaWarps) slices *= 2;
  dim3 grid((unsigned)heads, (unsigned)B, (unsigned)slices);
  evw::small_attention_kernel<<<grid, evw::kSaWarps * 32, smem, (cudaStream_t)stream>>>((const __half*)qkv, (__half*)out, S, heads,
                                                                                        head_dim, scale, slices);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_act_f16(const float* x, void* out, int64_t n, int mode, void* stream) {
  EVW_CHECK_ARG(x && out && n >= 0 && mode >= 0 && mode <= 4, "evw_act_f16: bad arguments");
  if (n == 0) return EVW_OK;
  evw::act_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, (__half*)out, n, mode);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
