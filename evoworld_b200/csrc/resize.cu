// Pillow-exact bilinear resize of 8-bit images (the PIL round trip of the 24 reprojected memory
// panoramas of a segment, `transforms.Resize((height, width))` in dataset/CameraTrajDataset.py:597-600 applied at
// unified_loop_consistency.py:422).  Pillow (src/libImaging/Resample.c) runs a separable triangle filter whose support
// grows with the down-scaling factor: coefficients normalised in double precision and rounded to 22-bit fixed point
// (computed on the host, evoworld_b200/image_ops.py::pil_resize_tables), a horizontal pass into an 8-bit intermediate, then
// a vertical pass; each pass accumulates integers from 1 << 21 and clips (sum >> 22) to [0, 255].  Integer arithmetic on
// bytes: bit-exact against PIL.Image.resize.  HBM-bound (1000 x 2000 -> 576 x 1024: 6 MB in, 3 MB intermediate, 1.8 MB out
// per image); one thread per output pixel (3 channels), taps read through L1.
#include "common.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace evw {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in [N, H, W, 3] -> out [N, H, w, 3].  One block = 256 consecutive output pixels of one row: the input span they need
// (<= kSpanBytes) is staged in shared memory with aligned 4-byte loads (15 scattered byte loads per thread before), every
// thread then takes its taps from shared memory.
constexpr int kHTile = 256;
constexpr int kSpanBytes = 8192;  // input bytes one block may need; larger spans (down-scaling by > 10) use the direct kernel

__global__ void __launch_bounds__(kHTile)
resize_h_smem_kernel(const uint8_t* __restrict__ in, long long in_bytes, uint8_t* __restrict__ out, int W, int w,
                     const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  __shared__ __align__(16) uint8_t span[kSpanBytes + 8];
  const long long row = blockIdx.y;
  const int x0 = blockIdx.x * kHTile;
  const int x1 = min(w, x0 + kHTile) - 1;  // last output pixel of the block
  const int lo = bounds[2 * x0], hi = bounds[2 * x1] + bounds[2 * x1 + 1];  // input pixels [lo, hi)
  const long long first = (row * W + lo) * 3, last = (row * W + hi) * 3;    // byte range of the span
  const uintptr_t base = reinterpret_cast<uintptr_t>(in);
  const uintptr_t a0 = (base + (uintptr_t)first) & ~(uintptr_t)3;            // aligned start (>= base: allocations are aligned)
  const int shift = (int)(base + (uintptr_t)first - a0);
  const int words = (int)((base + (uintptr_t)last - a0 + 3) >> 2);
  const bool aligned_ok = a0 >= base;
  for (int i = threadIdx.x; i < words; i += kHTile) {
    const uintptr_t a = a0 + 4ull * i;
    uint32_t v;
    if (aligned_ok && a + 4 <= base + (uintptr_t)in_bytes) {
      v = *reinterpret_cast<const uint32_t*>(a);
    } else {  // first / last word of the tensor: byte by byte inside the allocation
      v = 0;
      for (int j = 0; j < 4; ++j)
        if (a + j >= base && a + j < base + (uintptr_t)in_bytes) v |= (uint32_t)(*reinterpret_cast<const uint8_t*>(a + j)) << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(span + 4 * i) = v;
  }
  __syncthreads();
  const int xo = x0 + threadIdx.x;
  if (xo >= w) return;
  const int mylo = bounds[2 * xo], n = bounds[2 * xo + 1];
  const int* k = kk + (long long)xo * ksize;
  const uint8_t* p = span + shift + (mylo - lo) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < n; ++t) {
    const int c = __ldg(k + t);
    s0 += p[3 * t] * c;
    s1 += p[3 * t + 1] * c;
    s2 += p[3 * t + 2] * c;
  }
  uint8_t* o = out + (row * w + xo) * 3;
  o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// direct variant (any span): one thread per output pixel, taps from global memory
__global__ void __launch_bounds__(256)
resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long rows /* N * H */, int W, int w,
                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * w) return;
  const long long row = idx / w;
  const int xo = (int)(idx - row * w);
  const int lo = bounds[2 * xo], n = bounds[2 * xo + 1];
  const int* k = kk + (long long)xo * ksize;
  const uint8_t* p = in + (row * W + lo) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int t = 0; t < n; ++t) {
    const int c = __ldg(k + t);
    s0 += p[3 * t] * c;
    s1 += p[3 * t + 1] * c;
    s2 += p[3 * t + 2] * c;
  }
  uint8_t* o = out + idx * 3;
  o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// in [N, H, w, 3] -> out [N, h, w, 3]; one thread per output byte quad along the row (w * 3 bytes, 4 at a time when aligned)
__global__ void __launch_bounds__(256)
resize_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int N, int H, int h, long long row_bytes /* w * 3 */,
                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const long long quads = (row_bytes + 3) / 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * h * quads) return;
  const long long qi = idx % quads;
  const long long r = idx / quads;
  const int yo = (int)(r % h);
  const long long img = r / h;
  const int lo = bounds[2 * yo], n = bounds[2 * yo + 1];
  const int* k = kk + (long long)yo * ksize;
  const long long b0 = qi * 4;
  const int nb = (int)min(4ll, row_bytes - b0);
  int s[4] = {1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1), 1 << (kPrecisionBits - 1)};
  const uint8_t* p = in + (img * H + lo) * row_bytes + b0;
  const bool word = nb == 4 && (row_bytes & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) & 3) == 0);
  for (int t = 0; t < n; ++t) {
    const int c = __ldg(k + t);
    if (word) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(p + (long long)t * row_bytes);
      s[0] += (int)(v & 255u) * c; s[1] += (int)((v >> 8) & 255u) * c; s[2] += (int)((v >> 16) & 255u) * c; s[3] += (int)(v >> 24) * c;
    } else {
      for (int j = 0; j < nb; ++j) s[j] += p[(long long)t * row_bytes + j] * c;
    }
  }
  uint8_t* o = out + (img * h + yo) * row_bytes + b0;
  if (word && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    *reinterpret_cast<uint32_t*>(o) = (uint32_t)clip8(s[0]) | ((uint32_t)clip8(s[1]) << 8) | ((uint32_t)clip8(s[2]) << 16) |
                                      ((uint32_t)clip8(s[3]) << 24);
  } else {
    for (int j = 0; j < nb; ++j) o[j] = clip8(s[j]);
  }
}

}  // namespace
}  // namespace evw

// in uint8 [N, H, W, 3] -> out uint8 [N, h, w, 3]; tmp uint8 [N, H, w, 3] (unused when w == W: may be null).
// bounds_x int32 [w, 2] / kx int32 [w, ksx], bounds_y [h, 2] / ky [h, ksy]: Pillow's coefficient tables (device memory).
extern "C" int evw_resize_pil_u8(const uint8_t* in, uint8_t* tmp, uint8_t* out, int N, int H, int W, int h, int w,
                                 const int* bounds_x, const int* kx, int ksx, const int* bounds_y, const int* ky, int ksy,
                                 void* stream) {
  EVW_CHECK_ARG(in && out && N >= 1 && H >= 1 && W >= 1 && h >= 1 && w >= 1, "evw_resize_pil_u8: bad arguments");
  EVW_CHECK_ARG((w == W || (bounds_x && kx && ksx >= 1)) && (h == H || (bounds_y && ky && ksy >= 1)) &&
                    (w == W || h == H || tmp),
                "evw_resize_pil_u8: missing tables / intermediate");
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* src = in;
  if (w != W) {  // horizontal pass first, as Pillow does
    uint8_t* dst = (h == H) ? out : tmp;
    const long long n = (long long)N * H * w;
    // input pixels 256 consecutive outputs can span: 256 * scale + the filter support on both sides
    const long long span = (long long)((256.0 * W) / w + 2.0 * ksx + 4.0) * 3;
    if (span <= evw::kSpanBytes && (long long)N * H <= 65535) {
      dim3 grid((unsigned)((w + evw::kHTile - 1) / evw::kHTile), (unsigned)((long long)N * H));
      evw::resize_h_smem_kernel<<<grid, evw::kHTile, 0, st>>>(src, (long long)N * H * W * 3, dst, W, w, bounds_x, kx, ksx);
    } else {
      evw::resize_h_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, (long long)N * H, W, w, bounds_x, kx, ksx);
    }
    EVW_LAUNCH_CHECK();
    src = dst;
  }
  if (h != H) {
    const long long row_bytes = (long long)w * 3;
    const long long n = (long long)N * h * ((row_bytes + 3) / 4);
    evw::resize_v_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, out, N, H, h, row_bytes, bounds_y, ky, ksy);
    EVW_LAUNCH_CHECK();
  } else if (w == W) {
    EVW_CUDA(cudaMemcpyAsync(out, in, (size_t)N * H * W * 3, cudaMemcpyDeviceToDevice, st));
  }
  return EVW_OK;
}
