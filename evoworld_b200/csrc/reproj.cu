// Hot path 2 — 3D-memory reprojection kernels (HBM-bound gather/scatter; no tensor cores).
//
//   evw_plucker                 utils/plucker_embedding.py:221-255
//   evw_equi2pers_u8            equilib.Equi2Pers (pyequilib 0.5.8), unified_loop_consistency.py:329
//   evw_lift_depth              third_party/vggt/vggt/utils/geometry.py:12-111
//   evw_pack_points             reproject_vggt_open3d_utils.py:286-292 (+ Open3D's f64->f32 upload)
//   evw_conf_select             reproject_vggt_open3d_utils.py:294-310
//   evw_splat_cubemap_equirect  reproject_vggt_open3d_utils.py:617-711 + :542-614
//
// The integer part of the path (pixel index, z-buffer key, winning point index, cube->equirect
// gather index) is specified so that it is bit-reproducible: the projection uses explicit fmaf /
// IEEE division in a fixed order and is mirrored instruction-for-instruction by oracle/reproj_oracle.c.
#include "common.h"
#include <mutex>
#include <math_constants.h>
#include <cstdlib>

namespace {

constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------------------
// Plücker
// ------------------------------------------------------------------------------------------
__global__ void plucker_kernel(const float* __restrict__ ray, const float* __restrict__ c2w,
                               float* __restrict__ out, int T, int HW) {
  extern __shared__ float s_c2w[];  // [T,12]
  for (int i = threadIdx.x; i < T * 12; i += blockDim.x) s_c2w[i] = c2w[i];
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float dx = ray[p * 3 + 0], dy = ray[p * 3 + 1], dz = ray[p * 3 + 2];
  for (int n = 0; n < T; ++n) {
    const float* m = s_c2w + n * 12;
    // same summation order as a plain i,j loop: ((R0*dx + R1*dy) + R2*dz), no contraction
    float wx = __fadd_rn(__fadd_rn(__fmul_rn(m[0], dx), __fmul_rn(m[1], dy)), __fmul_rn(m[2], dz));
    float wy = __fadd_rn(__fadd_rn(__fmul_rn(m[4], dx), __fmul_rn(m[5], dy)), __fmul_rn(m[6], dz));
    float wz = __fadd_rn(__fadd_rn(__fmul_rn(m[8], dx), __fmul_rn(m[9], dy)), __fmul_rn(m[10], dz));
    float tx = m[3], ty = m[7], tz = m[11];
    float* o = out + (size_t)n * 6 * HW + p;
    o[0 * (size_t)HW] = wx;
    o[1 * (size_t)HW] = wy;
    o[2 * (size_t)HW] = wz;
    o[3 * (size_t)HW] = __fsub_rn(__fmul_rn(ty, wz), __fmul_rn(tz, wy));
    o[4 * (size_t)HW] = __fsub_rn(__fmul_rn(tz, wx), __fmul_rn(tx, wz));
    o[5 * (size_t)HW] = __fsub_rn(__fmul_rn(tx, wy), __fmul_rn(ty, wx));
  }
}

// ------------------------------------------------------------------------------------------
// Equirect -> perspective, uint8 bilinear (horizontal and vertical wrap as pyequilib's sampler)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
equi2pers_kernel(const uint8_t* __restrict__ equi, const float* __restrict__ pix2dir, uint8_t* __restrict__ out, int C,
                 int He, int We, int Hp, int Wp) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  __shared__ float A[9];  // one matrix per frame: read once per block instead of nine global loads per pixel
  if (threadIdx.y == 0 && threadIdx.x < 9) A[threadIdx.x] = pix2dir[b * 9 + threadIdx.x];
  __syncthreads();
  if (x >= Wp || y >= Hp) return;
  float fx = (float)x, fy = (float)y;
  float mx = __fadd_rn(__fadd_rn(__fmul_rn(A[0], fx), __fmul_rn(A[1], fy)), A[2]);
  float my = __fadd_rn(__fadd_rn(__fmul_rn(A[3], fx), __fmul_rn(A[4], fy)), A[5]);
  float mz = __fadd_rn(__fadd_rn(__fmul_rn(A[6], fx), __fmul_rn(A[7], fy)), A[8]);
  float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
  float phi = asinf(__fdiv_rn(mz, nrm));
  float theta = atan2f(my, mx);
  const float PI = 3.14159265358979323846f;
  float ui = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(theta, PI), (float)We), __fmul_rn(2.0f, PI)), 0.5f);
  float uj = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(phi, __fmul_rn(0.5f, PI)), (float)He), PI), 0.5f);
  // theta in [-pi, pi] and phi in [-pi/2, pi/2] put ui in [-We + 0.5, 0.5] and uj in [-He + 0.5, 0.5]: |u| < extent, where
  // fmod is the identity.  libdevice's fmodf and the integer '%' were most of this kernel's instructions; the general
  // path stays behind a never-taken branch so the result is the same expression in every case.
  if (!(fabsf(ui) < (float)We)) ui = fmodf(ui, (float)We);
  if (ui < 0.f) ui = __fadd_rn(ui, (float)We);
  if (!(fabsf(uj) < (float)He)) uj = fmodf(uj, (float)He);
  if (uj < 0.f) uj = __fadd_rn(uj, (float)He);
  float x0f = floorf(ui), y0f = floorf(uj);
  float dx = __fsub_rn(ui, x0f), dy = __fsub_rn(uj, y0f);
  int x0 = (int)x0f, y0 = (int)y0f;  // in [0, extent]: one conditional subtraction is the modulo
  if (x0 >= We) x0 -= We;
  if (y0 >= He) y0 -= He;
  int x1 = x0 + 1, y1 = y0 + 1;
  if (x1 >= We) x1 -= We;
  if (y1 >= He) y1 -= He;
  float wx0 = __fsub_rn(1.0f, dx), wy0 = __fsub_rn(1.0f, dy);
  // 32-bit offsets inside one image plane; the plane base is the only 64-bit address
  const unsigned o00 = (unsigned)y0 * We + x0, o01 = (unsigned)y0 * We + x1;
  const unsigned o10 = (unsigned)y1 * We + x0, o11 = (unsigned)y1 * We + x1;
  const size_t plane_in = (size_t)He * We, plane_out = (size_t)Hp * Wp;
  const uint8_t* img = equi + (size_t)b * C * plane_in;
  uint8_t* dst = out + (size_t)b * C * plane_out + (unsigned)y * Wp + x;
  for (int c = 0; c < C; ++c, img += plane_in, dst += plane_out) {
    float q00 = __ldg(img + o00), q01 = __ldg(img + o01), q10 = __ldg(img + o10), q11 = __ldg(img + o11);
    float top = __fadd_rn(__fmul_rn(q00, wx0), __fmul_rn(q01, dx));
    float bot = __fadd_rn(__fmul_rn(q10, wx0), __fmul_rn(q11, dx));
    float v = __fadd_rn(__fmul_rn(top, wy0), __fmul_rn(bot, dy));
    v = fminf(fmaxf(v, 0.f), 255.f);
    *dst = (uint8_t)v;  // truncation, as astype(uint8)
  }
}

// Pure-yaw fast path (the only rotation EvoWorld uses: unified_loop_consistency.py:329 passes pitch = roll = 0).  A rotation
// about the vertical axis only shifts the longitude, so the per-pixel (theta, phi) -> (ui, uj) of the yaw = 0 camera is
// tabulated once per (Hp, Wp, fov, He, We) with exactly the expressions of equi2pers_kernel, and a frame is a pure gather:
// ui = ui0 + yaw_shift (in source pixels), wrapped — no asinf / atan2f / sqrtf per pixel and frame (the general kernel
// spent its time there: 8.5 % of the HBM roofline, profiles/r01g_secondary_bench.log).
__global__ void equi2pers_table_kernel(const float* __restrict__ pix2dir0, float2* __restrict__ table, int He, int We, int Hp,
                                       int Wp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= Wp || y >= Hp) return;
  const float* A = pix2dir0;
  float fx = (float)x, fy = (float)y;
  float mx = __fadd_rn(__fadd_rn(__fmul_rn(A[0], fx), __fmul_rn(A[1], fy)), A[2]);
  float my = __fadd_rn(__fadd_rn(__fmul_rn(A[3], fx), __fmul_rn(A[4], fy)), A[5]);
  float mz = __fadd_rn(__fadd_rn(__fmul_rn(A[6], fx), __fmul_rn(A[7], fy)), A[8]);
  float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(mx, mx), __fmul_rn(my, my)), __fmul_rn(mz, mz)));
  float phi = asinf(__fdiv_rn(mz, nrm));
  float theta = atan2f(my, mx);
  const float PI = 3.14159265358979323846f;
  float ui = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(theta, PI), (float)We), __fmul_rn(2.0f, PI)), 0.5f);
  float uj = __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(phi, __fmul_rn(0.5f, PI)), (float)He), PI), 0.5f);
  if (!(fabsf(uj) < (float)He)) uj = fmodf(uj, (float)He);
  if (uj < 0.f) uj = __fadd_rn(uj, (float)He);
  table[(size_t)y * Wp + x] = make_float2(ui, uj);  // ui unwrapped (in [-We + 0.5, 0.5]); uj wrapped to [0, He)
}

__global__ void __launch_bounds__(256)
equi2pers_yaw_kernel(const uint8_t* __restrict__ equi, const float2* __restrict__ table, const float* __restrict__ shift_px,
                     uint8_t* __restrict__ out, int C, int He, int We, int Hp, int Wp) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= Wp || y >= Hp) return;
  const float2 t = __ldg(table + (size_t)y * Wp + x);
  float ui = __fadd_rn(t.x, __ldg(shift_px + b));  // |shift| <= We / 2 (the host reduces the yaw to [-pi, pi])
  const float uj = t.y;
  const float fWe = (float)We;
  if (!(fabsf(ui) < fWe)) ui = fmodf(ui, fWe);
  if (ui < 0.f) ui = __fadd_rn(ui, fWe);
  float x0f = floorf(ui), y0f = floorf(uj);
  float dx = __fsub_rn(ui, x0f), dy = __fsub_rn(uj, y0f);
  int x0 = (int)x0f, y0 = (int)y0f;
  if (x0 >= We) x0 -= We;
  if (y0 >= He) y0 -= He;
  int x1 = x0 + 1, y1 = y0 + 1;
  if (x1 >= We) x1 -= We;
  if (y1 >= He) y1 -= He;
  float wx0 = __fsub_rn(1.0f, dx), wy0 = __fsub_rn(1.0f, dy);
  const unsigned o00 = (unsigned)y0 * We + x0, o01 = (unsigned)y0 * We + x1;
  const unsigned o10 = (unsigned)y1 * We + x0, o11 = (unsigned)y1 * We + x1;
  const size_t plane_in = (size_t)He * We, plane_out = (size_t)Hp * Wp;
  const uint8_t* img = equi + (size_t)b * C * plane_in;
  uint8_t* dst = out + (size_t)b * C * plane_out + (unsigned)y * Wp + x;
  for (int c = 0; c < C; ++c, img += plane_in, dst += plane_out) {
    float q00 = __ldg(img + o00), q01 = __ldg(img + o01), q10 = __ldg(img + o10), q11 = __ldg(img + o11);
    float top = __fadd_rn(__fmul_rn(q00, wx0), __fmul_rn(q01, dx));
    float bot = __fadd_rn(__fmul_rn(q10, wx0), __fmul_rn(q11, dx));
    float v = __fadd_rn(__fmul_rn(top, wy0), __fmul_rn(bot, dy));
    v = fminf(fmaxf(v, 0.f), 255.f);
    *dst = (uint8_t)v;  // truncation, as astype(uint8)
  }
}

// ------------------------------------------------------------------------------------------
// Depth lift (float64 world transform as the numpy reference)
// ------------------------------------------------------------------------------------------
// World position of pixel p of frame s: camera coordinates rounded to float32 (geometry.py:104-109), float64 world
// transform with c2w = [R^T | -R^T t] where -R^T t is evaluated in float32 (numpy matmul of f32 arrays).
__device__ __forceinline__ void lift_point(const float* __restrict__ depth, const float* __restrict__ extr,
                                           const float* __restrict__ intr, int s, int p, int H, int W, double& wx, double& wy,
                                           double& wz) {
  const float* E = extr + s * 12;
  const float* K = intr + s * 9;
  float r[9], t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = E[j * 4 + i];
    float acc = __fmul_rn(E[0 * 4 + i], E[3]);
    acc = __fadd_rn(acc, __fmul_rn(E[1 * 4 + i], E[7]));
    acc = __fadd_rn(acc, __fmul_rn(E[2 * 4 + i], E[11]));
    t[i] = -acc;
  }
  const int v = p / W, u = p - v * W;
  const double d = (double)depth[(size_t)s * H * W + p];
  const float xc = (float)(((double)u - (double)K[2]) * d / (double)K[0]);
  const float yc = (float)(((double)v - (double)K[5]) * d / (double)K[4]);
  const float zc = (float)d;
  wx = ((double)xc * r[0] + (double)yc * r[1]) + (double)zc * r[2] + (double)t[0];
  wy = ((double)xc * r[3] + (double)yc * r[4]) + (double)zc * r[5] + (double)t[1];
  wz = ((double)xc * r[6] + (double)yc * r[7]) + (double)zc * r[8] + (double)t[2];
}

// Fused lift + pack for the device-resident point memory: depth + NCHW float colours -> float4 {x,y,z,rgb} per pixel, the
// same bits as lift_kernel (f64) followed by pack_points_kernel (f64 -> f32 round, colour = trunc(x*255)), without the
// 24 B/point float64 intermediate: 16 B in (depth + 3 colour floats), 16 B out.
__global__ void lift_pack_kernel(const float* __restrict__ depth, const float* __restrict__ extr, const float* __restrict__ intr,
                                 const float* __restrict__ images, float4* __restrict__ out, int H, int W) {
  const int s = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (p >= HW) return;
  double wx, wy, wz;
  lift_point(depth, extr, intr, s, p, H, W, wx, wy, wz);
  const float* im = images + (size_t)s * 3 * HW + p;
  const unsigned r = (unsigned)(uint8_t)(int)__fmul_rn(im[0], 255.0f);
  const unsigned g = (unsigned)(uint8_t)(int)__fmul_rn(im[(size_t)HW], 255.0f);
  const unsigned b = (unsigned)(uint8_t)(int)__fmul_rn(im[2 * (size_t)HW], 255.0f);
  out[(size_t)s * HW + p] = make_float4((float)wx, (float)wy, (float)wz, __uint_as_float(r | (g << 8) | (b << 16)));
}

__global__ void lift_kernel(const float* __restrict__ depth, const float* __restrict__ extr,
                            const float* __restrict__ intr, double* __restrict__ out64,
                            float* __restrict__ out32, int H, int W) {
  int s = blockIdx.y;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const float* E = extr + s * 12;
  const float* K = intr + s * 9;
  // c2w = [R^T | -R^T t]; the reference evaluates -R^T t in float32 (numpy matmul of f32 arrays)
  float r[9], t[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = E[j * 4 + i];
    float acc = __fmul_rn(E[0 * 4 + i], E[3]);
    acc = __fadd_rn(acc, __fmul_rn(E[1 * 4 + i], E[7]));
    acc = __fadd_rn(acc, __fmul_rn(E[2 * 4 + i], E[11]));
    t[i] = -acc;
  }
  int v = p / W, u = p - v * W;
  double d = (double)depth[(size_t)s * H * W + p];
  // (u - cu) * depth / fu in float64, then cast to float32 (geometry.py:104-109)
  float xc = (float)(((double)u - (double)K[2]) * d / (double)K[0]);
  float yc = (float)(((double)v - (double)K[5]) * d / (double)K[4]);
  float zc = (float)d;
  double wx = ((double)xc * r[0] + (double)yc * r[1]) + (double)zc * r[2] + (double)t[0];
  double wy = ((double)xc * r[3] + (double)yc * r[4]) + (double)zc * r[5] + (double)t[1];
  double wz = ((double)xc * r[6] + (double)yc * r[7]) + (double)zc * r[8] + (double)t[2];
  size_t o = ((size_t)s * H * W + p) * 3;
  if (out64) {
    out64[o] = wx;
    out64[o + 1] = wy;
    out64[o + 2] = wz;
  }
  if (out32) {
    out32[o] = (float)wx;
    out32[o + 1] = (float)wy;
    out32[o + 2] = (float)wz;
  }
}

// ------------------------------------------------------------------------------------------
// Pack {xyz, rgb} -> float4
// ------------------------------------------------------------------------------------------
__global__ void pack_points_kernel(const double* __restrict__ xyz64, const float* __restrict__ xyz32,
                                   const uint8_t* __restrict__ rgb, const float* __restrict__ images,
                                   int HW, float4* __restrict__ out, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float x, y, z;
  if (xyz64) {
    x = (float)xyz64[i * 3];
    y = (float)xyz64[i * 3 + 1];
    z = (float)xyz64[i * 3 + 2];
  } else {
    x = xyz32[i * 3];
    y = xyz32[i * 3 + 1];
    z = xyz32[i * 3 + 2];
  }
  unsigned r, g, b;
  if (rgb) {
    r = rgb[i * 3];
    g = rgb[i * 3 + 1];
    b = rgb[i * 3 + 2];
  } else {
    int64_t s = i / HW, p = i - s * HW;
    const float* im = images + (size_t)s * 3 * HW + p;
    r = (unsigned)(uint8_t)(int)__fmul_rn(im[0], 255.0f);
    g = (unsigned)(uint8_t)(int)__fmul_rn(im[(size_t)HW], 255.0f);
    b = (unsigned)(uint8_t)(int)__fmul_rn(im[2 * (size_t)HW], 255.0f);
  }
  out[i] = make_float4(x, y, z, __uint_as_float(r | (g << 8) | (b << 16)));
}

// ------------------------------------------------------------------------------------------
// Confidence select: exact k-th order statistics by 3-pass radix select on float bits, numpy-lerp
// threshold, then order-preserving stream compaction.
// ------------------------------------------------------------------------------------------
struct SelectState {
  unsigned prefix;        // resolved high bits of the k_lo-th key
  unsigned mask;          // which bits of prefix are resolved
  unsigned long long k;   // remaining rank inside the current prefix bucket
  unsigned long long cnt_le;  // #keys <= key(k_lo)
  unsigned next_key;      // min key > key(k_lo)
  unsigned nan_count;
  float thr;
  unsigned pad;
};

__device__ __forceinline__ unsigned float_key(float f) {
  unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
  unsigned b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(b);
}

__host__ __device__ constexpr int radix_bits(int pass) { return pass == 2 ? 10 : 11; }
__host__ __device__ constexpr int radix_shift(int pass) { return pass == 0 ? 21 : (pass == 1 ? 10 : 0); }

// Block-wide search of the histogram bin that holds rank k (the scan the round-1 version ran as a kernel of its own after
// every histogram pass; every block of the NEXT pass now redoes it in its prologue: 8 KiB of L2 reads and a 512-thread scan
// against a kernel launch + a dependent-launch gap).  All threads return the same result.
struct BinHit {
  unsigned bin;              // bin that holds rank k (the last bin if k is beyond the valid count: NaNs present -> thr = NaN anyway)
  unsigned cnt;              // keys in that bin
  unsigned long long k_rem;  // rank inside the bin
  unsigned next_bin;         // next non-empty bin above it, 0xFFFFFFFF if none
};

template <int PASS>
__device__ BinHit find_bin(const unsigned* __restrict__ hist, unsigned long long k) {
  constexpr int nb = 1 << radix_bits(PASS);
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_krem;
  __shared__ unsigned s_bin, s_cnt, s_next;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
  const int per = (nb + nthreads - 1) / nthreads;
  const int lo = tid * per;
  __syncthreads();  // previous use of the shared scratch
  if (tid == 0) { s_bin = 0xFFFFFFFFu; s_next = 0xFFFFFFFFu; }
  unsigned long long mine = 0;
  for (int b = lo; b < lo + per && b < nb; ++b) mine += hist[PASS * 2048 + b];
  unsigned long long incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  unsigned long long run = incl - mine;
  for (int w = 0; w < warp; ++w) run += s_warp[w];
  if (run <= k && k < run + mine) {  // at most one thread
    for (int b = lo; b < lo + per && b < nb; ++b) {
      const unsigned c = hist[PASS * 2048 + b];
      if (k < run + c) {
        s_bin = (unsigned)b; s_cnt = c; s_krem = k - run;
        break;
      }
      run += c;
    }
  }
  __syncthreads();
  if (s_bin == 0xFFFFFFFFu) {
    if (tid == 0) { s_bin = nb - 1; s_cnt = hist[PASS * 2048 + nb - 1]; s_krem = 0; }
    __syncthreads();
  }
  const unsigned bin = s_bin;
  unsigned nxt = 0xFFFFFFFFu;
  for (int b = lo; b < lo + per && b < nb; ++b)
    if ((unsigned)b > bin && hist[PASS * 2048 + b]) { nxt = (unsigned)b; break; }
  for (int o = 16; o > 0; o >>= 1) nxt = min(nxt, __shfl_xor_sync(0xffffffffu, nxt, o));
  if (lane == 0 && nxt != 0xFFFFFFFFu) atomicMin(&s_next, nxt);
  __syncthreads();
  BinHit h;
  h.bin = bin; h.cnt = s_cnt; h.k_rem = s_krem; h.next_bin = s_next;
  return h;
}

// Histogram pass PASS over the keys that match the digits resolved so far.  Counts go to shared memory with
// warp-aggregated atomics (confidence maps concentrate on a few dozen top-digit bins: a plain shared atomicAdd per element
// serialised 8-16 ways), then to the global histogram.  PASS 0 counts NaNs; PASS 2 also tracks the smallest key ABOVE the
// 22-bit bucket, so that the successor of the selected key is known without another pass over the data.
template <int PASS>
__global__ void __launch_bounds__(512)
select_hist_kernel(const float* __restrict__ conf, int64_t n, SelectState* __restrict__ st, unsigned* __restrict__ hist,
                   unsigned long long k_lo) {
  __shared__ unsigned sh[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sh[i] = 0;
  unsigned prefix = 0, mask = 0;
  if (PASS >= 1) {
    const BinHit h0 = find_bin<0>(hist, k_lo);
    prefix = h0.bin << radix_shift(0);
    mask = ((1u << radix_bits(0)) - 1u) << radix_shift(0);
    if (PASS >= 2) {
      const BinHit h1 = find_bin<1>(hist, h0.k_rem);
      prefix |= h1.bin << radix_shift(1);
      mask |= ((1u << radix_bits(1)) - 1u) << radix_shift(1);
    }
  }
  __syncthreads();
  constexpr int shift = radix_shift(PASS);
  constexpr unsigned dmask = (1u << radix_bits(PASS)) - 1u;
  unsigned nan_local = 0, above = 0xFFFFFFFFu;
  // one element; executed by all 32 lanes of a warp together (the aggregated atomic of pass 0 is a warp collective)
  auto take = [&](float f, bool valid) {
    unsigned digit = 0xFFFFFFFFu;  // "no contribution"
    if (valid) {
      if (f != f) {
        if (PASS == 0) ++nan_local;
      } else {
        const unsigned k = float_key(f);
        if ((k & mask) == prefix) digit = (k >> shift) & dmask;
        else if (PASS == 2 && (k & mask) > prefix) above = min(above, k);
      }
    }
    if (PASS == 0) {  // every element contributes and the top digits collide: one atomic per distinct digit of the warp
      const unsigned peers = __match_any_sync(0xffffffffu, digit);
      if (digit != 0xFFFFFFFFu && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh[digit], (unsigned)__popc(peers));
    } else if (digit != 0xFFFFFFFFu) {  // only the keys of one bucket contribute: few, and spread over the digit
      atomicAdd(&sh[digit], 1u);
    }
  };
  // 16-byte loads, two in flight per thread: with one 4-byte load per iteration the pass was bound by DRAM latency
  // (20 us for 20 MB, profiles/r02u_reproj_launches.csv.gz), not by bandwidth or by the atomics
  const bool vec = (reinterpret_cast<uintptr_t>(conf) & 15) == 0;
  const int64_t n4 = vec ? (n >> 2) : 0;
  const float4* conf4 = reinterpret_cast<const float4*>(conf);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n4; base += 2 * stride) {  // warp-uniform trip count
    const int64_t i0 = base + threadIdx.x, i1 = i0 + stride;
    const bool v0 = i0 < n4, v1 = i1 < n4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (v0) a = conf4[i0];
    if (v1) b = conf4[i1];
    take(a.x, v0); take(a.y, v0); take(a.z, v0); take(a.w, v0);
    take(b.x, v1); take(b.y, v1); take(b.z, v1); take(b.w, v1);
  }
  for (int64_t base = 4 * n4 + (int64_t)blockIdx.x * blockDim.x; base < n; base += stride) {  // tail (all of it if unaligned)
    const int64_t i = base + threadIdx.x;
    const bool v = i < n;
    take(v ? conf[i] : 0.f, v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[PASS * 2048 + i], sh[i]);
  if (PASS == 0) {
    for (int o = 16; o > 0; o >>= 1) nan_local += __shfl_xor_sync(0xffffffffu, nan_local, o);
    if ((threadIdx.x & 31) == 0 && nan_local) atomicAdd(&st->nan_count, nan_local);
  }
  if (PASS == 2) {
    for (int o = 16; o > 0; o >>= 1) above = min(above, __shfl_xor_sync(0xffffffffu, above, o));
    if ((threadIdx.x & 31) == 0 && above != 0xFFFFFFFFu) atomicMin(&st->next_key, above);
  }
}

// thr = numpy _lerp(a, b, t) in float32:  a + (b-a)*t, and for t >= 0.5:  b - (b-a)*(1-t)
__global__ void __launch_bounds__(512)
select_thr_kernel(SelectState* st, const unsigned* __restrict__ hist, long long k_lo, long long k_hi, float gamma,
                  int use_threshold, float* out_thr) {
  float thr = 0.0f;
  if (use_threshold) {  // uniform
    const BinHit h0 = find_bin<0>(hist, (unsigned long long)k_lo);
    const BinHit h1 = find_bin<1>(hist, h0.k_rem);
    const BinHit h2 = find_bin<2>(hist, h1.k_rem);
    const unsigned key_a = (h0.bin << radix_shift(0)) | (h1.bin << radix_shift(1)) | h2.bin;
    // every key of the final bin equals key_a: #keys <= key_a = (#keys below the bin) + (#keys in it)
    const unsigned long long cnt_le = ((unsigned long long)k_lo - h2.k_rem) + h2.cnt;
    // successor of key_a: the next non-empty bin of the same 22-bit bucket, else the smallest key above the bucket (pass 2)
    const unsigned next_key = h2.next_bin != 0xFFFFFFFFu ? ((key_a & ~((1u << radix_bits(2)) - 1u)) | h2.next_bin) : st->next_key;
    if (st->nan_count) {
      thr = CUDART_NAN_F;
    } else {
      const float a = key_float(key_a);
      const float b = (cnt_le >= (unsigned long long)k_hi + 1ull) ? a : key_float(next_key);
      const float diff = __fsub_rn(b, a);
      thr = __fadd_rn(a, __fmul_rn(diff, gamma));
      if (gamma >= 0.5f) thr = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, gamma)));
    }
    if (threadIdx.x == 0) { st->prefix = key_a; st->cnt_le = cnt_le; }
  }
  if (threadIdx.x == 0) {
    st->thr = thr;
    if (out_thr) *out_thr = thr;
  }
}

constexpr int kCompactThreads = 256;
constexpr int kCompactItems = 8;  // items per thread, consecutive
constexpr int kCompactTile = kCompactThreads * kCompactItems;

__global__ void compact_count_kernel(const float* __restrict__ conf, int64_t n,
                                     const SelectState* __restrict__ st,
                                     unsigned* __restrict__ block_counts) {
  const float thr = st->thr;
  int64_t base = (int64_t)blockIdx.x * kCompactTile + (int64_t)threadIdx.x * kCompactItems;
  unsigned c = 0;
  if (base + kCompactItems <= n && (reinterpret_cast<uintptr_t>(conf) & 15) == 0) {  // two 16-byte loads
    const float4 a = *reinterpret_cast<const float4*>(conf + base), b = *reinterpret_cast<const float4*>(conf + base + 4);
    c = (a.x >= thr) + (a.y >= thr) + (a.z >= thr) + (a.w >= thr) + (b.x >= thr) + (b.y >= thr) + (b.z >= thr) + (b.w >= thr);
  } else {
#pragma unroll
    for (int j = 0; j < kCompactItems; ++j) {
      int64_t i = base + j;
      if (i < n && conf[i] >= thr) ++c;
    }
  }
  __shared__ unsigned sw[kCompactThreads / 32];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = 0;
    for (int w = 0; w < kCompactThreads / 32; ++w) s += sw[w];
    block_counts[blockIdx.x] = s;
  }
}

// single block exclusive scan over block_counts -> block_offsets (int64), total -> out_count
__global__ void compact_scan_kernel(const unsigned* __restrict__ block_counts, int nblocks,
                                    long long* __restrict__ block_offsets, long long* out_count) {
  __shared__ long long s_part[1024];
  const int tid = threadIdx.x;
  const int per = (nblocks + blockDim.x - 1) / blockDim.x;
  const int lo = tid * per, hi = min(nblocks, lo + per);
  long long sum = 0;
  for (int i = lo; i < hi; ++i) sum += block_counts[i];
  s_part[tid] = sum;
  __syncthreads();
  {  // exclusive scan of the 1024 partial sums: warp shuffles + one pass over the 32 warp totals
    __shared__ long long s_warp[32];
    const int lane = tid & 31, warp = tid >> 5;
    long long incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    long long base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    s_part[tid] = base + incl - sum;
    if (tid == (int)blockDim.x - 1) *out_count = base + incl;
  }
  __syncthreads();
  long long run = s_part[tid];
  for (int i = lo; i < hi; ++i) {
    block_offsets[i] = run;
    run += block_counts[i];
  }
}

__global__ void compact_scatter_kernel(const float* __restrict__ conf, const float4* __restrict__ pts,
                                       int64_t n, const SelectState* __restrict__ st,
                                       const long long* __restrict__ block_offsets,
                                       float4* __restrict__ out, long long* __restrict__ keep_idx) {
  const float thr = st->thr;
  int64_t base = (int64_t)blockIdx.x * kCompactTile + (int64_t)threadIdx.x * kCompactItems;
  unsigned flags = 0, c = 0;
  if (base + kCompactItems <= n && (reinterpret_cast<uintptr_t>(conf) & 15) == 0) {  // two 16-byte loads
    const float4 a = *reinterpret_cast<const float4*>(conf + base), b = *reinterpret_cast<const float4*>(conf + base + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < kCompactItems; ++j)
      if (v[j] >= thr) flags |= 1u << j;
    c = __popc(flags);
  } else {
#pragma unroll
    for (int j = 0; j < kCompactItems; ++j) {
      int64_t i = base + j;
      if (i < n && conf[i] >= thr) {
        flags |= 1u << j;
        ++c;
      }
    }
  }
  // block exclusive scan of c
  __shared__ unsigned sw[kCompactThreads / 32];
  unsigned incl = c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) sw[warp] = incl;
  __syncthreads();
  unsigned woff = 0;
  for (int w = 0; w < warp; ++w) woff += sw[w];
  long long pos = block_offsets[blockIdx.x] + woff + (incl - c);
#pragma unroll
  for (int j = 0; j < kCompactItems; ++j) {
    if (flags & (1u << j)) {
      int64_t i = base + j;
      if (out) out[pos] = pts[i];
      if (keep_idx) keep_idx[pos] = i;
      ++pos;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Splat: project every point into the 6 faces of G views, 64-bit atomicMin z-buffer
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void splat_one(const float* __restrict__ m, float x, float y, float z,
                                          float focal, float c, float z_near, int res,
                                          unsigned long long* __restrict__ zface, unsigned idx) {
  float zc = fmaf(m[8], x, fmaf(m[9], y, fmaf(m[10], z, m[11])));
  if (!(zc > z_near)) return;
  float xc = fmaf(m[0], x, fmaf(m[1], y, fmaf(m[2], z, m[3])));
  float yc = fmaf(m[4], x, fmaf(m[5], y, fmaf(m[6], z, m[7])));
  float u = fmaf(focal, __fdiv_rn(xc, zc), c);
  float v = fmaf(focal, __fdiv_rn(yc, zc), c);
  float fres = (float)res;
  if (!(u >= 0.f && u < fres && v >= 0.f && v < fres)) return;
  int px = (int)floorf(u), py = (int)floorf(v);
  unsigned long long key = ((unsigned long long)__float_as_uint(zc) << 32) | idx;
  unsigned long long* cell = zface + (size_t)py * res + px;
  // cheap pre-test (plain L2 read) saves most of the losing atomics
  if (key < ld_relaxed_u64(cell)) atomicMin(cell, key);
}

template <int G>
__global__ void __launch_bounds__(256)
splat_kernel(const float4* __restrict__ pts, int64_t n_cap, const long long* __restrict__ n_dev,
             const float* __restrict__ w2c /*[G,6,12]*/, int res, float focal, float z_near,
             unsigned long long* __restrict__ zbuf /*[G,6,res,res]*/) {
  __shared__ float s_m[G * 6 * 12];
  for (int i = threadIdx.x; i < G * 72; i += blockDim.x) s_m[i] = w2c[i];
  __syncthreads();
  int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
  const float c = 0.5f * (float)res;
  const size_t face_sz = (size_t)res * res;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = ld_stream_f4(pts + i);
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int f = 0; f < 6; ++f)
        splat_one(s_m + (g * 6 + f) * 12, p.x, p.y, p.z, focal, c, z_near, res,
                  zbuf + (size_t)(g * 6 + f) * face_sz, (unsigned)i);
    }
  }
}

// resolve: 4 output pixels per thread (12 bytes = 3 x u32 stores)
__global__ void resolve_kernel(const unsigned long long* __restrict__ zbuf /*[6,res,res]*/,
                               const float4* __restrict__ pts, const uint32_t* __restrict__ lut,
                               int res, int64_t npix, uint8_t* __restrict__ out) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // quad index
  int64_t p0 = q * 4;
  if (p0 >= npix) return;
  unsigned rgb[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    rgb[j] = 0;
    int64_t p = p0 + j;
    if (p < npix) {
      uint32_t e = lut[p];
      if (e != 0xFFFFFFFFu) {
        unsigned face = e >> 28, row = (e >> 14) & 0x3FFFu, col = e & 0x3FFFu;
        unsigned long long key = zbuf[((size_t)face * res + row) * res + col];
        if (key != kEmptyKey) {
          unsigned idx = (unsigned)(key & 0xFFFFFFFFull);
          rgb[j] = __float_as_uint(__ldg(&pts[idx].w)) & 0xFFFFFFu;
        }
      }
    }
  }
  if (p0 + 3 < npix) {
    uint32_t w0 = rgb[0] | (rgb[1] << 24);
    uint32_t w1 = (rgb[1] >> 8) | (rgb[2] << 16);
    uint32_t w2 = (rgb[2] >> 16) | (rgb[3] << 8);
    uint32_t* o = reinterpret_cast<uint32_t*>(out + p0 * 3);  // p0*3 is a multiple of 12
    o[0] = w0;
    o[1] = w1;
    o[2] = w2;
  } else {
    for (int j = 0; j < 4 && p0 + j < npix; ++j) {
      out[(p0 + j) * 3 + 0] = rgb[j] & 0xFF;
      out[(p0 + j) * 3 + 1] = (rgb[j] >> 8) & 0xFF;
      out[(p0 + j) * 3 + 2] = (rgb[j] >> 16) & 0xFF;
    }
  }
}

__global__ void zbuf_to_index_kernel(const unsigned long long* __restrict__ zbuf, int64_t n,
                                     long long* __restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long k = zbuf[i];
  idx[i] = (k == kEmptyKey) ? -1ll : (long long)(k & 0xFFFFFFFFull);
}

// face images from the z-buffer: out [6,res,res,3] u8 (HWC, as render_to_image returns)
__global__ void zbuf_to_rgb_kernel(const unsigned long long* __restrict__ zbuf,
                                   const float4* __restrict__ pts, int64_t n, uint8_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long k = zbuf[i];
  unsigned rgb = 0;
  if (k != kEmptyKey) rgb = __float_as_uint(__ldg(&pts[(unsigned)(k & 0xFFFFFFFFull)].w)) & 0xFFFFFFu;
  out[i * 3 + 0] = rgb & 0xFF;
  out[i * 3 + 1] = (rgb >> 8) & 0xFF;
  out[i * 3 + 2] = (rgb >> 16) & 0xFF;
}

// cube faces [B,6,3,res,res] u8 (CHW per face) -> equirect [B,H,W,3] through the lookup table
__global__ void cube_gather_kernel(const uint8_t* __restrict__ faces, const uint32_t* __restrict__ lut,
                                   int res, int64_t npix, uint8_t* __restrict__ out) {
  int b = blockIdx.y;
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  uint32_t e = lut[p];
  uint8_t r = 0, g = 0, bl = 0;
  if (e != 0xFFFFFFFFu) {
    unsigned face = e >> 28, row = (e >> 14) & 0x3FFFu, col = e & 0x3FFFu;
    const size_t plane = (size_t)res * res;
    const uint8_t* f = faces + ((size_t)b * 6 + face) * 3 * plane + (size_t)row * res + col;
    r = f[0];
    g = f[plane];
    bl = f[2 * plane];
  }
  uint8_t* o = out + ((size_t)b * npix + p) * 3;
  o[0] = r;
  o[1] = g;
  o[2] = bl;
}


// ------------------------------------------------------------------------------------------
// Cube splat: one camera-space transform per (point, view); the face is the major axis of X_c and the face-local
// coordinates are the exact signed permutations inv(T_face [Rz180]) of the reference's CUBEMAP_TRANSFORMS
// (reproject_vggt_open3d_utils.py:29-36,619-622,652-658), so six 90-degree pinhole renders cost one mat-vec.
//   front (x, y, z) | right (-z, y, x) | back (-x, y, -z) | left (z, y, -x) | top (-x, -z, -y) | bottom (-x, z, y)
// Specification (mirrored by oracle_splat_keys_cube): X_c by fmaf chains; face = argmax(|x|,|y|,|z|) with ties
// resolved x over y over z... exactly as coded below; keep z' > z_near; u = fmaf(f, x'/z', c) etc. as splat_one.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cube_splat_one(const float* __restrict__ m, float x, float y, float z, float focal, float c,
                                               float z_near, int res, unsigned long long* __restrict__ zview, unsigned low,
                                               int pretest) {  // low = key bits 0..31: point index, or colour | index mod 256
  const float xc = fmaf(m[0], x, fmaf(m[1], y, fmaf(m[2], z, m[3])));
  const float yc = fmaf(m[4], x, fmaf(m[5], y, fmaf(m[6], z, m[7])));
  const float zc = fmaf(m[8], x, fmaf(m[9], y, fmaf(m[10], z, m[11])));
  const float ax = fabsf(xc), ay = fabsf(yc), az = fabsf(zc);
  int face;
  float fx, fy, fz;
  if (az >= ax && az >= ay) {
    if (zc > 0.f) { face = 0; fx = xc; fy = yc; fz = zc; }        // front
    else          { face = 2; fx = -xc; fy = yc; fz = -zc; }      // back
  } else if (ax >= ay) {
    if (xc > 0.f) { face = 1; fx = -zc; fy = yc; fz = xc; }       // right
    else          { face = 3; fx = zc; fy = yc; fz = -xc; }       // left
  } else {
    if (yc > 0.f) { face = 5; fx = -xc; fy = zc; fz = yc; }       // bottom
    else          { face = 4; fx = -xc; fy = -zc; fz = -yc; }     // top
  }
  if (!(fz > z_near)) return;
  const float u = fmaf(focal, __fdiv_rn(fx, fz), c);
  const float v = fmaf(focal, __fdiv_rn(fy, fz), c);
  const float fres = (float)res;
  if (!(u >= 0.f && u < fres && v >= 0.f && v < fres)) return;
  const int px = (int)floorf(u), py = (int)floorf(v);
  const unsigned long long key = ((unsigned long long)__float_as_uint(fz) << 32) | low;
  unsigned long long* cell = zview + ((size_t)face * res + py) * res + px;
  if (!pretest || key < ld_relaxed_u64(cell)) atomicMin(cell, key);
}

int g_splat_ctas_per_sm = 0;  // 0 = EVW_SPLAT_CTAS_PER_SM or the default
int splat_ctas_per_sm() {
  if (g_splat_ctas_per_sm == 0) {
    const char* e = getenv("EVW_SPLAT_CTAS_PER_SM");
    const int v = e ? atoi(e) : 0;
    g_splat_ctas_per_sm = (v >= 1 && v <= 8) ? v : 8;
  }
  return g_splat_ctas_per_sm;
}

// v2 of the cube splat: P points per thread per trip (P independent 16-byte loads in flight) and each view's matrix is
// fetched from shared memory once per trip (3 x LDS.128) for all P points.  Same arithmetic, same keys.
template <int G, int P>
__global__ void __launch_bounds__(256)
cube_splat2_kernel(const float4* __restrict__ pts, int64_t n_cap, const long long* __restrict__ n_dev,
                   const float* __restrict__ w2c /*[G,12]*/, int res, float focal, float z_near, int pretest, int color_key,
                   unsigned long long* __restrict__ zbuf /*[G,6,res,res]*/) {
  __shared__ __align__(16) float s_m[G * 12];
  for (int i = threadIdx.x; i < G * 12; i += blockDim.x) s_m[i] = w2c[i];
  __syncthreads();
  const int64_t n = n_dev ? min((int64_t)*n_dev, n_cap) : n_cap;
  const float c = 0.5f * (float)res;
  const size_t view_sz = (size_t)6 * res * res;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * P) {
    float4 p[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const int64_t i = i0 + k * stride;
      p[k] = ld_stream_f4(pts + (i < n ? i : i0));
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float m[12];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 row = *reinterpret_cast<const float4*>(s_m + g * 12 + q * 4);
        m[q * 4 + 0] = row.x; m[q * 4 + 1] = row.y; m[q * 4 + 2] = row.z; m[q * 4 + 3] = row.w;
      }
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int64_t i = i0 + k * stride;
        // colour-key mode (EVW_SPLAT_COLOR_KEYS): the low key word is colour << 8 | index mod 256, so the resolve reads the
        // colour out of the key instead of gathering it from the cloud
        const unsigned low = color_key ? (((__float_as_uint(p[k].w) & 0xFFFFFFu) << 8) | ((unsigned)i & 0xFFu)) : (unsigned)i;
        if (i < n) cube_splat_one(m, p[k].x, p[k].y, p[k].z, focal, c, z_near, res, zbuf + (size_t)g * view_sz, low, pretest);
      }
    }
  }
}

// v2 of the resolve: the 4 x G dependent gather chains of a thread (lookup entry -> z-buffer key -> point colour) are
// issued level by level — all keys, then all colours — so 16 loads are in flight instead of one (v1 was
// long-scoreboard bound: 48 of 54 stall cycles per issue).
template <int G>
__global__ void __launch_bounds__(256)
resolve_multi2_kernel(const unsigned long long* __restrict__ zbuf /*[G,6,res,res]*/, const float4* __restrict__ pts,
                      const uint32_t* __restrict__ lut, int res, int64_t npix, int g_count, int color_key,
                      uint8_t* __restrict__ out /*[G,npix,3]*/) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t p0 = q * 4;
  if (p0 >= npix) return;  // npix % 4 == 0 (checked by the caller): a quad is never ragged
  const uint4 e4 = *reinterpret_cast<const uint4*>(lut + p0);
  const uint32_t e[4] = {e4.x, e4.y, e4.z, e4.w};
  const size_t view_sz = (size_t)6 * res * res;
  uint32_t cell[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned face = e[j] >> 28, row = (e[j] >> 14) & 0x3FFFu, col = e[j] & 0x3FFFu;
    cell[j] = (e[j] != 0xFFFFFFFFu) ? (face * res + row) * res + col : 0u;
  }
  unsigned long long key[G][4];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j) key[g][j] = (g < g_count) ? zbuf[(size_t)g * view_sz + cell[j]] : kEmptyKey;
  uint32_t rgb[G][4];
  if (color_key) {  // the colour travels in bits 8..31 of the key: no gather from the cloud
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool hit = key[g][j] != kEmptyKey && e[j] != 0xFFFFFFFFu;
        rgb[g][j] = hit ? ((unsigned)(key[g][j] >> 8) & 0xFFFFFFu) : 0u;
      }
  } else {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool hit = key[g][j] != kEmptyKey && e[j] != 0xFFFFFFFFu;
        const unsigned idx = hit ? (unsigned)(key[g][j] & 0xFFFFFFFFull) : 0u;
        const uint32_t w = __float_as_uint(__ldg(&pts[idx].w));
        rgb[g][j] = hit ? (w & 0xFFFFFFu) : 0u;
      }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (g >= g_count) break;
    uint32_t* o = reinterpret_cast<uint32_t*>(out + ((size_t)g * npix + p0) * 3);  // 12-byte aligned quad
    o[0] = rgb[g][0] | (rgb[g][1] << 24);
    o[1] = (rgb[g][1] >> 8) | (rgb[g][2] << 16);
    o[2] = (rgb[g][2] >> 16) | (rgb[g][3] << 8);
  }
}

template <int G>
int launch_cube_pass(const float4* pts, int64_t n_cap, const long long* n_dev, const float* w2c, int res, float focal,
                     float z_near, int flags, unsigned long long* zbuf, const uint32_t* lut, int64_t npix, int g_count,
                     uint8_t* out, cudaStream_t st, int what = 3 /* bit 0: splat, bit 1: resolve */) {
  const int pretest = flags & EVW_SPLAT_PRETEST;
  const int color_key = (flags & EVW_SPLAT_COLOR_KEYS) ? 1 : 0;
  if (n_cap > 0 && (what & 1)) {
    constexpr int P = 2;
    int64_t want = (n_cap + 256 * P - 1) / (256 * P);
    // resident CTAs per SM: < 8 leaves room for the neighbouring pass's resolve / clear to co-run (two-stream pipeline)
    int64_t cap = (int64_t)evw::sm_count() * ((flags & EVW_SPLAT_OVERLAP) ? splat_ctas_per_sm() : 8);
    cube_splat2_kernel<G, P><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(pts, n_cap, n_dev, w2c, res, focal, z_near,
                                                                                    pretest, color_key, zbuf);
  }
  if (!(what & 2)) {
    EVW_LAUNCH_CHECK();
    return 0;
  }
  const int64_t quads = (npix + 3) / 4;
  resolve_multi2_kernel<G><<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(zbuf, pts, lut, res, npix, g_count, color_key, out);
  EVW_LAUNCH_CHECK();  // a failed launch in an early pass is reported by that pass, not by the last one
  return 0;
}

template <int G>
int launch_splat(const float4* pts, int64_t n_cap, const long long* n_dev, const float* w2c, int res,
                 float focal, float z_near, unsigned long long* zbuf, cudaStream_t st) {
  int64_t want = (n_cap + 255) / 256;
  int grid = (int)(want < (int64_t)evw::sm_count() * 8 ? (want > 0 ? want : 1)
                                                         : (int64_t)evw::sm_count() * 8);
  splat_kernel<G><<<grid, 256, 0, st>>>(pts, n_cap, n_dev, w2c, res, focal, z_near, zbuf);
  return 0;
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" int evw_plucker(const float* ray, const float* c2w, float* out, int T, int H, int W,
                           void* stream) {
  EVW_CHECK_ARG(ray && c2w && out, "evw_plucker: null pointer");
  EVW_CHECK_ARG(T > 0 && H > 0 && W > 0 && T <= 4096, "evw_plucker: bad shape T=%d H=%d W=%d", T, H, W);
  int HW = H * W;
  plucker_kernel<<<(HW + 255) / 256, 256, T * 12 * sizeof(float), (cudaStream_t)stream>>>(
      ray, c2w, out, T, HW);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_equi2pers_u8(const uint8_t* equi, const float* pix2dir, uint8_t* out, int B, int C,
                                int He, int We, int Hp, int Wp, void* stream) {
  EVW_CHECK_ARG(equi && pix2dir && out, "evw_equi2pers_u8: null pointer");
  EVW_CHECK_ARG(B > 0 && C > 0 && He > 0 && We > 0 && Hp > 0 && Wp > 0 && B <= 65535,
                "evw_equi2pers_u8: bad shape");
  dim3 blk(32, 8), grd((Wp + 31) / 32, (Hp + 7) / 8, B);
  equi2pers_kernel<<<grd, blk, 0, (cudaStream_t)stream>>>(equi, pix2dir, out, C, He, We, Hp, Wp);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_equi2pers_table(const float* pix2dir0, float* table, int He, int We, int Hp, int Wp, void* stream) {
  EVW_CHECK_ARG(pix2dir0 && table, "evw_equi2pers_table: null pointer");
  EVW_CHECK_ARG(He > 0 && We > 0 && Hp > 0 && Wp > 0 && ((uintptr_t)table & 7) == 0, "evw_equi2pers_table: bad shape / alignment");
  dim3 blk(32, 8), grd((Wp + 31) / 32, (Hp + 7) / 8);
  equi2pers_table_kernel<<<grd, blk, 0, (cudaStream_t)stream>>>(pix2dir0, reinterpret_cast<float2*>(table), He, We, Hp, Wp);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_equi2pers_yaw_u8(const uint8_t* equi, const float* table, const float* shift_px, uint8_t* out, int B, int C,
                                    int He, int We, int Hp, int Wp, void* stream) {
  EVW_CHECK_ARG(equi && table && shift_px && out, "evw_equi2pers_yaw_u8: null pointer");
  EVW_CHECK_ARG(B > 0 && C > 0 && He > 0 && We > 0 && Hp > 0 && Wp > 0 && B <= 65535, "evw_equi2pers_yaw_u8: bad shape");
  dim3 blk(32, 8), grd((Wp + 31) / 32, (Hp + 7) / 8, B);
  equi2pers_yaw_kernel<<<grd, blk, 0, (cudaStream_t)stream>>>(equi, reinterpret_cast<const float2*>(table), shift_px, out, C, He,
                                                              We, Hp, Wp);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_lift_depth(const float* depth, const float* extr, const float* intr,
                              double* out_f64, float* out_f32, int S, int H, int W, void* stream) {
  EVW_CHECK_ARG(depth && extr && intr, "evw_lift_depth: null pointer");
  EVW_CHECK_ARG(out_f64 || out_f32, "evw_lift_depth: no output buffer");
  EVW_CHECK_ARG(S > 0 && H > 0 && W > 0 && S <= 65535, "evw_lift_depth: bad shape");
  dim3 grd((H * W + 255) / 256, S);
  lift_kernel<<<grd, 256, 0, (cudaStream_t)stream>>>(depth, extr, intr, out_f64, out_f32, H, W);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_lift_pack_points(const float* depth, const float* extr, const float* intr, const float* images_f32,
                                    float* out_pts4, int S, int H, int W, void* stream) {
  EVW_CHECK_ARG(depth && extr && intr && images_f32 && out_pts4, "evw_lift_pack_points: null pointer");
  EVW_CHECK_ARG(S > 0 && H > 0 && W > 0 && S <= 65535, "evw_lift_pack_points: bad shape");
  EVW_CHECK_ARG(((uintptr_t)out_pts4 & 15) == 0, "evw_lift_pack_points: out_pts4 must be 16-byte aligned");
  dim3 grd((H * W + 255) / 256, S);
  lift_pack_kernel<<<grd, 256, 0, (cudaStream_t)stream>>>(depth, extr, intr, images_f32, reinterpret_cast<float4*>(out_pts4), H, W);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_pack_points(const double* xyz_f64, const float* xyz_f32, const uint8_t* rgb_u8,
                               const float* images_f32, int S, int HW, float* out_pts4, int64_t N,
                               void* stream) {
  EVW_CHECK_ARG((xyz_f64 != nullptr) != (xyz_f32 != nullptr), "evw_pack_points: exactly one xyz source");
  EVW_CHECK_ARG((rgb_u8 != nullptr) != (images_f32 != nullptr), "evw_pack_points: exactly one colour source");
  EVW_CHECK_ARG(out_pts4 && N >= 0, "evw_pack_points: bad output");
  if (images_f32) EVW_CHECK_ARG((int64_t)S * HW == N, "evw_pack_points: N != S*HW");
  if (N == 0) return EVW_OK;
  pack_points_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      xyz_f64, xyz_f32, rgb_u8, images_f32, HW, reinterpret_cast<float4*>(out_pts4), N);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

static int64_t compact_blocks(int64_t n) { return (n + kCompactTile - 1) / kCompactTile; }

extern "C" int64_t evw_conf_select_workspace(int64_t n) {
  int64_t nb = compact_blocks(n > 0 ? n : 1);
  return 256 /*state*/ + 3 * 2048 * 4 /*hist*/ + evw::align_up(nb * 4, 256) + evw::align_up(nb * 8, 256);
}

extern "C" int evw_conf_select(const float* conf, const float* pts4_in, int64_t n, int64_t k_lo,
                               int64_t k_hi, float gamma, int use_threshold, float* pts4_out,
                               int64_t* keep_idx, int64_t* out_count, float* out_thr, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  EVW_CHECK_ARG(conf && out_count && workspace, "evw_conf_select: null pointer");
  EVW_CHECK_ARG(n > 0 && n < (1ll << 32), "evw_conf_select: n=%lld out of range", (long long)n);
  EVW_CHECK_ARG((pts4_in == nullptr) == (pts4_out == nullptr), "evw_conf_select: pts in/out mismatch");
  EVW_CHECK_ARG(0 <= k_lo && k_lo <= k_hi && k_hi < n, "evw_conf_select: bad ranks");
  if (workspace_bytes < evw_conf_select_workspace(n)) {
    evw::set_error("evw_conf_select: workspace %lld < %lld", (long long)workspace_bytes,
                   (long long)evw_conf_select_workspace(n));
    return EVW_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  SelectState* state = (SelectState*)ws;
  unsigned* hist = (unsigned*)(ws + 256);
  int64_t nb = compact_blocks(n);
  unsigned* bcount = (unsigned*)(ws + 256 + 3 * 2048 * 4);
  long long* boff = (long long*)((char*)bcount + evw::align_up(nb * 4, 256));
  int grid = evw::sm_count() * 4;
  // state + the three histograms: one memset node instead of an init kernel; next_key starts at "none"
  EVW_CUDA(cudaMemsetAsync(ws, 0, 256 + 3 * 2048 * 4, st));
  EVW_CUDA(cudaMemsetAsync(&state->next_key, 0xFF, sizeof(unsigned), st));
  if (use_threshold) {  // 3 histogram passes (each resolves the previous digits in its prologue) + 1 block for the threshold
    select_hist_kernel<0><<<grid, 512, 0, st>>>(conf, n, state, hist, (unsigned long long)k_lo);
    select_hist_kernel<1><<<grid, 512, 0, st>>>(conf, n, state, hist, (unsigned long long)k_lo);
    select_hist_kernel<2><<<grid, 512, 0, st>>>(conf, n, state, hist, (unsigned long long)k_lo);
  }
  select_thr_kernel<<<1, 512, 0, st>>>(state, hist, (long long)k_lo, (long long)k_hi, gamma, use_threshold, out_thr);
  compact_count_kernel<<<(unsigned)nb, kCompactThreads, 0, st>>>(conf, n, state, bcount);
  compact_scan_kernel<<<1, 1024, 0, st>>>(bcount, (int)nb, boff, (long long*)out_count);
  compact_scatter_kernel<<<(unsigned)nb, kCompactThreads, 0, st>>>(
      conf, reinterpret_cast<const float4*>(pts4_in), n, state, boff,
      reinterpret_cast<float4*>(pts4_out), (long long*)keep_idx);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int64_t evw_splat_workspace(int views_per_pass, int face_res) {
  return (int64_t)views_per_pass * 6 * face_res * face_res * 8;
}

static int splat_pass(int G, const float4* pts, int64_t n_cap, const long long* n_dev,
                      const float* w2c, int res, float focal, float z_near,
                      unsigned long long* zbuf, cudaStream_t st) {
  switch (G) {
    case 1: return launch_splat<1>(pts, n_cap, n_dev, w2c, res, focal, z_near, zbuf, st);
    case 2: return launch_splat<2>(pts, n_cap, n_dev, w2c, res, focal, z_near, zbuf, st);
    case 3: return launch_splat<3>(pts, n_cap, n_dev, w2c, res, focal, z_near, zbuf, st);
    case 4: return launch_splat<4>(pts, n_cap, n_dev, w2c, res, focal, z_near, zbuf, st);
    case 6: return launch_splat<6>(pts, n_cap, n_dev, w2c, res, focal, z_near, zbuf, st);
    case 8: return launch_splat<8>(pts, n_cap, n_dev, w2c, res, focal, z_near, zbuf, st);
    default: return -1;
  }
}

extern "C" int evw_splat_cubemap_equirect(const float* pts4, int64_t n_cap, const int64_t* n_dev,
                                          const float* w2c, int V, int face_res, float focal,
                                          float z_near, const uint32_t* lut, int outH, int outW,
                                          uint8_t* out, void* zbuf_workspace, int64_t workspace_bytes,
                                          int views_per_pass, void* stream) {
  EVW_CHECK_ARG((pts4 || n_cap == 0) && w2c && lut && out && zbuf_workspace,
                "evw_splat_cubemap_equirect: null pointer");
  EVW_CHECK_ARG(n_cap >= 0 && n_cap < (1ll << 32), "evw_splat_cubemap_equirect: n out of range");
  EVW_CHECK_ARG(V > 0 && face_res > 0 && face_res <= 16383 && outH > 0 && outW > 0,
                "evw_splat_cubemap_equirect: bad shape");
  int G = views_per_pass;
  EVW_CHECK_ARG(G == 1 || G == 2 || G == 3 || G == 4 || G == 6 || G == 8,
                "evw_splat_cubemap_equirect: views_per_pass must be one of 1,2,3,4,6,8");
  if (workspace_bytes < evw_splat_workspace(G, face_res)) {
    evw::set_error("evw_splat_cubemap_equirect: workspace too small");
    return EVW_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* zbuf = (unsigned long long*)zbuf_workspace;
  const size_t view_cells = (size_t)6 * face_res * face_res;
  const int64_t npix = (int64_t)outH * outW;
  for (int v0 = 0; v0 < V;) {
    int g = (V - v0 < G) ? (V - v0) : G;
    // a short tail uses the largest supported group size <= what is left
    while (!(g == 1 || g == 2 || g == 3 || g == 4 || g == 6 || g == 8)) --g;
    EVW_CUDA(cudaMemsetAsync(zbuf, 0xFF, (size_t)g * view_cells * 8, st));
    if (n_cap > 0) {
      if (splat_pass(g, reinterpret_cast<const float4*>(pts4), n_cap, (const long long*)n_dev,
                     w2c + (size_t)v0 * 72, face_res, focal, z_near, zbuf, st) != 0) {
        evw::set_error("evw_splat_cubemap_equirect: internal group size");
        return EVW_ERR_INVALID;
      }
    }
    for (int j = 0; j < g; ++j) {
      int64_t quads = (npix + 3) / 4;
      resolve_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, st>>>(
          zbuf + (size_t)j * view_cells, reinterpret_cast<const float4*>(pts4), lut, face_res, npix,
          out + (size_t)(v0 + j) * npix * 3);
    }
    v0 += g;
  }
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_splat_faces_debug(const float* pts4, int64_t n, const float* w2c, int V,
                                     int face_res, float focal, float z_near, int64_t* win_idx,
                                     void* zbuf_workspace, int64_t workspace_bytes, void* stream) {
  EVW_CHECK_ARG(pts4 && w2c && win_idx && zbuf_workspace, "evw_splat_faces_debug: null pointer");
  EVW_CHECK_ARG(n >= 0 && n < (1ll << 32) && V > 0 && face_res > 0, "evw_splat_faces_debug: bad shape");
  if (workspace_bytes < evw_splat_workspace(1, face_res)) {
    evw::set_error("evw_splat_faces_debug: workspace too small");
    return EVW_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* zbuf = (unsigned long long*)zbuf_workspace;
  const int64_t view_cells = (int64_t)6 * face_res * face_res;
  for (int v = 0; v < V; ++v) {
    EVW_CUDA(cudaMemsetAsync(zbuf, 0xFF, (size_t)view_cells * 8, st));
    if (n > 0)
      launch_splat<1>(reinterpret_cast<const float4*>(pts4), n, nullptr, w2c + (size_t)v * 72,
                      face_res, focal, z_near, zbuf, st);
    zbuf_to_index_kernel<<<(unsigned)((view_cells + 255) / 256), 256, 0, st>>>(
        zbuf, view_cells, (long long*)win_idx + (size_t)v * view_cells);
  }
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_splat_faces_u8(const float* pts4, int64_t n, const float* w2c, int V, int face_res,
                                  float focal, float z_near, uint8_t* faces_out, void* zbuf_workspace,
                                  int64_t workspace_bytes, void* stream) {
  EVW_CHECK_ARG(pts4 && w2c && faces_out && zbuf_workspace, "evw_splat_faces_u8: null pointer");
  EVW_CHECK_ARG(n >= 0 && n < (1ll << 32) && V > 0 && face_res > 0, "evw_splat_faces_u8: bad shape");
  if (workspace_bytes < evw_splat_workspace(1, face_res)) {
    evw::set_error("evw_splat_faces_u8: workspace too small");
    return EVW_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* zbuf = (unsigned long long*)zbuf_workspace;
  const int64_t view_cells = (int64_t)6 * face_res * face_res;
  for (int v = 0; v < V; ++v) {
    EVW_CUDA(cudaMemsetAsync(zbuf, 0xFF, (size_t)view_cells * 8, st));
    if (n > 0)
      launch_splat<1>(reinterpret_cast<const float4*>(pts4), n, nullptr, w2c + (size_t)v * 72,
                      face_res, focal, z_near, zbuf, st);
    zbuf_to_rgb_kernel<<<(unsigned)((view_cells + 255) / 256), 256, 0, st>>>(
        zbuf, reinterpret_cast<const float4*>(pts4), view_cells, faces_out + (size_t)v * view_cells * 3);
  }
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_cube_to_equirect_u8(const uint8_t* faces, const uint32_t* lut, int B, int face_res,
                                       int outH, int outW, uint8_t* out, void* stream) {
  EVW_CHECK_ARG(faces && lut && out, "evw_cube_to_equirect_u8: null pointer");
  EVW_CHECK_ARG(B > 0 && B <= 65535 && face_res > 0 && face_res <= 16383 && outH > 0 && outW > 0,
                "evw_cube_to_equirect_u8: bad shape");
  int64_t npix = (int64_t)outH * outW;
  dim3 grd((unsigned)((npix + 255) / 256), B);
  cube_gather_kernel<<<grd, 256, 0, (cudaStream_t)stream>>>(faces, lut, face_res, npix, out);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" void evw_set_splat_ctas_per_sm(int v) { g_splat_ctas_per_sm = (v >= 1 && v <= 8) ? v : 0; }

extern "C" int64_t evw_splat_workspace_flags(int views_per_pass, int face_res, int flags) {
  return evw_splat_workspace(views_per_pass, face_res) * ((flags & EVW_SPLAT_OVERLAP) ? 2 : 1);
}

namespace {
// Two internal streams for the pass pipeline (fork/join around the caller's stream with events; capturable).
constexpr int kMaxPassEvents = 64;
struct SplatStreams {
  std::mutex mu;  // the two streams + events are per device: one pass pipeline at a time per device (host threads serialise)
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  cudaEvent_t cleared[kMaxPassEvents] = {}, splatted[kMaxPassEvents] = {};
  int device = -1;
};
int splat_streams(SplatStreams** out) {
  static SplatStreams per_dev[16];
  static std::mutex init_mu;
  std::lock_guard<std::mutex> init_guard(init_mu);
  int dev = 0;
  EVW_CUDA(cudaGetDevice(&dev));
  EVW_CHECK_ARG(dev >= 0 && dev < 16, "evw_splat_cube_equirect: device index %d out of range", dev);
  SplatStreams& s = per_dev[dev];
  if (s.device != dev) {
    for (int i = 0; i < 2; ++i) {
      EVW_CUDA(cudaStreamCreateWithFlags(&s.s[i], cudaStreamNonBlocking));
      EVW_CUDA(cudaEventCreateWithFlags(&s.join[i], cudaEventDisableTiming));
    }
    EVW_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    for (int i = 0; i < kMaxPassEvents; ++i) {
      EVW_CUDA(cudaEventCreateWithFlags(&s.cleared[i], cudaEventDisableTiming));
      EVW_CUDA(cudaEventCreateWithFlags(&s.splatted[i], cudaEventDisableTiming));
    }
    s.device = dev;
  }
  *out = &s;
  return EVW_OK;
}

int cube_pass_dispatch(int G, const float4* p4, int64_t n_cap, const long long* n_dev, const float* m, int res, float focal,
                       float z_near, int flags, unsigned long long* zbuf, const uint32_t* lut, int64_t npix, int g,
                       uint8_t* o, cudaStream_t st, int what = 3) {
  const size_t view_cells = (size_t)6 * res * res;
  if (g < G) {  // short tail: single-view passes
    for (int j = 0; j < g; ++j) {
      const int rc = launch_cube_pass<1>(p4, n_cap, n_dev, m + (size_t)j * 12, res, focal, z_near, flags,
                                         zbuf + (size_t)j * view_cells, lut, npix, 1, o + (size_t)j * npix * 3, st, what);
      if (rc) return rc;
    }
    return 0;
  }
  switch (G) {
    case 1: return launch_cube_pass<1>(p4, n_cap, n_dev, m, res, focal, z_near, flags, zbuf, lut, npix, g, o, st, what);
    case 2: return launch_cube_pass<2>(p4, n_cap, n_dev, m, res, focal, z_near, flags, zbuf, lut, npix, g, o, st, what);
    case 4: return launch_cube_pass<4>(p4, n_cap, n_dev, m, res, focal, z_near, flags, zbuf, lut, npix, g, o, st, what);
    default: return launch_cube_pass<8>(p4, n_cap, n_dev, m, res, focal, z_near, flags, zbuf, lut, npix, g, o, st, what);
  }
}
}  // namespace

extern "C" int evw_splat_cube_equirect(const float* pts4, int64_t n_cap, const int64_t* n_dev, const float* w2c_front,
                                       int V, int face_res, float focal, float z_near, const uint32_t* lut, int outH,
                                       int outW, uint8_t* out, void* zbuf_workspace, int64_t workspace_bytes,
                                       int views_per_pass, int flags, void* stream) {
  EVW_CHECK_ARG((pts4 || n_cap == 0) && w2c_front && lut && out && zbuf_workspace, "evw_splat_cube_equirect: null pointer");
  EVW_CHECK_ARG(n_cap >= 0 && n_cap < (1ll << 32), "evw_splat_cube_equirect: n out of range");
  EVW_CHECK_ARG(V > 0 && face_res > 0 && face_res <= 16383 && outH > 0 && outW > 0, "evw_splat_cube_equirect: bad shape");
  EVW_CHECK_ARG(((int64_t)outH * outW) % 4 == 0, "evw_splat_cube_equirect: outH*outW must be a multiple of 4");
  const int G = views_per_pass;
  EVW_CHECK_ARG(G == 1 || G == 2 || G == 4 || G == 8, "evw_splat_cube_equirect: views_per_pass must be 1, 2, 4 or 8");
  if (workspace_bytes < evw_splat_workspace_flags(G, face_res, flags)) {
    evw::set_error("evw_splat_cube_equirect: workspace too small");
    return EVW_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* zbuf = (unsigned long long*)zbuf_workspace;
  const size_t view_cells = (size_t)6 * face_res * face_res;
  const int64_t npix = (int64_t)outH * outW;
  const float4* p4 = reinterpret_cast<const float4*>(pts4);
  const int passes = (V + G - 1) / G;
  const bool overlap = (flags & EVW_SPLAT_OVERLAP) && passes > 1;
  const bool by_role = overlap && (flags & EVW_SPLAT_OVERLAP_BY_ROLE) && passes <= kMaxPassEvents;
  SplatStreams* ss = nullptr;
  std::unique_lock<std::mutex> guard;
  if (overlap) {
    int rc = splat_streams(&ss);
    if (rc) return rc;
    guard = std::unique_lock<std::mutex>(ss->mu);  // held until every launch of this call is enqueued
    EVW_CUDA(cudaEventRecord(ss->fork, st));
    EVW_CUDA(cudaStreamWaitEvent(ss->s[0], ss->fork, 0));
    EVW_CUDA(cudaStreamWaitEvent(ss->s[1], ss->fork, 0));
  }
  if (by_role) {
    // stream 0 runs the splats back to back (they are L2-atomic bound and gain nothing from overlapping each other);
    // stream 1 clears the z-buffer halves and resolves: resolve(p) and clear(p + 2) overlap splat(p + 1).
    cudaStream_t sa = ss->s[0], sb = ss->s[1];
    auto zb_of = [&](int pass) { return zbuf + (size_t)(pass & 1) * G * view_cells; };
    auto clear = [&](int pass) -> int {
      EVW_CUDA(cudaMemsetAsync(zb_of(pass), 0xFF, (size_t)G * view_cells * 8, sb));
      EVW_CUDA(cudaEventRecord(ss->cleared[pass], sb));
      return EVW_OK;
    };
    int rc = clear(0);
    if (rc) return rc;
    if (passes > 1 && (rc = clear(1))) return rc;
    for (int pass = 0; pass < passes; ++pass) {
      const int v0 = pass * G;
      const int g = (V - v0 < G) ? (V - v0) : G;
      const float* m = w2c_front + (size_t)v0 * 12;
      uint8_t* o = out + (size_t)v0 * npix * 3;
      EVW_CUDA(cudaStreamWaitEvent(sa, ss->cleared[pass], 0));
      if ((rc = cube_pass_dispatch(G, p4, n_cap, (const long long*)n_dev, m, face_res, focal, z_near, flags, zb_of(pass), lut, npix, g, o, sa, 1)))
        return rc;
      EVW_CUDA(cudaEventRecord(ss->splatted[pass], sa));
      EVW_CUDA(cudaStreamWaitEvent(sb, ss->splatted[pass], 0));
      if ((rc = cube_pass_dispatch(G, p4, n_cap, (const long long*)n_dev, m, face_res, focal, z_near, flags, zb_of(pass), lut, npix, g, o, sb, 2)))
        return rc;
      if (pass + 2 < passes && (rc = clear(pass + 2))) return rc;
    }
  } else {
    // pass p runs clear -> splat -> resolve on internal stream p % 2 (or on the caller's stream) with its own half of the workspace
    for (int v0 = 0, pass = 0; v0 < V; v0 += G, ++pass) {
      const int g = (V - v0 < G) ? (V - v0) : G;
      cudaStream_t ps = overlap ? ss->s[pass & 1] : st;
      unsigned long long* zb = zbuf + (overlap ? (size_t)(pass & 1) * G * view_cells : 0);
      EVW_CUDA(cudaMemsetAsync(zb, 0xFF, (size_t)G * view_cells * 8, ps));
      const int rc = cube_pass_dispatch(G, p4, n_cap, (const long long*)n_dev, w2c_front + (size_t)v0 * 12, face_res, focal,
                                        z_near, flags, zb, lut, npix, g, out + (size_t)v0 * npix * 3, ps);
      if (rc) return rc;
    }
  }
  if (overlap) {
    for (int i = 0; i < 2; ++i) {
      EVW_CUDA(cudaEventRecord(ss->join[i], ss->s[i]));
      EVW_CUDA(cudaStreamWaitEvent(st, ss->join[i], 0));
    }
  }
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

extern "C" int evw_splat_cube_faces_debug(const float* pts4, int64_t n, const float* w2c_front, int V, int face_res,
                                          float focal, float z_near, int64_t* win_idx, void* zbuf_workspace,
                                          int64_t workspace_bytes, void* stream) {
  EVW_CHECK_ARG(pts4 && w2c_front && win_idx && zbuf_workspace, "evw_splat_cube_faces_debug: null pointer");
  EVW_CHECK_ARG(n >= 0 && n < (1ll << 32) && V > 0 && face_res > 0, "evw_splat_cube_faces_debug: bad shape");
  if (workspace_bytes < evw_splat_workspace(1, face_res)) {
    evw::set_error("evw_splat_cube_faces_debug: workspace too small");
    return EVW_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* zbuf = (unsigned long long*)zbuf_workspace;
  const int64_t view_cells = (int64_t)6 * face_res * face_res;
  for (int v = 0; v < V; ++v) {
    EVW_CUDA(cudaMemsetAsync(zbuf, 0xFF, (size_t)view_cells * 8, st));
    if (n > 0) {
      int64_t want = (n + 511) / 512, cap = (int64_t)evw::sm_count() * 8;
      cube_splat2_kernel<1, 2><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(
          reinterpret_cast<const float4*>(pts4), n, nullptr, w2c_front + (size_t)v * 12, face_res, focal, z_near, 1, 0, zbuf);
    }
    zbuf_to_index_kernel<<<(unsigned)((view_cells + 255) / 256), 256, 0, st>>>(zbuf, view_cells,
                                                                               (long long*)win_idx + (size_t)v * view_cells);
  }
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}
