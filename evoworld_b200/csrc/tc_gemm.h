// Host-side interface of the tcgen05 implicit-GEMM (tc_gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace evw {

constexpr int kMaxTaps = 20;  // 3x3 conv with split-precision activations: 9 head taps + 9 tail taps

// out[row, n] = s0 * (acc + bias[n]) + rowvec[(((row + rv_row0) / rv_div) % rv_mod) * rv_ld + n] + s1 * res1[row, n] + s2 * res2[row, n]
// geglu: acc columns come in interleaved [16 value | 16 gate] groups, out has N/2 columns:
//        out = (value + bias_v) * gelu(gate + bias_g), then the same affine tail.
struct GemmEpilogue {
  void* out = nullptr;
  int out_fp16 = 1;
  void* out_lo = nullptr;  // optional (fp16 output only): fp16 tail  half(v - float(half(v)))  of every stored value, same
                           // layout as out — the consumer of a split-precision operand reads head and tail as two sources
  const float* bias = nullptr;
  const float* rowvec = nullptr;
  long long rv_div = 1, rv_mod = 1;
  long long rv_row0 = 0;  // row offset added before the broadcast index (a GEMM over a row panel of a larger tensor)
  int rv_ld = 0;  // row stride of rowvec in floats (0 = output width)
  const void* res1 = nullptr;
  int res1_fp16 = 0;
  float s1 = 1.f;
  const float* res2 = nullptr;
  float s2 = 1.f;
  float s0 = 1.f;
  int geglu = 0;
  int act = 0;  // 1: exact GELU applied to s0 * acc + bias (epilogues without row vector / residuals only): the MLP fc1 of the
                // ViT-style blocks (VGGT, CLIP) writes its fp16 activation directly instead of an fp32 tensor + an activation pass
  // GroupNorm(32) statistics of the OUTPUT, accumulated by the epilogue so that the consumer GroupNorm skips its statistics
  // pass over the tensor: gn_stats double [instances, 32, 2] (sum, sum of squares; cleared by gemm_launch), instance of a
  // row = row / rows_per_inst, group of a column = n / gn_cg (gn_cg = N / 32).  Set through gemm_enable_gn_stats().
  double* gn_stats = nullptr;
  int gn_cg = 0;
};

// Activations are fp16 [B, T, Y, X, C] (channels last); weights fp16 [N, K_total] with
// K_total = sum over taps of the tap's channel count, tap-major.  Rows of the output are
// (b, t, y, x) flattened.  A tap reads the activation shifted by (dx, dy, dt); reads outside
// [0,X) x [0,Y) x [0,T) are zero (TMA out-of-bounds fill) — the convolution padding.
struct GemmProblem {
  const void* a0 = nullptr;  // source 0, C0 channels
  const void* a1 = nullptr;  // optional source 1 (1x1 shortcut input), C1 channels
  const void* w = nullptr;
  int B = 1, T = 1, Y = 1, X = 1, C0 = 0, C1 = 0, N = 0;
  int Tmap = 0;  // extent of the T dimension of the tensor map when it differs from T (0 = T); used by the
                 // stride-2 convolution, whose taps select one of 4 phase images through the T coordinate
  long long K_total = 0;
  int num_taps = 1;
  int8_t tap_dx[kMaxTaps] = {0}, tap_dy[kMaxTaps] = {0}, tap_dt[kMaxTaps] = {0}, tap_src[kMaxTaps] = {0};
  int block_n = 0;  // 0 = choose
  // Strided output view (TMA-store epilogue only): the output of pixel (x, y) goes to pixel (out_sx x + out_ox,
  // out_sy y + out_oy) of an image of out_sx X by out_sy Y pixels — one phase of a fused nearest-x2-upsample + 3x3 conv
  int out_sx = 1, out_sy = 1, out_ox = 0, out_oy = 0;
  GemmEpilogue ep;
};

struct alignas(64) GemmOp {
  alignas(64) unsigned char tmap_a0[128];
  alignas(64) unsigned char tmap_a1[128];
  alignas(64) unsigned char tmap_b[128];
  alignas(64) unsigned char tmap_bh[128];  // half-height weight box (CTA-pair mode: each CTA loads its own half)
  alignas(64) unsigned char tmap_r[128];   // fp32 residual (res1) as a 5-D map with 32-column boxes (TMA-staged residual)
  alignas(64) unsigned char tmap_o[128];   // output as a 5-D map with 32-row x 16-column boxes (TMA-store epilogue)
  unsigned char params[448];
  int grid = 0;
  int cluster = 0;  // 1: launch as clusters of two CTAs sharing each weight tile (TMA multicast)
  int smem_bytes = 0;
  int res_tma = 0;  // 1: res1 streams through a shared-memory ring filled by TMA (fp32 residual, BLOCK_N % 32 == 0)
  int store_tma = 0;  // 1: the epilogue stages each warp's 32 x 16 outputs in shared memory and stores them with TMA
  int strided_out = 0;  // 1: the output map is a strided view (GemmProblem::out_sx ...): only the TMA-store kernels honour it
  double flops = 0;
};

int gemm_plan(GemmOp* op, const GemmProblem& pr);
// Taps and output phase of a fused nearest-x2 up-sampling + 3x3 convolution (padding 1): output pixel (2y + py, 2x + px) is a
// 2x2 convolution of the LOW-resolution input — rows {y - 1, y} for py = 0, {y, y + 1} for py = 1 (columns likewise) — with
// the 3x3 weights that fall on the same source pixel summed: 4 GEMMs of K = 4 C instead of one of K = 9 C on a 4x larger
// input.  phase = 2 py + px; weights [N, 4 C], taps ordered (dy, dx) as filled here (packed by unet.py / vae.py: upconv_weights).
void upconv2x_phase(GemmProblem& pr, int phase);
// whether a GEMM over an X x Y image with N fp32 output columns can store through a strided output view (TMA-store geometry)
bool gemm_strided_out_ok(int X, int Y, int N);
int gemm_launch(const GemmOp& op, cudaStream_t stream);
int gemm_enable_gn_stats(GemmOp* op, double* stats, long long rows_per_inst);  // 0 = enabled, 1 = not available for this GEMM
void set_gemm_gn_stats(int on);     // -1 = EVW_GEMM_GN_STATS / default on, 0 = GroupNorms keep their own statistics pass, 1 = on
void set_gemm_store_tma(int mode);  // -1 = EVW_GEMM_STORE_TMA / default (1 = wherever possible), 0 = direct stores only
void set_gemm_cluster_mode(int on);  // -1 = EVW_GEMM_CLUSTER / default, 0 = off, 1 = on (takes effect at plan time)

}  // namespace evw
