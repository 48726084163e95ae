// VAE around the denoise loop: AutoencoderKLTemporalDecoder.encode / .decode as the pipeline calls
// them (evoworld/pipeline/pipeline_evoworld.py:307-328 `_encode_vae_image`, :358-385 `decode_latents`; the class is
// diffusers 0.31 models/autoencoders/autoencoder_kl_temporal_decoder.py), planned over the same building blocks as the
// UNet: 3x3 / temporal / strided convolutions and linears on the tcgen05 implicit GEMM, GroupNorm(+SiLU) with the sums
// from the producer's epilogue, fp32 residual stream, fp16 operands.  The single-head (d = C = 512) self-attention of the
// two mid blocks runs as three GEMMs per frame (Q K^T -> fp32 scores, row softmax -> fp16, P V) — the flash kernel of the
// UNet is built for head dim 64.
//
// Data layout in HBM (caller-provided workspace): activations channels-last, rows = (frame, y, x).
#include "common.h"
#include "plan_builder.h"

#include <cmath>
#include <cstdlib>

namespace evw {
namespace {

struct VaeConfig {
  int in_channels = 3, out_channels = 3, latent = 4;
  int boc[4] = {128, 256, 512, 512};
  int layers = 2;
  int cin_pad = 64, cout_pad = 16;
  int split = 1;  // encoder conv_in in split precision ([head | tail | head] inside the channel padding, as the UNet's)
};

struct VaeArgs {
  const float* in = nullptr;  // encode: images [N,3,H,W]; decode: latents [N,4,h,w] (already divided by scaling_factor)
  float* out = nullptr;       // encode: moments [N,8,H/8,W/8]; decode: frames [N,3,8h,8w]
};

struct VaePlan : PlanCore {
  int mode = -1, N = 0, F = 0, H = 0, W = 0;  // mode 0 = encode (H, W of the image), 1 = decode (H, W of the latent)
  void* ws = nullptr;
  long long ws_bytes = 0;
};

struct Vae {
  VaeConfig cfg;
  TensorMap tensors;
  ScalarMap scalars;
  std::unique_ptr<VaePlan> plan[2];
  VaeArgs args;
};

// ------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------
// softmax over the columns of fp32 scores [rows, cols] -> fp16 probabilities; one block per row.  A row (36 KiB at
// S = 9216) is read three times, the second and third time from L1 / L2.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, __half* __restrict__ p, int cols) {
  const float* row = s + (long long)blockIdx.x * cols;
  __half* out = p + (long long)blockIdx.x * cols;
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    sum += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x * 4; c < cols; c += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + c);
    __half2 a = __floats2half2_rn(__expf(v.x - m) * inv, __expf(v.y - m) * inv);
    __half2 b = __floats2half2_rn(__expf(v.z - m) * inv, __expf(v.w - m) * inv);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(out + c) = r;
  }
}

int softmax_rows(const float* s, __half* p, long long rows, int cols, cudaStream_t st) {
  EVW_CHECK_ARG(cols % 4 == 0 && rows > 0 && rows < (1ll << 31), "softmax_rows: cols=%d rows=%lld", cols, rows);
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(s, p, cols);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

// time_conv_out (Conv3d(3, 3, (3,1,1), padding (1,0,0)) over the frames of each video) fused with the channels-last ->
// NCHW conversion of the decoder output: y fp32 [B*T, HW, Npad] (first Co columns) -> out fp32 [B*T, Co, HW]
__global__ void time_conv_out_kernel(const float* __restrict__ y, const float* __restrict__ w /*[Co, Co, 3] (o, i, kt)*/,
                                     const float* __restrict__ bias, int T, long long HW, int Co, int Npad, long long frames,
                                     float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= frames * HW) return;
  const long long f = idx / HW, px = idx - f * HW;
  const int t = (int)(f % T);
  float acc[4];
  for (int o = 0; o < Co; ++o) acc[o] = bias[o];
  for (int kt = 0; kt < 3; ++kt) {
    const int tt = t + kt - 1;
    if (tt < 0 || tt >= T) continue;  // zero padding at the ends of the video
    const float4 v = *reinterpret_cast<const float4*>(y + ((f + kt - 1) * HW + px) * Npad);
    const float in[4] = {v.x, v.y, v.z, v.w};
    for (int o = 0; o < Co; ++o)
      for (int i = 0; i < Co; ++i) acc[o] = fmaf(w[(o * Co + i) * 3 + kt], in[i], acc[o]);
  }
  for (int o = 0; o < Co; ++o) out[(f * Co + o) * HW + px] = acc[o];
}

int time_conv_out(const float* y, const float* w, const float* bias, int T, long long HW, int Co, int Npad, long long frames,
                  float* out, cudaStream_t st) {
  EVW_CHECK_ARG(Co >= 1 && Co <= 4 && Npad % 4 == 0 && Npad >= 4, "time_conv_out: Co=%d Npad=%d", Co, Npad);
  const long long n = frames * HW;
  time_conv_out_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, w, bias, T, HW, Co, Npad, frames, out);
  EVW_LAUNCH_CHECK();
  return EVW_OK;
}

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
struct VaeBuilder : BuilderBase {
  Vae& V;
  VaePlan& P;
  __half *n16 = nullptr, *raw16 = nullptr, *h16 = nullptr, *resamp16 = nullptr, *in16 = nullptr, *q16 = nullptr, *k16 = nullptr,
         *vT16 = nullptr, *p16 = nullptr, *attn16 = nullptr;
  float *f0 = nullptr, *f2 = nullptr, *pp[2] = {nullptr, nullptr}, *y32 = nullptr, *scores = nullptr;

  VaeBuilder(Vae& v, VaePlan& p, void* ws, bool dry_) : BuilderBase(v.tensors, v.scalars, p, ws, dry_), V(v), P(p) {}

  // ResnetBlock2D without time embedding (resnet.py): x -> out, both fp32 [frames*hh*ww, C]
  void spatial_res(const std::string& sp, int frames, int hh, int ww, const float* x, int Cin, int Cout, float* out) {
    const long long S = (long long)hh * ww;
    const bool shortcut = Cin != Cout;
    gnorm(x, 0, Cin, nullptr, 0, frames, S, 1e-6f, sp + ".norm1", 1, n16, shortcut ? raw16 : nullptr);
    {
      GemmProblem pr;
      pr.a0 = n16; pr.w = W(sp + ".conv1.weight");
      pr.B = 1; pr.T = frames; pr.Y = hh; pr.X = ww; pr.C0 = Cin; pr.N = Cout; pr.K_total = 9LL * Cin;
      taps_conv3x3(pr);
      pr.ep.out = f0; pr.ep.out_fp16 = 0; pr.ep.bias = Wf(sp + ".conv1.bias");
      gemm(pr, sp + ".conv1");
    }
    gnorm(f0, 0, Cout, nullptr, 0, frames, S, 1e-6f, sp + ".norm2", 1, n16, nullptr);
    {
      GemmProblem pr;
      pr.a0 = n16; pr.w = W(sp + ".conv2.weight");
      pr.B = 1; pr.T = frames; pr.Y = hh; pr.X = ww; pr.C0 = Cout; pr.N = Cout; pr.K_total = 9LL * Cout;
      taps_conv3x3(pr);
      if (shortcut) {  // 1x1 conv_shortcut of the raw input as a tenth tap (weights and biases merged by vae.py)
        pr.a1 = raw16; pr.C1 = Cin; pr.K_total += Cin;
        pr.tap_dx[9] = pr.tap_dy[9] = pr.tap_dt[9] = 0; pr.tap_src[9] = 1; pr.num_taps = 10;
      } else {
        pr.ep.res1 = x; pr.ep.res1_fp16 = 0; pr.ep.s1 = 1.f;
      }
      pr.ep.out = out; pr.ep.out_fp16 = 0; pr.ep.bias = Wf(sp + ".conv2.bias");
      gemm(pr, sp + ".conv2");
    }
  }

  // SpatioTemporalResBlock without time embedding (unet_3d_blocks.py decoder blocks): frames = B videos of T frames
  void st_res(const std::string& pre, int B, int T, int hh, int ww, const float* x, int Cin, int Cout, float* out) {
    const long long S = (long long)hh * ww;
    const std::string tp = pre + ".temporal_res_block";
    spatial_res(pre + ".spatial_res_block", B * T, hh, ww, x, Cin, Cout, f2);
    gnorm(f2, 0, Cout, nullptr, 0, B, (long long)T * S, 1e-5f, tp + ".norm1", 1, n16, nullptr);
    auto tconv = [&](const std::string& name, GemmEpilogue ep) {
      GemmProblem pr;
      pr.a0 = n16; pr.w = W(name + ".weight");
      pr.B = B; pr.T = T; pr.Y = 1; pr.X = (int)S; pr.C0 = Cout; pr.N = Cout; pr.K_total = 3LL * Cout;
      pr.num_taps = 3;
      for (int i = 0; i < 3; ++i) { pr.tap_dx[i] = pr.tap_dy[i] = 0; pr.tap_dt[i] = (int8_t)(i - 1); pr.tap_src[i] = 0; }
      ep.bias = Wf(name + ".bias");
      pr.ep = ep;
      gemm(pr, name);
    };
    {
      GemmEpilogue ep;
      ep.out = f0; ep.out_fp16 = 0;
      tconv(tp + ".conv1", ep);
    }
    gnorm(f0, 0, Cout, nullptr, 0, B, (long long)T * S, 1e-5f, tp + ".norm2", 1, n16, nullptr);
    {
      // AlphaBlender (switch_spatial_to_temporal_mix): a x_s + (1 - a)(x_s + h) = x_s + (1 - a) h, a from vae.py
      const float alpha = (float)Sc(pre + ".time_mixer.alpha");
      GemmEpilogue ep;
      ep.out = out; ep.out_fp16 = 0; ep.s0 = 1.f - alpha; ep.res1 = f2; ep.s1 = 1.f;
      tconv(tp + ".conv2", ep);
    }
  }

  // Attention (one head of width C, GroupNorm on the input, residual): x -> out fp32 [frames*S, C]
  void attention(const std::string& pre, int frames, long long S, const float* x, int C, float* out) {
    const long long M = frames * S;
    if (S % 64 != 0 && fail.empty()) fail = "VAE attention needs (H/8)*(W/8) to be a multiple of 64";
    gnorm(x, 0, C, nullptr, 0, frames, S, 1e-6f, pre + ".group_norm", 0, n16, nullptr);
    { GemmEpilogue e; e.out = q16; e.out_fp16 = 1; linear(n16, M, C, pre + ".to_q", C, e); }
    { GemmEpilogue e; e.out = k16; e.out_fp16 = 1; linear(n16, M, C, pre + ".to_k", C, e); }
    const float scale = 1.0f / sqrtf((float)C);
    for (int f = 0; f < frames; ++f) {
      const std::string tag = pre + " frame " + std::to_string(f);
      {  // V^T = W_v X^T (the bias of to_v is folded into to_out's: softmax rows sum to one)
        GemmProblem pr;
        pr.a0 = W(pre + ".to_v.weight"); pr.w = n16 ? n16 + (long long)f * S * C : nullptr;
        if (dry) pr.w = nullptr;
        pr.X = C; pr.C0 = C; pr.N = (int)S; pr.K_total = C; pr.num_taps = 1;
        pr.ep.out = vT16; pr.ep.out_fp16 = 1;
        gemm(pr, tag + " V^T");
      }
      {  // scores = scale Q K^T, fp32
        GemmProblem pr;
        pr.a0 = dry ? nullptr : q16 + (long long)f * S * C; pr.w = dry ? nullptr : k16 + (long long)f * S * C;
        pr.X = (int)S; pr.C0 = C; pr.N = (int)S; pr.K_total = C; pr.num_taps = 1;
        pr.ep.out = scores; pr.ep.out_fp16 = 0; pr.ep.s0 = scale;
        gemm(pr, tag + " QK^T");
      }
      { const float* s = scores; __half* p = p16; const long long rows = S; const int cols = (int)S;
        push([=](cudaStream_t st) { return softmax_rows(s, p, rows, cols, st); }, 1, tag + " softmax", p, rows * cols, 1); }
      {  // O = P V
        GemmProblem pr;
        pr.a0 = p16; pr.w = vT16;
        pr.X = (int)S; pr.C0 = (int)S; pr.N = C; pr.K_total = S; pr.num_taps = 1;
        pr.ep.out = dry ? nullptr : attn16 + (long long)f * S * C; pr.ep.out_fp16 = 1;
        gemm(pr, tag + " PV");
      }
    }
    {
      GemmEpilogue e;
      e.out = out; e.out_fp16 = 0; e.res1 = x; e.s1 = 1.f;
      linear(attn16, M, C, pre + ".to_out.0", C, e);
    }
  }

  void alloc_common(long long maxMC, long long M_in, long long M_out, int frames, long long S_attn, int C_attn) {
    n16 = bump.take<__half>(maxMC);
    raw16 = bump.take<__half>(maxMC);
    resamp16 = bump.take<__half>(maxMC);
    in16 = bump.take<__half>(M_in * V.cfg.cin_pad);
    f0 = bump.take<float>(maxMC);
    f2 = bump.take<float>(maxMC);
    pp[0] = bump.take<float>(maxMC);
    pp[1] = bump.take<float>(maxMC);
    y32 = bump.take<float>(M_out * V.cfg.cout_pad);
    stats = bump.take<double>(64LL * std::max(frames, 1));
    const long long Ma = (long long)frames * S_attn;
    q16 = bump.take<__half>(Ma * C_attn);
    k16 = bump.take<__half>(Ma * C_attn);
    attn16 = bump.take<__half>(Ma * C_attn);
    vT16 = bump.take<__half>(S_attn * C_attn);
    scores = bump.take<float>(S_attn * S_attn);
    p16 = bump.take<__half>(S_attn * S_attn);
  }

  int finish(const char* what) {
    if (!fail.empty()) {
      set_error("evw_vae %s plan: %s", what, fail.c_str());
      return EVW_ERR_STATE;
    }
    return EVW_OK;
  }

  // ---- Encoder (vae.py Encoder) + quant_conv (folded into conv_out by vae.py): images -> moments
  int build_encode() {
    const VaeConfig& c = V.cfg;
    const int N = P.N;
    EVW_CHECK_ARG(P.H % 8 == 0 && P.W % 8 == 0, "evw_vae_encode: image size %dx%d must be divisible by 8", P.H, P.W);
    int lh[4], lw[4];
    long long lM[4];
    for (int l = 0; l < 4; ++l) { lh[l] = P.H >> l; lw[l] = P.W >> l; lM[l] = (long long)N * lh[l] * lw[l]; }
    long long maxMC = 0;
    for (int l = 0; l < 4; ++l) maxMC = std::max(maxMC, lM[l] * std::max(c.boc[l], l ? c.boc[l - 1] : c.boc[0]));
    const long long S3 = (long long)lh[3] * lw[3];
    alloc_common(maxMC, lM[0], lM[3], N, S3, c.boc[3]);
    Vae* v = &V;
    {
      __half* o = in16; const int Cin = c.in_channels, Cpad = c.cin_pad, sp = c.split; const long long HW = (long long)P.H * P.W;
      push([=](cudaStream_t st) { return nchw_to_nhwc_f16(v->args.in, N, Cin, HW, Cpad, sp, o, st); }, 1, "images -> fp16 channels-last");
    }
    int cur = 0;
    conv3x3_simple(in16, N, lh[0], lw[0], c.cin_pad, "encoder.conv_in", c.boc[0], pp[cur]);
    const float* x = pp[cur];
    int xC = c.boc[0];
    for (int i = 0; i < 4; ++i) {
      const std::string bp = "encoder.down_blocks." + std::to_string(i);
      for (int j = 0; j < c.layers; ++j) {
        spatial_res(bp + ".resnets." + std::to_string(j), N, lh[i], lw[i], x, xC, c.boc[i], pp[cur ^ 1]);
        cur ^= 1; x = pp[cur]; xC = c.boc[i];
      }
      if (i < 3) {
        // Downsample2D(padding=0): zero-pad right / bottom by one, conv 3x3 stride 2, on the 4 phase images:
        // in[2y + ky] = phase (ky & 1) at row y + (ky >> 1); rows / columns past the image read as zero (TMA fill)
        { const float* xi = x; __half* o = resamp16; const long long n = N; const int hh = lh[i], ww = lw[i], C = xC;
          push([=](cudaStream_t st) { return downsplit(xi, o, n, hh, ww, C, st); }, 1, "downsplit"); }
        GemmProblem pr;
        pr.a0 = resamp16; pr.w = W(bp + ".downsamplers.0.conv.weight");
        pr.B = N; pr.T = 1; pr.Tmap = 4; pr.Y = lh[i + 1]; pr.X = lw[i + 1]; pr.C0 = xC; pr.N = xC; pr.K_total = 9LL * xC;
        pr.num_taps = 9;
        int t = 0;
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx, ++t) {
            pr.tap_dx[t] = (int8_t)(kx >> 1);
            pr.tap_dy[t] = (int8_t)(ky >> 1);
            pr.tap_dt[t] = (int8_t)(2 * (ky & 1) + (kx & 1));  // phase image index
            pr.tap_src[t] = 0;
          }
        pr.ep.out = pp[cur ^ 1]; pr.ep.out_fp16 = 0; pr.ep.bias = Wf(bp + ".downsamplers.0.conv.bias");
        gemm(pr, bp + ".downsamplers.0.conv");
        cur ^= 1; x = pp[cur];
      }
    }
    {
      const int C = c.boc[3];
      spatial_res("encoder.mid_block.resnets.0", N, lh[3], lw[3], x, C, C, pp[cur ^ 1]);
      cur ^= 1; x = pp[cur];
      attention("encoder.mid_block.attentions.0", N, S3, x, C, pp[cur ^ 1]);
      cur ^= 1; x = pp[cur];
      spatial_res("encoder.mid_block.resnets.1", N, lh[3], lw[3], x, C, C, pp[cur ^ 1]);
      cur ^= 1; x = pp[cur];
    }
    gnorm(x, 0, c.boc[3], nullptr, 0, N, S3, 1e-6f, "encoder.conv_norm_out", 1, n16, nullptr);
    conv3x3_simple(n16, N, lh[3], lw[3], c.boc[3], "encoder.conv_out", c.cout_pad, y32);
    {
      const float* y = y32; const int Co = 2 * c.latent, Np = c.cout_pad;
      push([=](cudaStream_t st) { return nhwc_to_nchw_f32(y, N, Co, S3, Np, 0, v->args.out, st); }, 1, "moments -> NCHW");
    }
    return finish("encode");
  }

  // ---- TemporalDecoder (autoencoder_kl_temporal_decoder.py): latents [N = B*F, 4, h, w] -> frames [N, 3, 8h, 8w]
  int build_decode() {
    const VaeConfig& c = V.cfg;
    const int N = P.N, F = P.F, B = N / F;
    EVW_CHECK_ARG(F >= 1 && N % F == 0, "evw_vae_decode: %d latents are not a multiple of num_frames=%d", N, F);
    int lh[4], lw[4];
    long long lM[4];
    for (int l = 0; l < 4; ++l) { lh[l] = P.H << l; lw[l] = P.W << l; lM[l] = (long long)N * lh[l] * lw[l]; }
    const int rev[4] = {c.boc[3], c.boc[2], c.boc[1], c.boc[0]};
    long long maxMC = 0;
    {
      int prev = rev[0];
      for (int s = 0; s < 4; ++s) {
        maxMC = std::max(maxMC, lM[s] * std::max(prev, rev[s]));
        if (s < 3) maxMC = std::max(maxMC, lM[s + 1] * rev[s]);
        prev = rev[s];
      }
    }
    const long long S0 = (long long)lh[0] * lw[0];
    alloc_common(maxMC, lM[0], lM[3], N, S0, rev[0]);
    Vae* v = &V;
    {
      __half* o = in16; const int Cin = c.latent, Cpad = c.cin_pad;
      push([=](cudaStream_t st) { return nchw_to_nhwc_f16(v->args.in, N, Cin, S0, Cpad, 1, o, st); }, 1, "latents -> fp16 channels-last");
    }
    int cur = 0;
    conv3x3_simple(in16, N, lh[0], lw[0], c.cin_pad, "decoder.conv_in", rev[0], pp[cur]);
    const float* x = pp[cur];
    int xC = rev[0];
    {
      const std::string mp = "decoder.mid_block";
      st_res(mp + ".resnets.0", B, F, lh[0], lw[0], x, xC, xC, pp[cur ^ 1]);
      cur ^= 1; x = pp[cur];
      for (int i = 1; i < c.layers; ++i) {
        attention(mp + ".attentions." + std::to_string(i - 1), N, S0, x, xC, pp[cur ^ 1]);
        cur ^= 1; x = pp[cur];
        st_res(mp + ".resnets." + std::to_string(i), B, F, lh[0], lw[0], x, xC, xC, pp[cur ^ 1]);
        cur ^= 1; x = pp[cur];
      }
    }
    for (int s = 0; s < 4; ++s) {
      const std::string bp = "decoder.up_blocks." + std::to_string(s);
      for (int j = 0; j < c.layers + 1; ++j) {
        st_res(bp + ".resnets." + std::to_string(j), B, F, lh[s], lw[s], x, xC, rev[s], pp[cur ^ 1]);
        cur ^= 1; x = pp[cur]; xC = rev[s];
      }
      if (s < 3) {
        if (gemm_strided_out_ok(lw[s], lh[s], xC)) {  // four 2x2 phase convolutions of the low-resolution tensor
          { const float* xi = x; __half* o = resamp16; const long long n = lM[s] * xC;
            push([=](cudaStream_t st) { return cast_f16(xi, o, n, st); }, 1, "cast (upsampler operand)"); }
          upconv2x(resamp16, N, lh[s], lw[s], xC, bp + ".upsamplers.0.conv", xC, pp[cur ^ 1]);
        } else {  // the literal form on the up-sampled fp16 image
          { const float* xi = x; __half* o = resamp16; const long long n = N; const int hh = lh[s], ww = lw[s], C = xC;
            push([=](cudaStream_t st) { return upsample2x(xi, o, n, hh, ww, C, st); }, 1, "upsample2x"); }
          conv3x3_simple(resamp16, N, lh[s + 1], lw[s + 1], xC, bp + ".upsamplers.0.conv", xC, pp[cur ^ 1]);
        }
        cur ^= 1; x = pp[cur];
      }
    }
    const long long S3 = (long long)lh[3] * lw[3];
    gnorm(x, 0, xC, nullptr, 0, N, S3, 1e-6f, "decoder.conv_norm_out", 1, n16, nullptr);
    conv3x3_simple(n16, N, lh[3], lw[3], xC, "decoder.conv_out", c.cout_pad, y32);
    {
      const float* y = y32; const float* w = Wf("decoder.time_conv_out.weight"); const float* b = Wf("decoder.time_conv_out.bias");
      const int Co = c.out_channels, Np = c.cout_pad;
      push([=](cudaStream_t st) { return time_conv_out(y, w, b, F, S3, Co, Np, N, v->args.out, st); }, 1, "time_conv_out -> NCHW");
    }
    return finish("decode");
  }

  int build() { return P.mode == 0 ? build_encode() : build_decode(); }
};

int vae_ensure_plan(Vae* V, int mode, int N, int F, int H, int W, void* ws, long long ws_bytes) {
  auto& slot = V->plan[mode];
  if (slot && slot->N == N && slot->F == F && slot->H == H && slot->W == W && slot->ws == ws) return EVW_OK;
  EVW_CHECK_ARG(N >= 1 && N <= 64 && H >= 8 && W >= 8, "evw_vae: unsupported shape N=%d H=%d W=%d", N, H, W);
  EVW_CHECK_ARG(((uintptr_t)ws & 1023) == 0, "evw_vae: workspace must be 1024-byte aligned");
  auto plan = std::make_unique<VaePlan>();
  plan->mode = mode; plan->N = N; plan->F = F; plan->H = H; plan->W = W; plan->ws = ws; plan->ws_bytes = ws_bytes;
  {
    VaePlan sizing = *plan;
    VaeBuilder dry(*V, sizing, nullptr, true);
    int rc = dry.build();
    if (rc) return rc;
    if (dry.bump.off > ws_bytes) {
      set_error("evw_vae: workspace %lld bytes < required %lld", ws_bytes, dry.bump.off);
      return EVW_ERR_WORKSPACE;
    }
  }
  VaeBuilder b(*V, *plan, ws, false);
  int rc = b.build();
  if (rc) return rc;
  slot = std::move(plan);
  return EVW_OK;
}

// EVW_VAE_DEBUG=1: synchronise after every op and print the absolute maximum / non-finite count of its output
int vae_run(Vae* V, int mode, cudaStream_t st) {
  static const bool debug = getenv("EVW_VAE_DEBUG") != nullptr;
  VaePlan& P = *V->plan[mode];
  size_t i = 0;
  for (auto& op : P.ops) {
    int rc = op(st);
    if (rc) return rc;
    if (debug) {
      const OpMeta& m = P.meta[i];
      cudaError_t e = cudaStreamSynchronize(st);
      double amax = 0;
      long long bad = 0;
      if (e == cudaSuccess && m.out && m.n > 0) {
        std::vector<char> host((size_t)m.n * (m.fp16 ? 2 : 4));
        cudaMemcpy(host.data(), m.out, host.size(), cudaMemcpyDeviceToHost);
        for (long long j = 0; j < m.n; ++j) {
          const float v = m.fp16 ? __half2float(reinterpret_cast<const __half*>(host.data())[j]) : reinterpret_cast<const float*>(host.data())[j];
          if (!(v == v) || v > 3.0e38f || v < -3.0e38f) ++bad;
          else if (fabs(v) > amax) amax = fabs(v);
        }
      }
      fprintf(stderr, "[evw_vae] op %4zu %-80s %s absmax %.4g nonfinite %lld\n", i, m.label.c_str(),
              e == cudaSuccess ? "ok" : cudaGetErrorString(e), amax, bad);
      if (e != cudaSuccess) {
        set_error("op %zu (%s): %s", i, m.label.c_str(), cudaGetErrorString(e));
        return EVW_ERR_CUDA;
      }
    }
    ++i;
  }
  return EVW_OK;
}

}  // namespace
}  // namespace evw

using evw::Vae;

extern "C" int evw_vae_create(void** handle, const int* cfg_ints, int n_ints, const char* const* tensor_names,
                              const void* const* tensor_ptrs, int n_tensors, const char* const* scalar_names,
                              const double* scalar_values, int n_scalars) {
  EVW_CHECK_ARG(handle && cfg_ints && n_ints >= 8, "evw_vae_create: bad config array");
  auto* V = new Vae();
  evw::VaeConfig& c = V->cfg;
  int k = 0;
  c.in_channels = cfg_ints[k++]; c.out_channels = cfg_ints[k++]; c.latent = cfg_ints[k++];
  for (int i = 0; i < 4; ++i) c.boc[i] = cfg_ints[k++];
  c.layers = cfg_ints[k++];
  bool ok = c.in_channels * 3 <= c.cin_pad && c.latent * 3 <= c.cin_pad && 2 * c.latent <= c.cout_pad && c.out_channels <= 4 &&
            c.layers >= 1;
  for (int i = 0; i < 4; ++i) ok = ok && c.boc[i] % 64 == 0 && c.boc[i] >= 64;
  if (!ok) {
    delete V;
    evw::set_error("evw_vae_create: unsupported configuration (block widths must be multiples of 64)");
    return EVW_ERR_INVALID;
  }
  for (int i = 0; i < n_tensors; ++i) V->tensors[tensor_names[i]] = tensor_ptrs[i];
  for (int i = 0; i < n_scalars; ++i) V->scalars[scalar_names[i]] = scalar_values[i];
  *handle = V;
  return EVW_OK;
}

extern "C" int evw_vae_destroy(void* handle) {
  delete (Vae*)handle;
  return EVW_OK;
}

extern "C" int64_t evw_vae_workspace_bytes(void* handle, int mode, int N, int num_frames, int H, int W) {
  if (!handle || (mode != 0 && mode != 1)) return -1;
  Vae* V = (Vae*)handle;
  evw::VaePlan sizing;
  sizing.mode = mode; sizing.N = N; sizing.F = num_frames; sizing.H = H; sizing.W = W;
  evw::VaeBuilder dry(*V, sizing, nullptr, true);
  if (dry.build() != 0) return -1;
  return evw::align_up(dry.bump.off, 1024) + 1024;
}

extern "C" int evw_vae_encode(void* handle, const float* images, float* moments, int N, int H, int W, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  EVW_CHECK_ARG(handle && images && moments && workspace, "evw_vae_encode: null pointer");
  Vae* V = (Vae*)handle;
  int rc = evw::vae_ensure_plan(V, 0, N, 1, H, W, workspace, workspace_bytes);
  if (rc) return rc;
  V->args.in = images; V->args.out = moments;
  return evw::vae_run(V, 0, (cudaStream_t)stream);
}

extern "C" int evw_vae_decode(void* handle, const float* latents, float* frames, int N, int num_frames, int h, int w,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  EVW_CHECK_ARG(handle && latents && frames && workspace, "evw_vae_decode: null pointer");
  Vae* V = (Vae*)handle;
  int rc = evw::vae_ensure_plan(V, 1, N, num_frames, h, w, workspace, workspace_bytes);
  if (rc) return rc;
  V->args.in = latents; V->args.out = frames;
  return evw::vae_run(V, 1, (cudaStream_t)stream);
}

extern "C" int evw_vae_plan_info(void* handle, int mode, int64_t* launches, double* flops, int64_t* gn_fused) {
  EVW_CHECK_ARG(handle && (mode == 0 || mode == 1), "evw_vae_plan_info: bad arguments");
  Vae* V = (Vae*)handle;
  if (!V->plan[mode]) {
    evw::set_error("evw_vae_plan_info: no plan yet (call encode / decode first)");
    return EVW_ERR_STATE;
  }
  if (launches) *launches = V->plan[mode]->launches;
  if (flops) *flops = V->plan[mode]->flops;
  if (gn_fused) *gn_fused = V->plan[mode]->gn_fused;
  return EVW_OK;
}
