// Host-side orchestration of the denoise step (hot path 1): one C-ABI call enqueues the whole
// UNetSpatioTemporalConditionModel.forward (evoworld/trainer/unet_plucker.py:355-488, blocks from
// diffusers 0.31 unet_3d_blocks.py) — or the whole loop body of pipeline_evoworld.py:689-725 — as a
// pre-planned sequence of sm_100a kernels on the caller's stream.  No allocation, no host sync.
//
// Data layout in HBM (caller-provided workspace): activations channels-last, rows = (b, t, y, x);
// fp32 residual stream + skip tensors, fp16 GEMM operands produced by the normalisation kernels.
#include "common.h"
#include "plan_builder.h"
#include "tc_gemm.h"
#include "unet_elem.h"

#include <cmath>
#include <cstdlib>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace evw {
namespace {

struct Config {
  int in_channels = 18, out_channels = 4;
  int boc[4] = {320, 640, 1280, 1280};
  int heads[4] = {5, 10, 20, 20};
  int down_attn[4] = {1, 1, 1, 0};
  int layers = 2;
  int cross_dim = 1024;
  int add_dim = 256;
  int temb_dim = 1280;
  int cin_pad = 64, cout_pad = 16;
  int temb_total = 0, xattn_total = 0;
  float eps_cross = 1e-6f, eps_plain = 1e-5f, eps_mid = 1e-5f, eps_up = 1e-6f;
  // split precision (scalar "precision.split", set by unet.py::pack_parameters which also packs the matching weights):
  // the low-FLOP layers that dominate the fp16 rounding error of the step (tools/precision_sim.py: conv_in, conv_out,
  // proj_in, proj_out and the fp16 storage of the spatial conv1 output = 0.75 of the error variance, < 4 % of the FLOPs)
  // take their operands as fp16 head + fp16 tail and keep conv1's output in fp32
  int split = 0;
};

constexpr int kSmall = 16384;  // halves per small staging region (B <= 8 rows of <= 2048 columns)

struct CallArgs {  // per-call pointers/scalars referenced by the planned ops
  const float* sample = nullptr;   // [B,T,Cin,h,w] (forward) or null
  const float* latents = nullptr;  // [1,T,4,h,w]   (denoise step)
  const float* cond = nullptr;     // [2,T,Cc,h,w]
  float* latents_out = nullptr;
  float* out = nullptr;            // [B,T,Co,h,w]
  const float* ehs = nullptr;      // [B,1,cross]
  const float* added = nullptr;    // [B,3]
  float timestep = 0.f, sigma = 1.f, sigma_next = 0.f, g_min = 1.f, g_max = 3.f;
  int mode = 0;  // 0 = forward, 1 = denoise step
};

// Pointers baked into a captured graph of the plan (kernel parameters); the per-step scalars are not (set_step_args).
struct GraphKey {
  int mode = -1;
  const void *sample = nullptr, *latents = nullptr, *cond = nullptr, *latents_out = nullptr, *out = nullptr, *ehs = nullptr,
             *added = nullptr;
  bool operator==(const GraphKey& o) const {
    return mode == o.mode && sample == o.sample && latents == o.latents && cond == o.cond && latents_out == o.latents_out &&
           out == o.out && ehs == o.ehs && added == o.added;
  }
};

struct Plan : PlanCore {
  int B = 0, T = 0, h = 0, w = 0;
  void* ws = nullptr;
  long long ws_bytes = 0;
  float *tsteps = nullptr, *dargs = nullptr;  // device-side per-step scalars (in the workspace)
  // plan-owned copies of the small / per-call tensors, so that the captured graph sees constant addresses:
  // encoder_hidden_states [B, cross], added_time_ids [B, 3], and for forward() the sample and the output
  float *ehs_stage = nullptr, *added_stage = nullptr, *sample_stage = nullptr, *out_stage = nullptr;
  // whole-plan CUDA graphs, one per distinct set of caller pointers (the denoise loop re-uses its buffers: one capture)
  bool warm = false, graphs_ok = true;
  cudaStream_t capture_stream = nullptr;
  std::vector<std::pair<GraphKey, cudaGraphExec_t>> graphs;
  long long graph_replays = 0;
  ~Plan() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    if (capture_stream) cudaStreamDestroy(capture_stream);
  }
};

struct UNet {
  Config cfg;
  std::unordered_map<std::string, const void*> tensors;
  std::unordered_map<std::string, double> scalars;
  std::unique_ptr<Plan> plan;
  CallArgs args;
  std::string err;
};

struct Builder : BuilderBase {
  UNet& U;
  Plan& P;
  int B, T;
  int lh[4], lw[4];
  long long lS[4], lM[4];
  // shared scratch
  __half *n16 = nullptr, *raw16 = nullptr, *h16 = nullptr, *qkv16 = nullptr, *attn16 = nullptr, *ff16 = nullptr, *in16 = nullptr,
         *resamp16 = nullptr, *small16 = nullptr, *lo16 = nullptr;
  float *f0 = nullptr, *f1 = nullptr, *f2 = nullptr, *pp[2] = {nullptr, nullptr}, *y32 = nullptr;
  float *emb = nullptr, *emb2 = nullptr, *temb_all = nullptr, *xattn_all = nullptr, *tsteps = nullptr, *dargs = nullptr;

  Builder(UNet& u, Plan& p, void* ws, bool dry_) : BuilderBase(u.tensors, u.scalars, p, ws, dry_), U(u), P(p) {}

  // split-precision linear: (a_hi + a_lo) (W_hi + W_lo)^T without the tail x tail term, as three taps
  // [a_hi | a_lo | a_hi] x [W_hi | W_hi | W_lo] (weight [N, 3K], packed by unet.py)
  void linear_split(const __half* a_hi, const __half* a_lo, long long M, int K, const std::string& wname, int N, GemmEpilogue ep) {
    GemmProblem pr;
    pr.a0 = a_hi; pr.a1 = a_lo; pr.w = W(wname + ".weight");
    pr.X = (int)M; pr.C0 = K; pr.C1 = K; pr.N = N; pr.K_total = 3LL * K; pr.num_taps = 3;
    pr.tap_src[0] = 0; pr.tap_src[1] = 1; pr.tap_src[2] = 0;
    ep.bias = Wf(wname + ".bias");
    pr.ep = ep;
    gemm(pr, wname);
  }
  void lnorm(const float* x, const float* rowvec, long long rv_div, long long rv_mod, long long rows, int C,
             const std::string& name, __half* out) {
    const float* g = Wf(name + ".weight");
    const float* b = Wf(name + ".bias");
    push([=](cudaStream_t st) { return layer_norm(x, rowvec, rv_div, rv_mod, rows, C, 1e-5f, g, b, out, st); }, 1, name, out, rows * C, 1);
  }

  // ---- SpatioTemporalResBlock (diffusers resnet.py): x = cat(x0, x1) -> out
  void resblock(const std::string& pre, int lvl, const float* x0, int C0, const float* x1, int C1, int Cout, float eps,
                float* out) {
    const int Cin = C0 + C1;
    const long long M = lM[lvl], S = lS[lvl];
    const int BF = B * T;
    const bool shortcut = Cin != Cout;
    const std::string sp = pre + ".spatial_res_block", tp = pre + ".temporal_res_block";
    gnorm(x0, 0, C0, x1, C1, BF, S, eps, sp + ".norm1", 1, n16, shortcut ? raw16 : nullptr);
    {
      GemmProblem pr;
      pr.a0 = n16; pr.w = W(sp + ".conv1.weight");
      pr.B = 1; pr.T = BF; pr.Y = lh[lvl]; pr.X = lw[lvl]; pr.C0 = Cin; pr.N = Cout; pr.K_total = 9LL * Cin;
      taps_conv3x3(pr);
      // split precision: the conv1 output feeds GroupNorm statistics and stays fp32 (f0 is free inside a res block)
      if (U.cfg.split) { pr.ep.out = f0; pr.ep.out_fp16 = 0; }
      else { pr.ep.out = h16; pr.ep.out_fp16 = 1; }
      pr.ep.bias = Wf(sp + ".conv1.bias");
      pr.ep.rowvec = temb_all + (long long)Sc(sp + ".time_emb_proj.offset");
      pr.ep.rv_ld = U.cfg.temb_total; pr.ep.rv_div = (long long)T * S; pr.ep.rv_mod = B;
      gemm(pr, sp + ".conv1");
    }
    if (U.cfg.split) gnorm(f0, 0, Cout, nullptr, 0, BF, S, eps, sp + ".norm2", 1, n16, nullptr);
    else gnorm(h16, 1, Cout, nullptr, 0, BF, S, eps, sp + ".norm2", 1, n16, nullptr);
    {
      GemmProblem pr;
      pr.a0 = n16; pr.w = W(sp + ".conv2.weight");
      pr.B = 1; pr.T = BF; pr.Y = lh[lvl]; pr.X = lw[lvl]; pr.C0 = Cout; pr.N = Cout; pr.K_total = 9LL * Cout;
      taps_conv3x3(pr);
      if (shortcut) {
        pr.a1 = raw16; pr.C1 = Cin; pr.K_total += Cin;
        pr.tap_dx[9] = pr.tap_dy[9] = pr.tap_dt[9] = 0; pr.tap_src[9] = 1; pr.num_taps = 10;
      } else {
        pr.ep.res1 = x0; pr.ep.res1_fp16 = 0; pr.ep.s1 = 1.f;
      }
      pr.ep.out = f2; pr.ep.out_fp16 = 0; pr.ep.bias = Wf(sp + ".conv2.bias");
      gemm(pr, sp + ".conv2");
    }
    // temporal part: GroupNorm statistics over (T, h, w) per batch element
    gnorm(f2, 0, Cout, nullptr, 0, B, (long long)T * S, eps, tp + ".norm1", 1, n16, nullptr);
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
      ep.out = h16; ep.out_fp16 = 1;
      ep.rowvec = temb_all + (long long)Sc(tp + ".time_emb_proj.offset");
      ep.rv_ld = U.cfg.temb_total; ep.rv_div = (long long)T * S; ep.rv_mod = B;
      tconv(tp + ".conv1", ep);
    }
    gnorm(h16, 1, Cout, nullptr, 0, B, (long long)T * S, eps, tp + ".norm2", 1, n16, nullptr);
    {
      // AlphaBlender: alpha x_s + (1 - alpha)(x_s + h) = x_s + (1 - alpha) h
      const float alpha = (float)Sc(pre + ".time_mixer.alpha");
      GemmEpilogue ep;
      ep.out = out; ep.out_fp16 = 0; ep.s0 = 1.f - alpha; ep.res1 = f2; ep.s1 = 1.f;
      tconv(tp + ".conv2", ep);
    }
    (void)M;
  }

  // FeedForward (GEGLU projection -> net.2), optionally over row panels (EVW_FF_PANEL_MB > 0): the 4C-wide fp16 intermediate
  // of a panel would be written and read back while still in the 126 MB L2, every panel re-using the same buffer — at
  // level 0 the intermediate is 660 MB per feed-forward (1.3 GB of HBM traffic for 0.6 TFLOP).  MEASURED SLOWER and therefore
  // off by default: 114.7 ms per step without panels, 114.8 / 117.7 / 119.0 ms with 160 / 80 / 40 MB panels
  // (profiles/r02ag_bench_ff_panel_ab.log) — the extra launches and partial last waves cost more than the L2 hits return.
  // ep2 = epilogue of net.2 for the whole tensor; its row-indexed operands are offset per panel here.
  void feed_forward(const __half* a16, long long M, int C, const std::string& ff, GemmEpilogue ep2) {
    static const long long panel_mb = [] { const char* e = getenv("EVW_FF_PANEL_MB"); return e ? atoll(e) : 0ll; }();
    const long long bytes_per_row = 4LL * C * 2;
    long long rows_per_panel = M;
    if (panel_mb > 0 && M * bytes_per_row > panel_mb * (1ll << 20) * 3 / 2) {
      const long long panels = (M * bytes_per_row + panel_mb * (1ll << 20) - 1) / (panel_mb * (1ll << 20));
      rows_per_panel = ((M + panels - 1) / panels + 127) / 128 * 128;  // whole 128-row tiles
    }
    const int o_esz = ep2.out_fp16 ? 2 : 4;
    for (long long r0 = 0; r0 < M; r0 += rows_per_panel) {
      const long long rows = std::min(rows_per_panel, M - r0);
      GemmEpilogue e1;
      e1.out = ff16; e1.out_fp16 = 1; e1.geglu = 1;
      linear(dry ? a16 : a16 + r0 * C, rows, C, ff + ".net.0.proj", 8 * C, e1);
      GemmEpilogue e2 = ep2;
      if (!dry) {
        e2.out = (char*)ep2.out + r0 * C * o_esz;
        if (ep2.out_lo) e2.out_lo = (char*)ep2.out_lo + r0 * C * 2;
        if (ep2.res1) e2.res1 = (const char*)ep2.res1 + r0 * C * (ep2.res1_fp16 ? 2 : 4);
        if (ep2.res2) e2.res2 = ep2.res2 + r0 * C;
      }
      e2.rv_row0 = ep2.rv_row0 + r0;
      linear(ff16, rows, 4 * C, ff + ".net.2", C, e2);
    }
  }

  // ---- TransformerSpatioTemporalModel (diffusers transformer_temporal.py); in-place on x allowed (out may == x)
  void transformer(const std::string& pre, int lvl, const float* x, int C, int heads, float* out) {
    const long long M = lM[lvl], S = lS[lvl];
    const int BF = B * T;
    const std::string sb = pre + ".transformer_blocks.0", tb = pre + ".temporal_transformer_blocks.0";
    const float alpha = (float)Sc(pre + ".time_mixer.alpha");
    const float* tpos = Wf(pre + ".time_pos_embed.table");  // [32, C]
    const float* xv_s = xattn_all + (long long)Sc(sb + ".attn2.offset");
    const float* xv_t = xattn_all + (long long)Sc(tb + ".attn2.offset");
    auto ep_f32 = [](float* o) { GemmEpilogue e; e.out = o; e.out_fp16 = 0; return e; };
    auto ep_f16 = [](__half* o) { GemmEpilogue e; e.out = o; e.out_fp16 = 1; return e; };

    const bool split = U.cfg.split != 0;
    gnorm(x, 0, C, nullptr, 0, BF, S, 1e-6f, pre + ".norm", 0, n16, nullptr, split ? lo16 : nullptr);
    if (split) linear_split(n16, lo16, M, C, pre + ".proj_in", C, ep_f32(f0));
    else linear(n16, M, C, pre + ".proj_in", C, ep_f32(f0));
    // --- spatial block
    lnorm(f0, nullptr, 1, 1, M, C, sb + ".norm1", n16);
    linear(n16, M, C, sb + ".attn1.qkv", 3 * C, ep_f16(qkv16), false);
    {
      const __half* q = qkv16; __half* o = attn16; int S_ = (int)S;
      push([=](cudaStream_t st) { return spatial_attention(q, o, BF, S_, heads, st); }, 1, sb + ".attn1.sdpa", o, M * C, 1);
      P.flops += 4.0 * BF * heads * (double)S * (double)S * 64.0;
    }
    {
      // x += attn1(...);  x += attn2(norm2(x), ehs): one key => softmax == 1 => to_out(to_v(ehs_b)) broadcast
      GemmEpilogue ep = ep_f32(f0);
      ep.res1 = f0; ep.rowvec = xv_s; ep.rv_ld = U.cfg.xattn_total; ep.rv_div = (long long)T * S; ep.rv_mod = B;
      linear(attn16, M, C, sb + ".attn1.to_out.0", C, ep);
    }
    lnorm(f0, nullptr, 1, 1, M, C, sb + ".norm3", n16);
    {
      GemmEpilogue ep = ep_f32(f0);
      ep.res1 = f0;
      feed_forward(n16, M, C, sb + ".ff", ep);  // f0 = x_spatial
    }
    // --- temporal block on x_spatial + time_pos_embed[t]
    lnorm(f0, tpos, S, T, M, C, tb + ".norm_in", n16);
    {
      GemmEpilogue ep = ep_f32(f1);
      ep.res1 = f0; ep.rowvec = tpos; ep.rv_ld = C; ep.rv_div = S; ep.rv_mod = T;
      feed_forward(n16, M, C, tb + ".ff_in", ep);  // f1 = ff_in(...) + (x_s + emb)
    }
    lnorm(f1, nullptr, 1, 1, M, C, tb + ".norm1", n16);
    linear(n16, M, C, tb + ".attn1.qkv", 3 * C, ep_f16(qkv16), false);
    {
      const __half* q = qkv16; __half* o = attn16; int B_ = B, T_ = T;
      push([=](cudaStream_t st) { return temporal_attention(q, o, B_, T_, S, heads, st); }, 1, tb + ".attn1.sdpa", o, M * C, 1);
      P.flops += 4.0 * B * S * heads * (double)T * (double)T * 64.0;
    }
    {
      GemmEpilogue ep = ep_f32(f1);
      ep.res1 = f1; ep.rowvec = xv_t; ep.rv_ld = U.cfg.xattn_total; ep.rv_div = (long long)T * S; ep.rv_mod = B;
      linear(attn16, M, C, tb + ".attn1.to_out.0", C, ep);
    }
    lnorm(f1, nullptr, 1, 1, M, C, tb + ".norm3", n16);
    {
      // time_mixer: alpha x_s + (1 - alpha)(y + ff(y))  -> fp16 operand of proj_out
      // (in place over n16: a panel's rows are rewritten only after its own GEGLU projection has consumed them)
      GemmEpilogue ep = ep_f16(n16);
      ep.s0 = 1.f - alpha; ep.res1 = f1; ep.s1 = 1.f - alpha; ep.res2 = f0; ep.s2 = alpha;
      if (split) ep.out_lo = lo16;
      feed_forward(n16, M, C, tb + ".ff", ep);
    }
    {
      GemmEpilogue ep = ep_f32(out);
      ep.res1 = x;
      if (split) linear_split(n16, lo16, M, C, pre + ".proj_out", C, ep);
      else linear(n16, M, C, pre + ".proj_out", C, ep);
    }
  }

  int build() {
    const Config& c = U.cfg;
    B = P.B; T = P.T;
    const int BF = B * T;
    lh[0] = P.h; lw[0] = P.w;
    for (int l = 1; l < 4; ++l) { lh[l] = (lh[l - 1] + 1) / 2; lw[l] = (lw[l - 1] + 1) / 2; }
    long long maxMC = 0, maxMCin = 0;
    for (int l = 0; l < 4; ++l) {
      lS[l] = (long long)lh[l] * lw[l];
      lM[l] = lS[l] * BF;
      maxMC = std::max(maxMC, lM[l] * c.boc[l]);
    }
    // largest concatenated resblock input: up path, level l: boc[l] + boc[l] or boc[l+1] + boc[l] ...
    for (int l = 0; l < 4; ++l) {
      int cmax = 2 * c.boc[l];
      if (l + 1 < 4) cmax = std::max(cmax, c.boc[l + 1] + c.boc[l]);
      maxMCin = std::max(maxMCin, lM[l] * cmax);
    }
    n16 = bump.take<__half>(maxMCin);
    raw16 = bump.take<__half>(maxMCin);
    h16 = bump.take<__half>(maxMC);
    if (c.split) lo16 = bump.take<__half>(maxMC);  // fp16 tails of the split-precision operands
    qkv16 = bump.take<__half>(3 * maxMC);
    attn16 = bump.take<__half>(maxMC);
    ff16 = bump.take<__half>(4 * maxMC);
    in16 = bump.take<__half>(lM[0] * c.cin_pad);
    resamp16 = bump.take<__half>(maxMC * 4);  // upsampled / phase-split operand of the resampling convs
    f0 = bump.take<float>(maxMC);
    f1 = bump.take<float>(maxMC);
    f2 = bump.take<float>(maxMC);
    // the up-sampling convolutions write [M_{l-1}, boc[l]] (e.g. 640 channels at level 0)
    long long maxPP = maxMC;
    for (int l = 1; l < 4; ++l) maxPP = std::max(maxPP, lM[l - 1] * c.boc[l]);
    pp[0] = bump.take<float>(maxPP);
    pp[1] = bump.take<float>(maxPP);
    y32 = bump.take<float>(lM[0] * c.cout_pad);
    stats = bump.take<double>(64LL * std::max(BF, 1));  // [insts, 32, 2] double sums
    small16 = bump.take<__half>(5LL * kSmall);
    emb = bump.take<float>(8LL * c.temb_dim);
    emb2 = bump.take<float>(8LL * c.temb_dim);
    temb_all = bump.take<float>((long long)B * c.temb_total);
    xattn_all = bump.take<float>((long long)B * c.xattn_total);
    tsteps = bump.take<float>(64);
    dargs = bump.take<float>(64);
    P.tsteps = tsteps; P.dargs = dargs;
    P.ehs_stage = bump.take<float>((long long)B * c.cross_dim);
    P.added_stage = bump.take<float>(64);
    P.sample_stage = bump.take<float>((long long)BF * c.in_channels * lS[0]);
    P.out_stage = bump.take<float>((long long)BF * c.out_channels * lS[0]);

    UNet* u = &U;
    const long long HW = lS[0];
    // ---- inputs -> fp16 channels-last, padded to 64 channels
    {
      __half* o = in16; int Cin = c.in_channels, Cpad = c.cin_pad, T_ = T, B_ = B, sp_ = c.split; const float* da = dargs;
      push([=](cudaStream_t st) {
        const CallArgs& a = u->args;
        if (a.mode == 1) return pre_concat(a.latents, a.cond, B_, T_, 4, Cin - 4, HW, da, Cpad, sp_, o, st);
        return nchw_to_nhwc_f16(a.sample, (long long)B_ * T_, Cin, HW, Cpad, sp_, o, st);
      }, 1, "pre (scale+concat+layout)");
    }
    // ---- time / added-id embeddings (unet_plucker.py:384-414)
    {
      float* ts = tsteps; __half* s16 = small16; int B_ = B, add_dim = c.add_dim, c0 = c.boc[0];
      push([=](cudaStream_t st) {  // ts[0..B) = timestep is written by set_step_args before the plan runs
        const CallArgs& a = u->args;
        int rc = timestep_embed(ts, B_, c0, s16, st);
        if (rc) return rc;
        return timestep_embed(a.added, B_ * 3, add_dim, s16 + kSmall, st);  // [B*3, 256] == [B, 768]
      }, 2, "timestep embeddings");
      GemmEpilogue e1; e1.out = emb; e1.out_fp16 = 0;
      linear(small16, B, c.boc[0], "time_embedding.linear_1", c.temb_dim, e1);
      { float* x = emb; __half* o = small16 + 2 * kSmall; long long n = (long long)B * c.temb_dim;
        push([=](cudaStream_t st) { return silu_f16(x, o, n, st); }, 1, "silu"); }
      GemmEpilogue e2; e2.out = emb; e2.out_fp16 = 0;
      linear(small16 + 2 * kSmall, B, c.temb_dim, "time_embedding.linear_2", c.temb_dim, e2);
      GemmEpilogue e3; e3.out = emb2; e3.out_fp16 = 0;
      linear(small16 + kSmall, B, 3 * c.add_dim, "add_embedding.linear_1", c.temb_dim, e3);
      { float* x = emb2; __half* o = small16 + 2 * kSmall; long long n = (long long)B * c.temb_dim;
        push([=](cudaStream_t st) { return silu_f16(x, o, n, st); }, 1, "silu"); }
      GemmEpilogue e4; e4.out = emb; e4.out_fp16 = 0; e4.res1 = emb;  // emb = time_emb + add_emb
      linear(small16 + 2 * kSmall, B, c.temb_dim, "add_embedding.linear_2", c.temb_dim, e4);
      // every time_emb_proj(SiLU(emb)) of the 44 res blocks in one GEMM
      { float* x = emb; __half* o = small16 + 3 * kSmall; long long n = (long long)B * c.temb_dim;
        push([=](cudaStream_t st) { return silu_f16(x, o, n, st); }, 1, "silu"); }
      GemmEpilogue e5; e5.out = temb_all; e5.out_fp16 = 0;
      linear(small16 + 3 * kSmall, B, c.temb_dim, "temb_proj_all", c.temb_total, e5);
      // every single-key cross attention: to_out(to_v(ehs_b)) + bias, folded weights, one GEMM
      { __half* o = small16 + 4 * kSmall; long long n = (long long)B * c.cross_dim;
        push([=](cudaStream_t st) { return cast_f16(u->args.ehs, o, n, st); }, 1, "cast"); }
      GemmEpilogue e6; e6.out = xattn_all; e6.out_fp16 = 0;
      linear(small16 + 4 * kSmall, B, c.cross_dim, "xattn_all", c.xattn_total, e6);
    }
    // ---- conv_in
    std::vector<std::pair<float*, int>> skips;  // (tensor, channels)
    float* s0 = bump.take<float>(lM[0] * c.boc[0]);
    conv3x3_simple(in16, BF, lh[0], lw[0], c.cin_pad, "conv_in", c.boc[0], s0);
    skips.push_back({s0, c.boc[0]});
    const float* x = s0;
    int xC = c.boc[0];
    // ---- down blocks
    for (int i = 0; i < 4; ++i) {
      const int Cout = c.boc[i];
      const float eps = c.down_attn[i] ? c.eps_cross : c.eps_plain;
      const std::string bp = "down_blocks." + std::to_string(i);
      for (int j = 0; j < c.layers; ++j) {
        float* o = bump.take<float>(lM[i] * Cout);
        resblock(bp + ".resnets." + std::to_string(j), i, x, xC, nullptr, 0, Cout, eps, o);
        if (c.down_attn[i]) transformer(bp + ".attentions." + std::to_string(j), i, o, Cout, c.heads[i], o);
        skips.push_back({o, Cout});
        x = o; xC = Cout;
      }
      if (i < 3) {
        // Downsample2D: conv 3x3 stride 2 pad 1 on the 4 phase images
        { const float* xi = x; __half* o = resamp16; long long n = BF; int hh = lh[i], ww = lw[i], C = xC;
          push([=](cudaStream_t st) { return downsplit(xi, o, n, hh, ww, C, st); }, 1, "downsplit"); }
        float* o = bump.take<float>(lM[i + 1] * Cout);
        GemmProblem pr;
        pr.a0 = resamp16; pr.w = W(bp + ".downsamplers.0.conv.weight");
        pr.B = BF; pr.T = 1; pr.Tmap = 4; pr.Y = lh[i + 1]; pr.X = lw[i + 1]; pr.C0 = xC; pr.N = Cout; pr.K_total = 9LL * xC;
        pr.num_taps = 9;
        int t = 0;
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx, ++t) {
            pr.tap_dx[t] = (int8_t)(kx == 0 ? -1 : 0);
            pr.tap_dy[t] = (int8_t)(ky == 0 ? -1 : 0);
            pr.tap_dt[t] = (int8_t)((ky != 1 ? 2 : 0) + (kx != 1 ? 1 : 0));  // phase image index
            pr.tap_src[t] = 0;
          }
        pr.ep.out = o; pr.ep.out_fp16 = 0; pr.ep.bias = Wf(bp + ".downsamplers.0.conv.bias");
        gemm(pr, bp + ".downsamplers.0.conv");
        skips.push_back({o, Cout});
        x = o; xC = Cout;
      }
    }
    // ---- mid block
    {
      const int C = c.boc[3];
      resblock("mid_block.resnets.0", 3, x, xC, nullptr, 0, C, c.eps_mid, pp[0]);
      transformer("mid_block.attentions.0", 3, pp[0], C, c.heads[3], pp[0]);
      resblock("mid_block.resnets.1", 3, pp[0], C, nullptr, 0, C, c.eps_mid, pp[1]);
      x = pp[1]; xC = C;
    }
    int cur = 1;  // x lives in pp[cur]
    // ---- up blocks
    for (int i = 0; i < 4; ++i) {
      const int lvl = 3 - i;
      const int Cout = c.boc[lvl];
      const std::string bp = "up_blocks." + std::to_string(i);
      for (int j = 0; j < c.layers + 1; ++j) {
        auto sk = skips.back();
        skips.pop_back();
        float* o = pp[cur ^ 1];
        resblock(bp + ".resnets." + std::to_string(j), lvl, x, xC, sk.first, sk.second, Cout, c.eps_up, o);
        if (c.down_attn[lvl]) transformer(bp + ".attentions." + std::to_string(j), lvl, o, Cout, c.heads[lvl], o);
        cur ^= 1;
        x = o; xC = Cout;
      }
      if (i < 3) {
        // Upsample2D: nearest x2 then conv 3x3.  Where the geometry allows strided TMA stores: four 2x2 phase convolutions of
        // the low-resolution tensor (K = 4C instead of 9C on a 4x larger input; the fp16 operand is a plain cast of x);
        // else the literal form on the up-sampled fp16 image.
        EVW_CHECK_ARG(lh[lvl - 1] == 2 * lh[lvl] && lw[lvl - 1] == 2 * lw[lvl],
                      "latent size %dx%d must be divisible by 8 (three x2 resampling stages)", P.h, P.w);
        float* o = pp[cur ^ 1];
        if (gemm_strided_out_ok(lw[lvl], lh[lvl], Cout)) {
          { const float* xi = x; __half* o16 = resamp16; long long n = lM[lvl] * xC;
            push([=](cudaStream_t st) { return cast_f16(xi, o16, n, st); }, 1, "cast (upsampler operand)"); }
          upconv2x(resamp16, BF, lh[lvl], lw[lvl], xC, bp + ".upsamplers.0.conv", Cout, o);
        } else {
          { const float* xi = x; __half* o16 = resamp16; long long n = BF; int hh = lh[lvl], ww = lw[lvl], C = xC;
            push([=](cudaStream_t st) { return upsample2x(xi, o16, n, hh, ww, C, st); }, 1, "upsample2x"); }
          conv3x3_simple(resamp16, BF, lh[lvl - 1], lw[lvl - 1], xC, bp + ".upsamplers.0.conv", Cout, o);
        }
        cur ^= 1;
        x = o;
      }
    }
    // ---- out
    if (c.split) {
      // conv_out in split precision: 9 taps on the head + 9 on the tail of the normalised input, weight tail in the
      // padded output rows [Co, 2 Co) (added back by the post kernel)
      gnorm(x, 0, xC, nullptr, 0, BF, lS[0], 1e-5f, "conv_norm_out", 1, n16, nullptr, lo16);
      GemmProblem pr;
      pr.a0 = n16; pr.a1 = lo16; pr.w = W("conv_out.weight");
      pr.B = 1; pr.T = BF; pr.Y = lh[0]; pr.X = lw[0]; pr.C0 = xC; pr.C1 = xC; pr.N = c.cout_pad; pr.K_total = 18LL * xC;
      taps_conv3x3(pr);
      for (int i = 0; i < 9; ++i) {
        pr.tap_dx[9 + i] = pr.tap_dx[i]; pr.tap_dy[9 + i] = pr.tap_dy[i]; pr.tap_dt[9 + i] = 0; pr.tap_src[9 + i] = 1;
      }
      pr.num_taps = 18;
      pr.ep.out = y32; pr.ep.out_fp16 = 0; pr.ep.bias = Wf("conv_out.bias");
      gemm(pr, "conv_out");
    } else {
      gnorm(x, 0, xC, nullptr, 0, BF, lS[0], 1e-5f, "conv_norm_out", 1, n16, nullptr);
      conv3x3_simple(n16, BF, lh[0], lw[0], xC, "conv_out", c.cout_pad, y32);
    }
    {
      const float* y = y32; int Co = c.out_channels, Np = c.cout_pad, T_ = T, B_ = B, fold = c.split; const float* da = dargs;
      push([=](cudaStream_t st) {
        const CallArgs& a = u->args;
        if (a.mode == 1) return post_cfg_euler(y, T_, Co, HW, Np, fold, da, a.latents_out, st);
        return nhwc_to_nchw_f32(y, (long long)B_ * T_, Co, HW, Np, fold, a.out, st);
      }, 1, "post (CFG+Euler / layout)");
    }
    if (!fail.empty()) {
      set_error("evw_unet plan: %s", fail.c_str());
      return EVW_ERR_STATE;
    }
    return EVW_OK;
  }
};

int ensure_plan(UNet* U, int B, int T, int h, int w, void* ws, long long ws_bytes) {
  if (U->plan && U->plan->B == B && U->plan->T == T && U->plan->h == h && U->plan->w == w && U->plan->ws == ws) return EVW_OK;
  EVW_CHECK_ARG(B >= 1 && B <= 8 && T >= 1 && T <= 32 && h >= 8 && w >= 8, "evw_unet: unsupported shape B=%d T=%d h=%d w=%d", B, T, h, w);
  EVW_CHECK_ARG(((uintptr_t)ws & 1023) == 0, "evw_unet: workspace must be 1024-byte aligned");
  auto plan = std::make_unique<Plan>();
  plan->B = B; plan->T = T; plan->h = h; plan->w = w; plan->ws = ws; plan->ws_bytes = ws_bytes;
  {
    Plan sizing = *plan;
    Builder dry(*U, sizing, nullptr, true);
    int rc = dry.build();
    if (rc) return rc;
    if (dry.bump.off > ws_bytes) {
      set_error("evw_unet: workspace %lld bytes < required %lld", ws_bytes, dry.bump.off);
      return EVW_ERR_WORKSPACE;
    }
  }
  Builder b(*U, *plan, ws, false);
  int rc = b.build();
  if (rc) return rc;
  U->plan = std::move(plan);
  return EVW_OK;
}

int debug_stats(const void* p, long long n, int fp16, cudaStream_t st, double* absmax, long long* bad) {
  std::vector<char> host((size_t)n * (fp16 ? 2 : 4));
  if (cudaMemcpyAsync(host.data(), p, host.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
  cudaStreamSynchronize(st);
  double am = 0;
  long long b = 0;
  for (long long i = 0; i < n; ++i) {
    float v = fp16 ? __half2float(reinterpret_cast<const __half*>(host.data())[i]) : reinterpret_cast<const float*>(host.data())[i];
    if (!(v == v) || v > 3.0e38f || v < -3.0e38f) ++b;
    else if (fabs(v) > am) am = fabs(v);
  }
  *absmax = am;
  *bad = b;
  return 0;
}

// EVW_UNET_PROFILE=<path>: time every op with CUDA events (serialised) and append "index ms label" lines
int run_plan_profiled(UNet* U, cudaStream_t st, const char* path) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  FILE* f = fopen(path, "a");
  size_t i = 0;
  for (auto& op : U->plan->ops) {
    cudaEventRecord(e0, st);
    int rc = op(st);
    if (rc) return rc;
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (f) fprintf(f, "%zu %.6f %s\n", i, ms, U->plan->meta[i].label.empty() ? "-" : U->plan->meta[i].label.c_str());
    ++i;
  }
  if (f) {
    fprintf(f, "END\n");
    fclose(f);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return EVW_OK;
}

// Replay the plan as ONE CUDA graph launch.  The first call of a plan runs eagerly (it also performs the one-time
// cudaFuncSetAttribute calls of the kernels); the second call captures the ~550 launches on an internal stream (the
// caller's stream may be the legacy default stream, which cannot be captured) and from then on every call with the same
// caller pointers is: set_step_args kernel + cudaGraphLaunch on the caller's stream.  EVW_UNET_GRAPH=0 disables.
int run_plan_graph(UNet* U, cudaStream_t st, bool* done) {
  Plan& P = *U->plan;
  *done = false;
  static const bool enabled = [] { const char* e = getenv("EVW_UNET_GRAPH"); return !(e && atoi(e) == 0); }();
  if (!enabled || !P.graphs_ok) return EVW_OK;
  if (!P.warm) {
    P.warm = true;
    return EVW_OK;
  }
  const CallArgs& a = U->args;
  GraphKey key;
  key.mode = a.mode; key.sample = a.sample; key.latents = a.latents; key.cond = a.cond; key.latents_out = a.latents_out;
  key.out = a.out; key.ehs = a.ehs; key.added = a.added;
  cudaGraphExec_t exec = nullptr;
  for (auto& g : P.graphs)
    if (g.first == key) exec = g.second;
  if (!exec) {
    if (P.graphs.size() >= 8) {  // callers that hand over fresh buffers every call: stay eager
      P.graphs_ok = false;
      return EVW_OK;
    }
    if (!P.capture_stream) EVW_CUDA(cudaStreamCreateWithFlags(&P.capture_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    EVW_CUDA(cudaStreamBeginCapture(P.capture_stream, cudaStreamCaptureModeThreadLocal));
    int rc = EVW_OK;
    for (auto& op : P.ops)
      if ((rc = op(P.capture_stream))) break;
    cudaError_t e = cudaStreamEndCapture(P.capture_stream, &graph);
    if (rc || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      P.graphs_ok = false;  // fall back to eager launches for this plan
      return rc;
    }
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      cudaGetLastError();
      P.graphs_ok = false;
      return EVW_OK;
    }
    P.graphs.emplace_back(key, exec);
  }
  EVW_CUDA(cudaGraphLaunch(exec, st));
  ++P.graph_replays;
  *done = true;
  return EVW_OK;
}

int run_plan(UNet* U, cudaStream_t st) {
  static const bool debug = getenv("EVW_UNET_DEBUG") != nullptr;
  {
    const CallArgs& a = U->args;
    int rc = set_step_args(U->plan->tsteps, U->plan->B, U->plan->dargs, a.timestep, a.sigma, a.sigma_next, a.g_min, a.g_max, st);
    if (rc) return rc;
  }
  if (const char* prof = getenv("EVW_UNET_PROFILE")) return run_plan_profiled(U, st, prof);
  if (!debug) {
    bool done = false;
    int rc = run_plan_graph(U, st, &done);
    if (rc || done) return rc;
  }
  size_t i = 0;
  for (auto& op : U->plan->ops) {
    int rc = op(st);
    if (rc) return rc;
    if (debug) {
      const OpMeta& m = U->plan->meta[i];
      cudaError_t e = cudaStreamSynchronize(st);
      double amax = 0;
      long long bad = 0;
      if (e == cudaSuccess && m.out && m.n > 0) debug_stats(m.out, m.n, m.fp16, st, &amax, &bad);
      fprintf(stderr, "[evw_unet] op %4zu %-70s %s absmax %.4g nonfinite %lld\n", i, m.label.c_str(),
              e == cudaSuccess ? "ok" : cudaGetErrorString(e), amax, bad);
      if (const char* dir = getenv("EVW_UNET_DUMP_DIR")) {
        if (e == cudaSuccess && m.out && m.n > 0) {
          std::vector<char> host((size_t)m.n * (m.fp16 ? 2 : 4));
          cudaMemcpy(host.data(), m.out, host.size(), cudaMemcpyDeviceToHost);
          std::string path = std::string(dir) + "/op" + std::to_string(i) + (m.fp16 ? ".f16" : ".f32");
          if (FILE* f = fopen(path.c_str(), "wb")) {
            fwrite(host.data(), 1, host.size(), f);
            fclose(f);
          }
          if (FILE* f = fopen((std::string(dir) + "/index.txt").c_str(), "a")) {
            fprintf(f, "%zu %s %lld %d\n", i, m.label.empty() ? "-" : m.label.c_str(), m.n, m.fp16);
            fclose(f);
          }
        }
      }
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

using evw::UNet;

extern "C" int evw_unet_create(void** handle, const int* cfg_ints, int n_ints, const float* cfg_floats, int n_floats,
                               const char* const* tensor_names, const void* const* tensor_ptrs, int n_tensors,
                               const char* const* scalar_names, const double* scalar_values, int n_scalars) {
  EVW_CHECK_ARG(handle && cfg_ints && n_ints >= 19 && cfg_floats && n_floats >= 4, "evw_unet_create: bad config arrays");
  auto* U = new UNet();
  evw::Config& c = U->cfg;
  int k = 0;
  c.in_channels = cfg_ints[k++]; c.out_channels = cfg_ints[k++];
  for (int i = 0; i < 4; ++i) c.boc[i] = cfg_ints[k++];
  for (int i = 0; i < 4; ++i) c.heads[i] = cfg_ints[k++];
  for (int i = 0; i < 4; ++i) c.down_attn[i] = cfg_ints[k++];
  c.layers = cfg_ints[k++]; c.cross_dim = cfg_ints[k++]; c.add_dim = cfg_ints[k++];
  c.temb_total = cfg_ints[k++]; c.xattn_total = cfg_ints[k++];
  c.temb_dim = c.boc[0] * 4;
  c.eps_cross = cfg_floats[0]; c.eps_plain = cfg_floats[1]; c.eps_mid = cfg_floats[2]; c.eps_up = cfg_floats[3];
  bool ok = c.in_channels <= c.cin_pad && c.out_channels <= c.cout_pad && c.layers >= 1 && c.cross_dim % 64 == 0 &&
            c.add_dim * 3 % 64 == 0;
  for (int i = 0; i < 4; ++i) ok = ok && c.boc[i] % 64 == 0 && c.boc[i] == c.heads[i] * 64;
  if (!ok) {
    delete U;
    evw::set_error("evw_unet_create: unsupported configuration (channels must be multiples of 64 with head dim 64)");
    return EVW_ERR_INVALID;
  }
  for (int i = 0; i < n_tensors; ++i) U->tensors[tensor_names[i]] = tensor_ptrs[i];
  for (int i = 0; i < n_scalars; ++i) U->scalars[scalar_names[i]] = scalar_values[i];
  {
    auto it = U->scalars.find("precision.split");
    c.split = (it != U->scalars.end() && it->second != 0.0) ? 1 : 0;
    if (c.split && (3 * c.in_channels > c.cin_pad || 2 * c.out_channels > c.cout_pad)) {
      delete U;
      evw::set_error("evw_unet_create: precision.split needs 3*in_channels <= %d and 2*out_channels <= %d", c.cin_pad, c.cout_pad);
      return EVW_ERR_INVALID;
    }
  }
  *handle = U;
  return EVW_OK;
}

extern "C" int evw_unet_destroy(void* handle) {
  delete (UNet*)handle;
  return EVW_OK;
}

extern "C" int64_t evw_unet_workspace_bytes(void* handle, int B, int T, int h, int w) {
  if (!handle) return -1;
  UNet* U = (UNet*)handle;
  evw::Plan sizing;
  sizing.B = B; sizing.T = T; sizing.h = h; sizing.w = w;
  evw::Builder dry(*U, sizing, nullptr, true);
  if (dry.build() != 0) return -1;
  return evw::align_up(dry.bump.off, 1024) + 1024;
}

extern "C" int evw_unet_forward(void* handle, const float* sample, float timestep, const float* ehs,
                                const float* added_time_ids, float* out, int B, int T, int h, int w, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  EVW_CHECK_ARG(handle && sample && ehs && added_time_ids && out && workspace, "evw_unet_forward: null pointer");
  UNet* U = (UNet*)handle;
  int rc = evw::ensure_plan(U, B, T, h, w, workspace, workspace_bytes);
  if (rc) return rc;
  evw::CallArgs& a = U->args;
  a = evw::CallArgs();
  evw::Plan& P = *U->plan;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t hw = (size_t)h * w, bf = (size_t)B * T;
  EVW_CUDA(cudaMemcpyAsync(P.sample_stage, sample, bf * U->cfg.in_channels * hw * sizeof(float), cudaMemcpyDeviceToDevice, st));
  EVW_CUDA(cudaMemcpyAsync(P.ehs_stage, ehs, (size_t)B * U->cfg.cross_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
  EVW_CUDA(cudaMemcpyAsync(P.added_stage, added_time_ids, (size_t)B * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  a.mode = 0; a.sample = P.sample_stage; a.timestep = timestep; a.ehs = P.ehs_stage; a.added = P.added_stage; a.out = P.out_stage;
  rc = evw::run_plan(U, st);
  if (rc) return rc;
  EVW_CUDA(cudaMemcpyAsync(out, P.out_stage, bf * U->cfg.out_channels * hw * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return EVW_OK;
}

extern "C" int evw_denoise_step(void* handle, float* latents, const float* cond_latents, float sigma, float sigma_next,
                                const float* ehs, const float* added_time_ids, float g_min, float g_max, int T, int h,
                                int w, void* workspace, int64_t workspace_bytes, void* stream) {
  EVW_CHECK_ARG(handle && latents && cond_latents && ehs && added_time_ids && workspace, "evw_denoise_step: null pointer");
  EVW_CHECK_ARG(sigma > 0.f, "evw_denoise_step: sigma must be positive");
  UNet* U = (UNet*)handle;
  int rc = evw::ensure_plan(U, 2, T, h, w, workspace, workspace_bytes);
  if (rc) return rc;
  evw::CallArgs& a = U->args;
  a = evw::CallArgs();
  evw::Plan& P = *U->plan;
  cudaStream_t st = (cudaStream_t)stream;
  EVW_CUDA(cudaMemcpyAsync(P.ehs_stage, ehs, (size_t)2 * U->cfg.cross_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
  EVW_CUDA(cudaMemcpyAsync(P.added_stage, added_time_ids, (size_t)2 * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  a.mode = 1; a.latents = latents; a.latents_out = latents; a.cond = cond_latents; a.sigma = sigma; a.sigma_next = sigma_next;
  a.timestep = 0.25f * logf(sigma);
  a.ehs = P.ehs_stage; a.added = P.added_stage; a.g_min = g_min; a.g_max = g_max;
  return evw::run_plan(U, st);
}

extern "C" int64_t evw_unet_graph_replays(void* handle) {
  UNet* U = (UNet*)handle;
  return (U && U->plan) ? U->plan->graph_replays : -1;
}

extern "C" int64_t evw_unet_gn_fused(void* handle) {
  UNet* U = (UNet*)handle;
  return (U && U->plan) ? U->plan->gn_fused : -1;
}

extern "C" int evw_unet_plan_info(void* handle, int64_t* launches, double* flops) {
  EVW_CHECK_ARG(handle, "evw_unet_plan_info: null handle");
  UNet* U = (UNet*)handle;
  if (!U->plan) {
    evw::set_error("evw_unet_plan_info: no plan yet (call forward first)");
    return EVW_ERR_STATE;
  }
  if (launches) *launches = U->plan->launches + 1;  // + set_step_args
  if (flops) *flops = U->plan->flops;
  return EVW_OK;
}
