// Shared machinery of the pre-planned networks (unet_host.cu: the denoise step; vae_host.cu: VAE encode / temporal
// decode): a plan is a list of closures that enqueue kernels on a stream, built once per shape over a caller-provided
// workspace (bump-allocated), with the tcgen05 implicit GEMM (tc_gemm.h) and the normalisation kernels (unet_elem.h) as the
// building blocks.  No allocation and no host synchronisation when a plan runs.
#pragma once
#include "common.h"
#include "tc_gemm.h"
#include "unet_elem.h"

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace evw {

using Op = std::function<int(cudaStream_t)>;

struct OpMeta {
  std::string label;
  const void* out = nullptr;
  long long n = 0;
  int fp16 = 0;
};

struct PlanCore {
  std::vector<OpMeta> meta;
  std::vector<Op> ops;
  long long launches = 0;
  double flops = 0;
  int gn_fused = 0;  // GroupNorms whose statistics come from the producing GEMM's epilogue
};

struct Bump {
  char* base;
  long long off = 0;
  explicit Bump(void* b) : base((char*)b) {}
  template <typename T>
  T* take(long long n) {
    off = align_up(off, 1024);
    T* p = base ? (T*)(base + off) : nullptr;
    off += n * (long long)sizeof(T);
    return p;
  }
};

using TensorMap = std::unordered_map<std::string, const void*>;
using ScalarMap = std::unordered_map<std::string, double>;

struct BuilderBase {
  const TensorMap& tensors;
  const ScalarMap& scalars;
  PlanCore& core;
  Bump bump;
  bool dry;  // size computation only
  double* stats = nullptr;  // [instances, 32, 2] GroupNorm sums (set by the derived builder)
  std::string fail;
  // GroupNorm statistics from the producer's epilogue: the most recent GEMM that wrote each tensor, and the plan position
  // of the last op that used the statistics scratch (a GroupNorm, or a GEMM already asked to fill it)
  struct Writer { std::shared_ptr<GemmOp> op; long long rows; int N; size_t idx; };
  std::map<const void*, Writer> last_writer;
  size_t stats_busy_idx = 0;

  BuilderBase(const TensorMap& t, const ScalarMap& s, PlanCore& c, void* ws, bool dry_) : tensors(t), scalars(s), core(c), bump(ws), dry(dry_) {}

  const void* W(const std::string& name) {
    auto it = tensors.find(name);
    if (it == tensors.end()) {
      if (fail.empty()) fail = "missing tensor '" + name + "'";
      return nullptr;
    }
    return it->second;
  }
  const float* Wf(const std::string& name) { return (const float*)W(name); }
  double Sc(const std::string& name) {
    auto it = scalars.find(name);
    if (it == scalars.end()) {
      if (fail.empty()) fail = "missing scalar '" + name + "'";
      return 0;
    }
    return it->second;
  }
  void push(Op op, int launches = 1, const std::string& label = "", const void* out = nullptr, long long n = 0, int fp16 = 0) {
    core.launches += launches;
    if (!dry) {
      if (out) last_writer.erase(out);
      core.ops.push_back(std::move(op));
      core.meta.push_back(OpMeta{label, out, n, fp16});
    }
  }

  // ---- planned GEMM
  void gemm(GemmProblem pr, const std::string& label = "gemm") {
    if (dry) {
      core.launches += 1;
      return;
    }
    if (!fail.empty()) return;
    auto op = std::make_shared<GemmOp>();
    int rc = gemm_plan(op.get(), pr);
    if (rc != 0) {
      fail = std::string("gemm_plan: ") + evw_last_error();
      return;
    }
    core.flops += op->flops;
    const long long rows = (long long)pr.B * pr.T * pr.Y * pr.X;
    push([op](cudaStream_t st) { return gemm_launch(*op, st); }, 1,
         label + " " + std::to_string(rows) + "x" + std::to_string(pr.N) + "x" + std::to_string(pr.K_total), pr.ep.out,
         rows * (pr.ep.geglu ? pr.N / 2 : pr.N), pr.ep.out_fp16);
    if (!pr.ep.geglu && !pr.ep.out_lo) last_writer[pr.ep.out] = Writer{op, rows, pr.N, core.ops.size()};
  }
  static void taps_conv3x3(GemmProblem& pr) {
    pr.num_taps = 9;
    int i = 0;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx, ++i) {
        pr.tap_dx[i] = (int8_t)dx; pr.tap_dy[i] = (int8_t)dy; pr.tap_dt[i] = 0; pr.tap_src[i] = 0;
      }
  }
  void linear(const __half* a, long long M, int K, const std::string& wname, int N, GemmEpilogue ep, bool bias = true) {
    GemmProblem pr;
    pr.a0 = a; pr.w = W(wname + ".weight");
    pr.X = (int)M; pr.C0 = K; pr.N = N; pr.K_total = K; pr.num_taps = 1;
    if (bias) ep.bias = Wf(wname + ".bias");
    pr.ep = ep;
    gemm(pr, wname);
  }
  void conv3x3_simple(const __half* a, int frames, int hh, int ww, int Cin, const std::string& name, int N, float* outp) {
    GemmProblem pr;
    pr.a0 = a; pr.w = W(name + ".weight");
    pr.B = 1; pr.T = frames; pr.Y = hh; pr.X = ww; pr.C0 = Cin; pr.N = N; pr.K_total = 9LL * Cin;
    taps_conv3x3(pr);
    pr.ep.out = outp; pr.ep.out_fp16 = 0; pr.ep.bias = Wf(name + ".bias");
    gemm(pr, name);
  }
  // diffusers Upsample2D (nearest x2 + Conv2d 3x3) as four 2x2 phase convolutions of the low-resolution input a16
  // [frames, hh, ww, C] -> outp fp32 [frames, 2hh, 2ww, N]; weights [4][N, 4C] packed by ops.upconv_weights
  void upconv2x(const __half* a16, int frames, int hh, int ww, int C, const std::string& name, int N, float* outp) {
    const __half* w4 = (const __half*)W(name + ".weight4");
    for (int phase = 0; phase < 4; ++phase) {
      GemmProblem pr;
      pr.a0 = a16; pr.w = (w4 && !dry) ? w4 + (size_t)phase * N * 4 * C : w4;
      pr.B = 1; pr.T = frames; pr.Y = hh; pr.X = ww; pr.C0 = C; pr.N = N;
      upconv2x_phase(pr, phase);
      pr.ep.out = outp; pr.ep.out_fp16 = 0; pr.ep.bias = Wf(name + ".bias");
      gemm(pr, name + ".phase" + std::to_string(phase));
    }
    if (!dry) last_writer.erase(outp);  // four partial writers: nobody owns the whole tensor's GroupNorm sums
  }
  void gnorm(const void* src0, int src0_fp16, int C0, const float* src1, int C1, long long insts, long long rows, float eps,
             const std::string& name, int silu, __half* out, __half* raw, __half* out_lo = nullptr) {
    const float* g = Wf(name + ".weight");
    const float* b = Wf(name + ".bias");
    double* st_ = stats;
    // The GEMM that produced src0 accumulates the statistics in its epilogue when it can (single source, the whole
    // tensor written by that GEMM, nothing else using the statistics scratch in between): 1 kernel instead of 2.
    int have = 0;
    if (!dry && !src1) {
      auto it = last_writer.find(src0);
      if (it != last_writer.end() && it->second.N == C0 && it->second.rows == insts * rows && it->second.idx > stats_busy_idx &&
          gemm_enable_gn_stats(it->second.op.get(), stats, rows) == 0) {
        have = 1;
        ++core.gn_fused;
      }
    }
    // 2 kernels (stats, apply); the statistics clear is a memset node and is not counted as a kernel launch
    push([=](cudaStream_t st) { return group_norm(src0, src0_fp16, C0, src1, C1, insts, rows, eps, g, b, silu, st_, out, raw, out_lo, st, have); },
         have ? 1 : 2, name, out, insts * rows * (C0 + C1), 1);
    if (!dry) stats_busy_idx = core.ops.size();
  }
};

}  // namespace evw
