// Host launchers of the memory-bound UNet kernels (unet_elem.cu) and the attention kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace evw {

// GroupNorm(32 groups) over `insts` instances of `rows_per_inst` rows; input = concat(src0[C0], src1[C1]) along
// channels (src1 may be null); src0 is fp32 or fp16, src1 fp32.  out fp16 [rows, C]; raw_out (optional) = fp16 copy
// of the un-normalised input; out_lo (optional) = fp16 tail of the output (value - float(half(value))) for
// split-precision consumers.  stats: double [insts, 32, 2] scratch.
int group_norm(const void* src0, int src0_fp16, int C0, const float* src1, int C1, long long insts,
               long long rows_per_inst, float eps, const float* gamma, const float* beta, int do_silu, double* stats,
               __half* out, __half* raw_out, __half* out_lo, cudaStream_t st, int have_stats = 0);
// LayerNorm over C of (x[row] + rowvec[(row / rv_div) % rv_mod]) -> fp16
int layer_norm(const float* x, const float* rowvec, long long rv_div, long long rv_mod, long long rows, int C, float eps,
               const float* gamma, const float* beta, __half* out, cudaStream_t st, float* out32 = nullptr);
// qkv fp16 [B*T*S, 3*heads*64] rows (b,t,s) -> out fp16 [B*T*S, heads*64], attention over t
int temporal_attention(const __half* qkv, __half* out, int B, int T, long long S, int heads, cudaStream_t st);
// qkv fp16 [F*S, 3*heads*64] rows (f,s) -> out fp16 [F*S, heads*64], attention over s (tcgen05 flash attention)
int spatial_attention(const __half* qkv, __half* out, int F, int S, int heads, cudaStream_t st);
int upsample2x(const float* x, __half* out, long long n, int h, int w, int C, cudaStream_t st);
int downsplit(const float* x, __half* out, long long n, int h, int w, int C, cudaStream_t st);
// split: conv_in operand in split precision inside the channel padding ([head | tail | head], 3 Cin <= Cpad);
// fold: conv_out's weight tail lives in output columns [Co, 2Co) and is added back here
// per-step scalars on the device (so a captured graph of the plan stays valid from step to step):
// tsteps[0..B) = timestep; dargs = {1/sqrt(sigma^2+1), sigma, sigma_next, g_min, g_max}
int set_step_args(float* tsteps, int B, float* dargs, float timestep, float sigma, float sigma_next, float g_min, float g_max,
                  cudaStream_t st);
int pre_concat(const float* latents, const float* cond, int Bc, int T, int Cl, int Cc, long long HW, const float* dargs, int Cpad,
               int split, __half* out, cudaStream_t st);
int nchw_to_nhwc_f16(const float* x, long long frames, int Cin, long long HW, int Cpad, int split, __half* out, cudaStream_t st);
int nhwc_to_nchw_f32(const float* y, long long frames, int Co, long long HW, int Npad, int fold, float* out, cudaStream_t st);
int post_cfg_euler(const float* y, int T, int Cl, long long HW, int Npad, int fold, const float* dargs, float* latents,
                   cudaStream_t st);
int timestep_embed(const float* t, int n, int dim, __half* out, cudaStream_t st);
int silu_f16(const float* x, __half* out, long long n, cudaStream_t st);
int cast_f16(const float* x, __half* out, long long n, cudaStream_t st);
int fill_f32(float* out, float v, int n, cudaStream_t st);
int add_f32(const float* a, const float* b, float* out, long long n, cudaStream_t st);

}  // namespace evw
