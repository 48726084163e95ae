/*
 * evoworld_b200 — C ABI of the B200-native (sm_100a) hot paths of EvoWorld.
 *
 * The reference (JiahaoPlus/EvoWorld) has no FFI and no native code; every entry point below is
 * the boundary a maintainer would bind (ctypes) underneath the named reference function.  All
 * citations are file:line relative to the reference checkout.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `h_`;
 *   - the caller owns every buffer (torch allocates them; pass tensor.data_ptr());
 *   - no hidden allocation, no host synchronisation inside, work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return value: 0 = EVW_OK, negative = error; evw_last_error() gives the message
 *     (thread-local).  The Python shims raise RuntimeError on non-zero.
 */
#ifndef EVOWORLD_B200_H_
#define EVOWORLD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVW_OK 0
#define EVW_ERR_INVALID (-1)   /* bad argument (shape, alignment, null pointer)          */
#define EVW_ERR_CUDA (-2)      /* a CUDA runtime / driver call failed                     */
#define EVW_ERR_WORKSPACE (-3) /* caller-provided workspace too small                     */
#define EVW_ERR_STATE (-4)     /* handle used before it was fully set up                  */

const char* evw_last_error(void);
/* ABI version, bumped on any signature change. */
int evw_abi_version(void);
/* Fills SM count / L2 bytes / compute capability (major*10+minor) of the current device. */
int evw_device_info(int* sm_count, int64_t* l2_bytes, int* cc);

/* ------------------------------------------------------------------------------------------
 * Hot path 2 — 3D-memory reprojection
 * ---------------------------------------------------------------------------------------- */

/* Plücker embedding.  Replaces utils/plucker_embedding.py:221-255 ray_c2w_to_plucker.
 * ray [H,W,3] f32 (camera-frame unit rays), c2w [T,3,4] f32, out [T,6,H,W] f32 contiguous:
 * out[n,0:3] = R_n d, out[n,3:6] = t_n x (R_n d). */
int evw_plucker(const float* ray, const float* c2w, float* out, int T, int H, int W, void* stream);

/* Equirectangular -> perspective bilinear warp of uint8 images.
 * Replaces equilib.Equi2Pers.__call__ (pyequilib 0.5.8; call site
 * unified_loop_consistency.py:178-183,329).  equi [B,C,He,We] u8, out [B,C,Hp,Wp] u8.
 * pix2dir [B,9] f32 row-major: M = R * G * K^-1 mapping the homogeneous output pixel (x,y,1) to a
 * direction in equilib's global frame (x fwd, y right, z down); composed by the host shim in
 * float64 (evoworld_b200/equi2pers.py).  phi = asin(Mz/|M|), theta = atan2(My,Mx),
 * ui = (theta-pi) We/2pi + .5 mod We, uj = (phi-pi/2) He/pi + .5 mod He, bilinear with wrap,
 * result truncated to uint8. */
int evw_equi2pers_u8(const uint8_t* equi, const float* pix2dir, uint8_t* out, int B, int C, int He,
                     int We, int Hp, int Wp, void* stream);

/* Pure-yaw fast path of the same warp (pitch = roll = 0 — the only call EvoWorld makes, unified_loop_consistency.py:329).
 * evw_equi2pers_table tabulates, once per (Hp, Wp, fov, He, We), the source coordinates (ui, uj) of every output pixel for
 * the yaw = 0 camera (pix2dir0 [9] f32 = G K^-1; table [Hp*Wp] float2, ui left unwrapped) with the expressions above;
 * evw_equi2pers_yaw_u8 then warps B frames as a pure gather: ui = ui0 + shift_px[b] (the yaw as a longitude shift in
 * source pixels: -yaw We / 2pi for z_down = False, reduced to [-We/2, We/2]), wrapped; same bilinear sampling. */
int evw_equi2pers_table(const float* pix2dir0, float* table, int He, int We, int Hp, int Wp, void* stream);
int evw_equi2pers_yaw_u8(const uint8_t* equi, const float* table, const float* shift_px, uint8_t* out, int B, int C,
                         int He, int We, int Hp, int Wp, void* stream);

/* Depth lift.  Replaces third_party/vggt/vggt/utils/geometry.py:12-111
 * unproject_depth_map_to_point_map.  depth [S,H,W] f32, extr [S,3,4] f32 (cam-from-world),
 * intr [S,3,3] f32.  Exactly one of out_f64 [S,H,W,3] / out_f32 [S,H,W,3] may be NULL. */
int evw_lift_depth(const float* depth, const float* extr, const float* intr, double* out_f64,
                   float* out_f32, int S, int H, int W, void* stream);

/* Fused lift + pack (device-resident point memory, the caller row next to the path: unified_loop_consistency.py:352-367
 * followed by reproject_vggt_open3d_utils.py:224-292): depth [S,H,W] f32, extr [S,3,4], intr [S,3,3], images NCHW
 * [S,3,H,W] f32 in [0,1] -> out_pts4 [S*H*W] float4 {x,y,z,rgb-bits}; bit-identical to evw_lift_depth (f64) followed by
 * evw_pack_points, without the float64 intermediate. */
int evw_lift_pack_points(const float* depth, const float* extr, const float* intr, const float* images_f32,
                         float* out_pts4, int S, int H, int W, void* stream);

/* Pack a point cloud for splatting: xyz (f64 or f32, exactly one non-NULL) [N,3] and colours.
 * Colour source is either rgb_u8 [N,3] or images_f32 NCHW [S,3,H,W] in [0,1] (then N=S*H*W and
 * the byte is trunc(x*255) as reproject_vggt_open3d_utils.py:286-292 _extract_colors).
 * out: float4 per point {x,y,z, bits(r | g<<8 | b<<16)} = 16 B/point. */
int evw_pack_points(const double* xyz_f64, const float* xyz_f32, const uint8_t* rgb_u8,
                    const float* images_f32, int S, int HW, float* out_pts4, int64_t N,
                    void* stream);

/* Workspace bytes for evw_conf_select on n values. */
int64_t evw_conf_select_workspace(int64_t n);

/* Confidence-percentile filter + order-preserving compaction.
 * Replaces reproject_vggt_open3d_utils.py:294-310 _apply_confidence_filter:
 *   thr = lerp(sorted[k_lo], sorted[k_hi], gamma) (numpy 'linear' percentile; k_lo/k_hi/gamma are
 *   computed by the host shim with numpy's own scalar arithmetic), keep = conf >= thr.
 * use_threshold=0 keeps everything >= 0.0 (the conf_thres == 0.0 branch).
 * pts4_in [n] float4 -> pts4_out [<=n] float4 in original order; *out_count (device int64) = kept;
 * *out_thr (device f32) = threshold; keep_idx (optional, may be NULL) [<=n] int64 original indices. */
int evw_conf_select(const float* conf, const float* pts4_in, int64_t n, int64_t k_lo, int64_t k_hi,
                    float gamma, int use_threshold, float* pts4_out, int64_t* keep_idx,
                    int64_t* out_count, float* out_thr, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* Cube->equirect lookup table entry: (face<<28) | (row<<14) | col, or 0xFFFFFFFF for "no face".
 * Built on the host by the shim with the reference's own arithmetic
 * (reproject_vggt_open3d_utils.py:542-614) so that indices are bit-identical; see
 * evoworld_b200/reprojection.py:build_cube_lut. */

/* Z-buffer bytes needed for `views_per_pass` views of six res x res faces. */
int64_t evw_splat_workspace(int views_per_pass, int face_res);

/* Point splat + cube->equirect resolve for V target views.
 * Replaces CubemapRenderer.render_cubemaps_to_panoramas (reproject_vggt_open3d_utils.py:668-711):
 * 6 x V Open3D point renders (:617-666) followed by cube_to_equirectangular_cuda (:542-614).
 *   pts4   [n_cap] float4 {x,y,z,rgb-bits};  n_dev: device int64 count (<= n_cap) or NULL => n_cap
 *   w2c    [V,6,3,4] f32 cam-from-world per (view, face), face order front,right,back,left,top,bottom
 *   lut    [outH,outW] u32 cube->equirect table;  out [V,outH,outW,3] u8
 *   pixel = floor(fx*x/z+cx), nearest z wins, ties -> lowest point index, empty = 0.
 * zbuf workspace >= evw_splat_workspace(views_per_pass, face_res). */
int evw_splat_cubemap_equirect(const float* pts4, int64_t n_cap, const int64_t* n_dev,
                               const float* w2c, int V, int face_res, float focal, float z_near,
                               const uint32_t* lut, int outH, int outW, uint8_t* out,
                               void* zbuf_workspace, int64_t workspace_bytes, int views_per_pass,
                               void* stream);

/* Cube formulation of the same render (the fast path behind render_cubemaps_to_panoramas): w2c_front [V,3,4] f32 is
 * the cam-from-world of the FRONT face only (= inv(target_c2w)); the other five faces are the exact signed axis
 * permutations of the reference's CUBEMAP_TRANSFORMS, so each (point, view) costs one transform and the point lands
 * in the face of its major axis.  Otherwise the contract of evw_splat_cubemap_equirect; views_per_pass in {1,2,4,8},
 * outH*outW % 4 == 0.  flags (bit set):
 *   EVW_SPLAT_PRETEST     read the cell before issuing the 64-bit atomic min;
 *   EVW_SPLAT_OVERLAP     two internal streams (forked from / joined to `stream` with events): pass p runs its clear,
 *                         splat and resolve on stream p % 2, so one pass's L2-atomic-bound splat overlaps its neighbour's
 *                         gather-latency-bound resolve and its clear; the workspace must then hold
 *                         two passes: evw_splat_workspace_flags(views_per_pass, face_res, flags);
 * All flag combinations except EVW_SPLAT_COLOR_KEYS produce identical bytes.  With EVW_SPLAT_OVERLAP the call uses two
 * library-owned streams per device (created on first use, forked from / joined to `stream` with events, capturable);
 * a per-device mutex serialises host threads that enter the call concurrently for the same device (the enqueue is
 * short; the reference's process model is one Python thread per process and GPU anyway). */
#define EVW_SPLAT_PRETEST 1
#define EVW_SPLAT_OVERLAP 2
#define EVW_SPLAT_COLOR_KEYS 16 /* OPTIONAL tie rule, off by default: the low key word carries colour << 8 | (index mod 256)
                                  instead of the index, so the resolve needs no gather (faster).  Output identical to the
                                  default unless two points of different colour have exactly the same float32 depth in one
                                  cell: the default keeps the lowest index (draw order), this keeps the lowest packed colour */
#define EVW_SPLAT_OVERLAP_BY_ROLE 8 /* with OVERLAP: one stream runs every splat, the other every clear + resolve (measured slower) */
int64_t evw_splat_workspace_flags(int views_per_pass, int face_res, int flags);
/* Tuning hook: resident splat CTAs per SM (1..8) while EVW_SPLAT_OVERLAP is set; fewer leaves SM slots for the
 * neighbouring pass's resolve and clear.  0 restores the default (EVW_SPLAT_CTAS_PER_SM or 8). */
void evw_set_splat_ctas_per_sm(int ctas);
int evw_splat_cube_equirect(const float* pts4, int64_t n_cap, const int64_t* n_dev, const float* w2c_front, int V,
                            int face_res, float focal, float z_near, const uint32_t* lut, int outH, int outW,
                            uint8_t* out, void* zbuf_workspace, int64_t workspace_bytes, int views_per_pass,
                            int flags, void* stream);
int evw_splat_cube_faces_debug(const float* pts4, int64_t n, const float* w2c_front, int V, int face_res, float focal,
                               float z_near, int64_t* win_idx, void* zbuf_workspace, int64_t workspace_bytes,
                               void* stream);

/* Same splat, but returns the raw per-face winners for parity tests:
 * win_idx [V,6,res,res] int64 (-1 = empty). */
int evw_splat_faces_debug(const float* pts4, int64_t n, const float* w2c, int V, int face_res,
                          float focal, float z_near, int64_t* win_idx, void* zbuf_workspace,
                          int64_t workspace_bytes, void* stream);

/* Per-face renders (API completeness for CubemapRenderer.render_face / render_cubemap,
 * reproject_vggt_open3d_utils.py:617-666): faces_out [V,6,res,res,3] u8 (HWC per face). */
int evw_splat_faces_u8(const float* pts4, int64_t n, const float* w2c, int V, int face_res,
                       float focal, float z_near, uint8_t* faces_out, void* zbuf_workspace,
                       int64_t workspace_bytes, void* stream);

/* cube_to_equirectangular_cuda (reproject_vggt_open3d_utils.py:542-614) on already rendered faces:
 * faces [B,6,3,res,res] u8 (CHW per face, face order as above), out [B,outH,outW,3] u8. */
int evw_cube_to_equirect_u8(const uint8_t* faces, const uint32_t* lut, int B, int face_res,
                            int outH, int outW, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hot path 1 — UNet denoise step (tensor-core kernels)
 * ---------------------------------------------------------------------------------------- */

/* Generic tcgen05 implicit GEMM (csrc/tc_gemm.cu); the contraction behind every nn.Linear,
 * Conv2d 3x3 (+1x1 shortcut) and Conv3d (3,1,1) of the diffusers SpatioTemporal UNet blocks that
 * evoworld/trainer/unet_plucker.py:163-244 instantiates.
 *   a0 fp16 [B,T,Y,X,C0] channels-last, a1 optional fp16 [B,T,Y,X,C1], w fp16 [N, K_total] tap-major,
 *   h_taps (HOST) int8 [num_taps,4] = (dx,dy,dt,src); reads outside the tensor are zero (padding).
 *   out[row,n] = s0*(acc+bias[n]) + rowvec[(row/rv_div)%rv_mod, n] + s1*res1[row,n] + s2*res2[row,n];
 *   geglu = 1: columns interleaved [16 value|16 gate], out has N/2 columns = (v+b)*gelu(g+b).
 *   geglu = 2: out = gelu(s0 * (acc + bias)) (exact erf form; no rowvec / res1 / res2 / out_lo): the fc1 of ViT-style MLPs.
 *   out/res1 are fp16 or fp32 (flags), bias/rowvec/res2 fp32.  C0, C1 multiples of 64; N multiple of 8.
 *   out_lo (optional, fp16 output only, N multiple of 16): fp16 tail half(v - float(half(v))) of every stored value —
 *   the split-precision operand format: a consumer reads head and tail as two tap sources (a0, a1) against
 *   [W_hi | W_hi | W_lo] and keeps ~22 significant bits of the product (used for proj_in / proj_out / conv_out,
 *   the low-FLOP layers that dominate the fp16 rounding error of the step: tools/precision_sim.py). */
int evw_gemm_f16(const void* a0, const void* a1, const void* w, int B, int T, int Y, int X, int C0, int C1,
                 int N, int num_taps, const int8_t* h_taps, void* out, int out_fp16, const float* bias,
                 const float* rowvec, int64_t rv_div, int64_t rv_mod, const void* res1, int res1_fp16, float s1,
                 const float* res2, float s2, float s0, int geglu, int block_n, void* out_lo, void* stream);

/* evw_gemm_f16 whose epilogue also accumulates the GroupNorm(32) statistics of its output, so that the GroupNorm that
 * consumes it (ResnetBlock2D.norm2 after conv1, TemporalResnetBlock.norm1 after the spatial block, ... — diffusers
 * resnet.py) needs no statistics pass of its own: gn_stats double [rows / gn_rows_per_inst, 32, 2] = per (instance, group)
 * sum and sum of squares of the stored values (cleared by the call).  Only for epilogues without GEGLU / res2 and with
 * res1 fp32 (or absent); all rows of a 128-row tile must fall into one instance — EVW_ERR_INVALID otherwise. */
int evw_gemm_f16_gn(const void* a0, const void* a1, const void* w, int B, int T, int Y, int X, int C0, int C1,
                    int N, int num_taps, const int8_t* h_taps, void* out, int out_fp16, const float* bias,
                    const float* rowvec, int64_t rv_div, int64_t rv_mod, const void* res1, int res1_fp16, float s1,
                    const float* res2, float s2, float s0, int geglu, int block_n, void* out_lo, double* gn_stats,
                    int64_t gn_rows_per_inst, void* stream);

/* diffusers Upsample2D (nearest x2, then Conv2d 3x3 padding 1) without materialising the up-sampled image: every output phase
 * (2y + py, 2x + px) is a 2x2 convolution of the low-resolution input with the 3x3 weights that land on the same source pixel
 * summed — four GEMMs of K = 4C that store through strided TMA maps.  a fp16 [F,h,w,C] -> out fp32 [F,2h,2w,N]; w4 fp16
 * [4][N,4C] (evoworld_b200/ops.py::upconv_weights).  Used by the UNet up blocks (unet_plucker.py:203-233 -> diffusers
 * UpBlockSpatioTemporal.upsamplers) and the VAE decoder. */
int evw_upconv2x_f16(const void* a, const void* w4, const float* bias, void* out, int F, int h, int w, int C, int N,
                     void* stream);

/* Launch mode of the GEMM (takes effect when an op is planned): 1 = always CTA pairs (clusters of two) on m-adjacent tiles
 * running tcgen05.mma.cta_group::2 with M = 256, each CTA holding half of the weight tile; 0 = always independent CTAs
 * (cta_group::1, M = 128); -1 = default (EVW_GEMM_CLUSTER, else automatic: pairs where K_total >= EVW_GEMM_PAIR_MIN_K,
 * 1024 by default — the main-loop-bound convolutions and wide linears — and single CTAs for the epilogue-bound K = 320 /
 * 640 linears).  All modes are bit-identical. */
void evw_set_gemm_cluster(int on);
/* Whether the UNet plan lets GEMM epilogues accumulate the statistics of the GroupNorm that follows (evw_gemm_f16_gn):
 * 1 = yes, 0 = every GroupNorm runs its own statistics pass, -1 = default (EVW_GEMM_GN_STATS, on).  Read at plan time. */
void evw_set_gemm_gn_stats(int on);
/* Output path of the GEMM epilogue (read at plan time): 1 = TMA stores from per-warp shared-memory slabs wherever the
 * geometry allows (default), 0 = per-thread global stores only, -1 = default (EVW_GEMM_STORE_TMA).  Bit-identical. */
void evw_set_gemm_store_tma(int mode);

/* Spatial self-attention (BasicTransformerBlock.attn1 -> F.scaled_dot_product_attention, head dim 64):
 * qkv fp16 [F*S, 3*heads*64] (columns [q|k|v], each [heads,64]) -> out fp16 [F*S, heads*64]; softmax over
 * the S tokens of each frame.  tcgen05 flash-attention forward (csrc/tc_attention.cu). */
int evw_spatial_attention_f16(const void* qkv, void* out, int F, int S, int heads, void* stream);
/* Micro-benchmark hook: pick the spatial-attention kernel variant at run time.  -2 = default (EVW_ATTN_VARIANT, else 0);
 * 0 = v8 (the default kernel), 1 = v8 with every 8th exponential on the FMA pipe, 2 = v8 with staggered groups,
 * 3 = v7 (one thread per query row), 4 = v7 with every 4th exponential on the FMA pipe (csrc/tc_attention.cu). */
void evw_set_attention_variant(int variant);

/* Temporal self-attention (TemporalBasicTransformerBlock.attn1): same qkv layout with rows ordered
 * (b, t, s); softmax over the T <= 32 frames of each (b, s, head). */
int evw_temporal_attention_f16(const void* qkv, void* out, int B, int T, int64_t S, int heads, void* stream);

/* GroupNorm(32) [+SiLU] over `insts` instances of `rows_per_inst` channels-last rows; input is the channel
 * concatenation of src0 (fp32 or fp16, C0 ch) and optional src1 (fp32, C1 ch); out fp16 [rows, C0+C1];
 * raw_out (optional) receives the un-normalised fp16 copy; out_lo (optional) the fp16 tail of the output (split-
 * precision operand, see evw_gemm_f16); stats_ws >= insts*64 doubles. */
int evw_group_norm_f16(const void* src0, int src0_fp16, int C0, const float* src1, int C1, int64_t insts,
                       int64_t rows_per_inst, float eps, const float* gamma, const float* beta, int do_silu,
                       void* stats_ws, void* out, void* raw_out, void* out_lo, void* stream);

/* LayerNorm over C of x[row] (+ rowvec[(row/rv_div)%rv_mod]) -> fp16 [rows, C]. */
int evw_layer_norm_f16(const float* x, const float* rowvec, int64_t rv_div, int64_t rv_mod, int64_t rows, int C,
                       float eps, const float* gamma, const float* beta, void* out, void* stream);

/* The same LayerNorm with an fp32 result (CLIP's pre_layrnorm produces the fp32 residual stream of the encoder). */
int evw_layer_norm_f32(const float* x, int64_t rows, int C, float eps, const float* gamma, const float* beta, float* out,
                       void* stream);

/* UNetSpatioTemporalConditionModel on the device (evoworld/trainer/unet_plucker.py:30-488).
 * evw_unet_create takes the packed parameters by name (see evoworld_b200/unet.py:pack_parameters for
 * the naming and layouts; the caller keeps the tensors alive) plus named host scalars (AlphaBlender
 * alphas, offsets into the batched time_emb_proj / cross-attention tables).
 *   cfg_ints  = {in_ch, out_ch, boc[4], heads[4], down_attn[4], layers_per_block, cross_dim,
 *                addition_time_embed_dim, temb_total, xattn_total};  cfg_floats = {eps_cross, eps_plain, eps_mid, eps_up} */
int evw_unet_create(void** handle, const int* cfg_ints, int n_ints, const float* cfg_floats, int n_floats,
                    const char* const* tensor_names, const void* const* tensor_ptrs, int n_tensors,
                    const char* const* scalar_names, const double* scalar_values, int n_scalars);
int evw_unet_destroy(void* handle);
/* Bytes of caller-provided workspace needed for a [B,T,*,h,w] call (-1 on error). */
int64_t evw_unet_workspace_bytes(void* handle, int B, int T, int h, int w);
/* forward (unet_plucker.py:355-488): sample fp32 [B,T,Cin,h,w], timestep (= 0.25 ln sigma), encoder_hidden_states
 * fp32 [B,1,cross_dim], added_time_ids fp32 [B,3] -> out fp32 [B,T,Cout,h,w].  workspace 1024-B aligned. */
int evw_unet_forward(void* handle, const float* sample, float timestep, const float* ehs, const float* added_time_ids,
                     float* out, int B, int T, int h, int w, void* workspace, int64_t workspace_bytes, void* stream);
/* One iteration of the denoise loop (pipeline_evoworld.py:689-725), in place on `latents` fp32 [1,T,4,h,w]:
 * cat([latents]*2)/sqrt(sigma^2+1) ++ cond_latents fp32 [2,T,Cin-4,h,w] -> UNet -> CFG with guidance
 * linspace(g_min,g_max,T) -> Euler v-prediction step sigma -> sigma_next. */
int evw_denoise_step(void* handle, float* latents, const float* cond_latents, float sigma, float sigma_next,
                     const float* ehs, const float* added_time_ids, float g_min, float g_max, int T, int h, int w,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* How many calls on the current plan were served by replaying its captured CUDA graph (one cudaGraphLaunch instead of
 * ~550 kernel launches; the first call of a plan runs eagerly, the second captures).  -1 without a plan.  EVW_UNET_GRAPH=0
 * disables graph replay. */
int64_t evw_unet_graph_replays(void* handle);
/* How many GroupNorms of the current plan take their statistics from the epilogue of the GEMM that produced their input
 * (evw_gemm_f16_gn) instead of a pass of their own.  -1 without a plan.  EVW_GEMM_GN_STATS=0 disables the fusion. */
int64_t evw_unet_gn_fused(void* handle);
/* Kernel launches and algorithmic FLOPs of the current plan (after the first forward / step). */
int evw_unet_plan_info(void* handle, int64_t* launches, double* flops);

/* Pillow-exact bilinear resize of 8-bit images: `transforms.Resize((height, width))` on the PIL copies of the reprojected
 * memory panoramas (dataset/CameraTrajDataset.py:597-600, applied at unified_loop_consistency.py:422) = PIL.Image.resize
 * (BILINEAR): horizontal pass into an 8-bit intermediate, then vertical, 22-bit fixed-point coefficients (libImaging
 * Resample.c).  in uint8 [N,H,W,3] -> out uint8 [N,h,w,3]; tmp uint8 [N,H,w,3]; the coefficient tables come from
 * evoworld_b200/image_ops.py::pil_resize_tables (device int32: bounds [out,2] = first input index and tap count, k [out,ks]). */
int evw_resize_pil_u8(const uint8_t* in, uint8_t* tmp, uint8_t* out, int N, int H, int W, int h, int w, const int* bounds_x,
                      const int* kx, int ksx, const int* bounds_y, const int* ky, int ksy, void* stream);

/* CLIP ViT image encoder pieces (transformers CLIPVisionModelWithProjection, called at
 * evoworld/pipeline/pipeline_evoworld.py:289; transformers models/clip/modeling_clip.py CLIPAttention / CLIPMLP).  The
 * linears use evw_gemm_f16 and the LayerNorms evw_layer_norm_f16 (evoworld_b200/clip.py).
 * evw_small_attention_f16: softmax(scale q k^T) v per (image, head) over S <= 1024 tokens with any head width <= 256
 * (ViT-H: S = 257, 16 heads of 80): qkv fp16 [B*S, 3*heads*head_dim] (q | k | v) -> out fp16 [B*S, heads*head_dim].
 * evw_act_f16: x fp32 -> fp16 through an activation, mode 0 = GELU (erf), 1 = quick_gelu, 2 = ReLU, 3 = identity (cast),
 * 4 = SiLU (2-4: VGGT's DPT residual units and camera-head modulation). */
int evw_small_attention_f16(const void* qkv, void* out, int B, int S, int heads, int head_dim, float scale, void* stream);
int evw_act_f16(const float* x, void* out, int64_t n, int mode, void* stream);

/* VAE around the denoise loop: diffusers AutoencoderKLTemporalDecoder as the pipeline calls it —
 * `vae.encode(image).latent_dist.mode()` (evoworld/pipeline/pipeline_evoworld.py:307-328, call sites :610-617) and
 * `vae.decode(latents / scaling_factor, num_frames=n).sample` (decode_latents, :358-385, call site :731).
 * evw_vae_create takes the packed parameters by name (evoworld_b200/vae.py:pack_parameters: fp16 tap-major convolution
 * weights, conv_shortcut merged into conv2, quant_conv folded into encoder.conv_out, to_v's bias folded into to_out's)
 * plus named host scalars (AlphaBlender alphas); cfg_ints = {in_channels, out_channels, latent_channels,
 * block_out_channels[4], layers_per_block}.  The caller keeps the tensors alive. */
int evw_vae_create(void** handle, const int* cfg_ints, int n_ints, const char* const* tensor_names,
                   const void* const* tensor_ptrs, int n_tensors, const char* const* scalar_names,
                   const double* scalar_values, int n_scalars);
int evw_vae_destroy(void* handle);
/* Workspace of one call: mode 0 = encode N images of H x W, mode 1 = decode N latents of H x W (N a multiple of
 * num_frames); -1 on error. */
int64_t evw_vae_workspace_bytes(void* handle, int mode, int N, int num_frames, int H, int W);
/* images fp32 [N,3,H,W] in [-1,1] -> moments fp32 [N, 2*latent, H/8, W/8] = (mean | logvar) of the posterior after
 * quant_conv (DiagonalGaussianDistribution.mode() == mean).  H, W divisible by 8, (H/8)*(W/8) a multiple of 64. */
int evw_vae_encode(void* handle, const float* images, float* moments, int N, int H, int W, void* workspace,
                   int64_t workspace_bytes, void* stream);
/* latents fp32 [N, latent, h, w] (already divided by scaling_factor), N = videos x num_frames -> frames fp32 [N,3,8h,8w]:
 * TemporalDecoder incl. the (3,1,1) time_conv_out over each video's frames.  h*w a multiple of 64. */
int evw_vae_decode(void* handle, const float* latents, float* frames, int N, int num_frames, int h, int w,
                   void* workspace, int64_t workspace_bytes, void* stream);
/* Kernel launches, algorithmic FLOPs and epilogue-fused GroupNorms of the current encode (0) / decode (1) plan. */
int evw_vae_plan_info(void* handle, int mode, int64_t* launches, double* flops, int64_t* gn_fused);

/* VGGT forward pieces (third_party/vggt/vggt/models/vggt.py:56-92 as called at
 * unified_loop_consistency.py:114-136).  Linears / convolutions: evw_gemm_f16; frame and global attention (head width 64):
 * evw_spatial_attention_f16; camera trunk attention: evw_small_attention_f16; LayerNorms: evw_layer_norm_f16 / _f32
 * (evoworld_b200/vggt.py).
 * evw_qknorm_rope_f16: layers/attention.py:54-58 in place on qkv fp16 [rows, 3*heads*64] — LayerNorm(64, eps) of every q and
 *   k head (gamma / beta fp32 [64]) followed by the 2-D rotary embedding (layers/rope.py:116-188): token (row %
 *   tokens_per_frame) has integer position pos_yx[tok] = (y, x); cos_t / sin_t fp32 [max_pos, 16] = cos / sin(pos * 100^(-j/16)).
 * evw_bilinear_ac_f32: heads/dpt_head.py:463-484 (F.interpolate bilinear, align_corners=True) on channels-last src fp32
 *   [F,h,w,C] -> dst [F,H,W,C] fp16 (out_fp16) or fp32, optional addend fp32 [H*W, C] added per pixel (the uv positional
 *   embedding of dpt_head.py:258-259).  C a multiple of 4.
 * evw_patchify_f16: the operand of the DINOv2 patch-embedding GEMM (models/aggregator.py:201 image normalisation +
 *   layers/patch_embed.py:66-77 Conv2d(k = stride = patch)): images fp32 [F,3,H,W] -> out fp16 [F (H/p) (W/p), Kp], columns
 *   (channel, ky, kx) of (x - mean3[c]) / std3[c] (mean3 / std3: HOST pointers), zero-padded from 3 p^2 to Kp.
 * evw_relu_inplace_f16: ResidualConvUnit's nn.ReLU(inplace=True) (heads/dpt_head.py:333,397,410): x fp32 [n] <- relu(x) in place
 *   (the skip connection then adds relu(x)) and out fp16 [n] = the same values (the convolution's operand).  n % 4 == 0.
 * evw_adaln_modulate_f32: heads/camera_head.py:118-122  out = gate * (xn * (1 + scale) + shift) + x, mod fp32 [rows, 3C] =
 *   (shift | scale | gate).
 * evw_dpt_activate_f32: heads/head_act.py:62-112 on x fp32 [rows, ld]: channels 0..n_ch-2 -> pts fp32 [rows, n_ch-1] (mode 0 =
 *   exp, 1 = inv_log), channel n_ch-1 -> conf fp32 [rows] = 1 + exp. */
int evw_qknorm_rope_f16(void* qkv, int64_t rows, int heads, int tokens_per_frame, const int* pos_yx, const float* q_gamma,
                        const float* q_beta, const float* k_gamma, const float* k_beta, const float* cos_t, const float* sin_t,
                        float eps, void* stream);
int evw_bilinear_ac_f32(const float* src, void* dst, int out_fp16, const float* addend, int F, int h, int w, int H, int W, int C,
                        void* stream);
int evw_patchify_f16(const float* images, void* out, int F, int H, int W, int patch, int Kp, const float* mean3, const float* std3,
                     void* stream);
int evw_relu_inplace_f16(float* x, void* out, int64_t n, void* stream);
int evw_adaln_modulate_f32(const float* xn, const float* mod, const float* x, float* out, int64_t rows, int C, void* stream);
int evw_dpt_activate_f32(const float* x, int64_t rows, int ld, int n_ch, int mode, float* pts, float* conf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVOWORLD_B200_H_ */
