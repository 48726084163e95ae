"""Thin Python wrappers over the tensor-core / elementwise C-ABI entry points of the UNet path.

These exist for the parity tests and micro-benchmarks; the denoise step itself is orchestrated in
C++ (csrc/unet_host.cu) behind evw_unet_forward so that one call enqueues the whole network.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

CONV3x3_TAPS = [(dx, dy, 0, 0) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]  # weight [N, ky, kx, C] tap-major
TEMPORAL_TAPS = [(0, 0, dt, 0) for dt in (-1, 0, 1)]
LINEAR_TAPS = [(0, 0, 0, 0)]


def gemm_f16(a0: torch.Tensor, w: torch.Tensor, taps: Sequence[Tuple[int, int, int, int]] = LINEAR_TAPS,
             a1: Optional[torch.Tensor] = None, bias=None, rowvec=None, rv_div: int = 1, rv_mod: int = 1, res1=None,
             s1: float = 1.0, res2=None, s2: float = 1.0, s0: float = 1.0, geglu: bool = False,
             out_dtype=torch.float16, block_n: int = 0, out: Optional[torch.Tensor] = None,
             out_lo: Optional[torch.Tensor] = None, gn_stats: Optional[torch.Tensor] = None,
             gn_rows_per_inst: int = 0, act: Optional[str] = None) -> torch.Tensor:
    """a0 fp16 [B,T,Y,X,C0] (or [M,C0] for a linear layer), w fp16 [N,K_total] -> [rows, N or N/2].
    act="gelu": exact GELU of s0 * (acc + bias) in the epilogue (no row vector / residual operands).
    gn_stats (float64 [rows / gn_rows_per_inst, 32, 2]): filled with the GroupNorm(32) sums of the output (evw_gemm_f16_gn)."""
    _lib.require_cuda(a0, "a0")
    if a0.dim() == 2:
        a0 = a0[None, None, None]
        if a1 is not None:
            a1 = a1[None, None, None]
    B, T, Y, X, C0 = a0.shape
    C1 = a1.shape[-1] if a1 is not None else 0
    N = w.shape[0]
    rows = B * T * Y * X
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((rows, n_out), dtype=out_dtype, device=a0.device)
    taps_arr = np.asarray(taps, dtype=np.int8).reshape(-1, 4)
    assert a0.dtype == torch.float16 and w.dtype == torch.float16
    with torch.cuda.device(a0.device):
        if gn_stats is not None:
            assert gn_stats.dtype == torch.float64 and gn_stats.is_contiguous() and gn_stats.numel() == rows // gn_rows_per_inst * 64
        _lib.check(_lib.lib().evw_gemm_f16_gn(
            _lib.ptr(a0), _lib.ptr(a1), _lib.ptr(w), B, T, Y, X, C0, C1, N, taps_arr.shape[0], taps_arr.tobytes(),
            _lib.ptr(out), 1 if out.dtype == torch.float16 else 0, _lib.ptr(bias), _lib.ptr(rowvec), rv_div, rv_mod,
            _lib.ptr(res1), 1 if (res1 is not None and res1.dtype == torch.float16) else 0, s1, _lib.ptr(res2), s2, s0,
            _epilogue_mode(geglu, act), block_n, _lib.ptr(out_lo), _lib.ptr(gn_stats), gn_rows_per_inst,
            _lib.stream_ptr(a0.device)), "evw_gemm_f16")
    return out


def _epilogue_mode(geglu: bool, act: Optional[str]) -> int:
    """The `geglu` argument of evw_gemm_f16: 0 = plain, 1 = GEGLU, 2 = GELU of acc + bias."""
    if act not in (None, "gelu"):
        raise ValueError(f"gemm_f16: activation {act!r} is not built into the epilogue")
    if act and geglu:
        raise ValueError("gemm_f16: GEGLU and act are exclusive")
    return 1 if geglu else (2 if act else 0)


def geglu_interleave(w: torch.Tensor, b: Optional[torch.Tensor] = None):
    """Reorder a GEGLU projection [2F, K] (= [value; gate]) into 32-row groups [16 value | 16 gate]
    so that value and gate of the same output column land in the same accumulator tile."""
    F2, K = w.shape
    F = F2 // 2
    assert F % 16 == 0
    v, g = w[:F].reshape(F // 16, 16, K), w[F:].reshape(F // 16, 16, K)
    wi = torch.stack([v, g], dim=1).reshape(F2, K).contiguous()
    bi = None
    if b is not None:
        bi = torch.stack([b[:F].reshape(F // 16, 16), b[F:].reshape(F // 16, 16)], dim=1).reshape(F2).contiguous()
    return wi, bi


def spatial_attention(qkv: torch.Tensor, frames: int, S: int, heads: int) -> torch.Tensor:
    """qkv fp16 [frames*S, 3*heads*64] -> fp16 [frames*S, heads*64]."""
    _lib.require_cuda(qkv, "qkv")
    out = torch.empty((frames * S, heads * 64), dtype=torch.float16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.lib().evw_spatial_attention_f16(_lib.ptr(qkv), _lib.ptr(out), frames, S, heads,
                                                        _lib.stream_ptr(qkv.device)), "evw_spatial_attention_f16")
    return out


def temporal_attention(qkv: torch.Tensor, B: int, T: int, S: int, heads: int) -> torch.Tensor:
    """qkv fp16 [B*T*S, 3*heads*64] rows (b,t,s) -> fp16 [B*T*S, heads*64], softmax over t."""
    _lib.require_cuda(qkv, "qkv")
    out = torch.empty((B * T * S, heads * 64), dtype=torch.float16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.lib().evw_temporal_attention_f16(_lib.ptr(qkv), _lib.ptr(out), B, T, S, heads,
                                                         _lib.stream_ptr(qkv.device)), "evw_temporal_attention_f16")
    return out


def group_norm(src0: torch.Tensor, gamma, beta, insts: int, eps: float, silu: bool, src1: Optional[torch.Tensor] = None,
               want_raw: bool = False, want_lo: bool = False):
    """src0 [rows, C0] fp32|fp16 (+ src1 [rows, C1] fp32) -> fp16 [rows, C0+C1] (and the raw fp16 copy)."""
    _lib.require_cuda(src0, "src0")
    rows, C0 = src0.shape
    C1 = src1.shape[1] if src1 is not None else 0
    out = torch.empty((rows, C0 + C1), dtype=torch.float16, device=src0.device)
    raw = torch.empty_like(out) if want_raw else None
    lo = torch.empty_like(out) if want_lo else None
    ws = torch.empty(insts * 64, dtype=torch.float64, device=src0.device)
    with torch.cuda.device(src0.device):
        _lib.check(_lib.lib().evw_group_norm_f16(_lib.ptr(src0), 1 if src0.dtype == torch.float16 else 0, C0, _lib.ptr(src1),
                                                 C1, insts, rows // insts, eps, _lib.ptr(gamma), _lib.ptr(beta),
                                                 1 if silu else 0, _lib.ptr(ws), _lib.ptr(out), _lib.ptr(raw), _lib.ptr(lo),
                                                 _lib.stream_ptr(src0.device)), "evw_group_norm_f16")
    res = (out,) + ((raw,) if want_raw else ()) + ((lo,) if want_lo else ())
    return res if len(res) > 1 else out


def layer_norm(x: torch.Tensor, gamma, beta, eps: float = 1e-5, rowvec=None, rv_div: int = 1, rv_mod: int = 1):
    _lib.require_cuda(x, "x")
    rows, C = x.shape
    out = torch.empty((rows, C), dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().evw_layer_norm_f16(_lib.ptr(x), _lib.ptr(rowvec), rv_div, rv_mod, rows, C, eps, _lib.ptr(gamma),
                                                 _lib.ptr(beta), _lib.ptr(out), _lib.stream_ptr(x.device)), "evw_layer_norm_f16")
    return out


def small_attention(qkv: torch.Tensor, B: int, S: int, heads: int, head_dim: int, scale: float) -> torch.Tensor:
    """qkv fp16 [B*S, 3*heads*head_dim] (q | k | v) -> fp16 [B*S, heads*head_dim]; S <= 1024, any head width <= 256."""
    _lib.require_cuda(qkv, "qkv")
    assert qkv.dtype == torch.float16 and qkv.shape == (B * S, 3 * heads * head_dim)
    out = torch.empty((B * S, heads * head_dim), dtype=torch.float16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.lib().evw_small_attention_f16(_lib.ptr(qkv), _lib.ptr(out), B, S, heads, head_dim, scale,
                                                      _lib.stream_ptr(qkv.device)), "evw_small_attention_f16")
    return out


def activation_f16(x: torch.Tensor, mode: str = "gelu") -> torch.Tensor:
    """x fp32 -> fp16 through GELU (erf form), quick_gelu, relu, identity (cast) or silu."""
    _lib.require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.is_contiguous()
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().evw_act_f16(_lib.ptr(x), _lib.ptr(out), x.numel(), {"gelu": 0, "quick_gelu": 1, "relu": 2, "identity": 3, "silu": 4}[mode],
                                          _lib.stream_ptr(x.device)), "evw_act_f16")
    return out


def layer_norm_f32(x: torch.Tensor, gamma, beta, eps: float = 1e-5) -> torch.Tensor:
    """LayerNorm over the last dimension of x fp32 [rows, C] with an fp32 result."""
    _lib.require_cuda(x, "x")
    rows, C = x.shape
    out = torch.empty((rows, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().evw_layer_norm_f32(_lib.ptr(x), rows, C, eps, _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(out),
                                                 _lib.stream_ptr(x.device)), "evw_layer_norm_f32")
    return out


def upconv_weights(w: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [N, C, 3, 3] of an Upsample2D (nearest x2 then 3x3 conv, padding 1) -> the four phase matrices fp16
    [4, N, 4C] of the fused form: phase = 2 py + px, taps (dy, dx) in the order of csrc/tc_gemm.cu::upconv2x_phase, the 3x3
    weights that fall on the same low-resolution pixel summed in fp32 (rows: py = 0 -> {k0}, {k1 + k2}; py = 1 -> {k0 + k1}, {k2})."""
    w = w.to(torch.float32)
    groups = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    phases = []
    for py in (0, 1):
        for px in (0, 1):
            taps = []
            for ky in groups[py]:
                for kx in groups[px]:
                    taps.append(w[:, :, ky][:, :, :, kx].sum(dim=(2, 3)))  # [N, C]
            phases.append(torch.cat(taps, dim=1))                            # [N, 4C]
    return torch.stack(phases).to(torch.float16).contiguous()


def upconv2x(a: torch.Tensor, w4: torch.Tensor, bias=None) -> torch.Tensor:
    """a fp16 [F, h, w, C] -> fp32 [F, 2h, 2w, N]: nearest x2 + 3x3 conv as four 2x2 phase convolutions (evw_upconv2x_f16)."""
    _lib.require_cuda(a, "a")
    F_, h, w, C = a.shape
    N = w4.shape[1]
    assert a.dtype == torch.float16 and w4.dtype == torch.float16 and w4.shape == (4, N, 4 * C)
    out = torch.empty((F_, 2 * h, 2 * w, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().evw_upconv2x_f16(_lib.ptr(a), _lib.ptr(w4), _lib.ptr(bias), _lib.ptr(out), F_, h, w, C, N,
                                               _lib.stream_ptr(a.device)), "evw_upconv2x_f16")
    return out


def qknorm_rope_(qkv: torch.Tensor, heads: int, tokens_per_frame: int, pos_yx: torch.Tensor, q_gamma, q_beta, k_gamma, k_beta,
                 cos_t: torch.Tensor, sin_t: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """In place on qkv fp16 [rows, 3*heads*64]: LayerNorm(64) of every q / k head, then the 2-D rotary embedding with the
    integer (y, x) position pos_yx int32 [tokens_per_frame, 2] of token row % tokens_per_frame (evw_qknorm_rope_f16)."""
    _lib.require_cuda(qkv, "qkv")
    rows = qkv.shape[0]
    assert qkv.dtype == torch.float16 and qkv.shape[1] == 3 * heads * 64 and rows % tokens_per_frame == 0
    assert pos_yx.dtype == torch.int32 and pos_yx.shape == (tokens_per_frame, 2)
    assert cos_t.dtype == torch.float32 and cos_t.shape[1] == 16 and sin_t.shape == cos_t.shape
    assert int(pos_yx.max()) < cos_t.shape[0]
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.lib().evw_qknorm_rope_f16(_lib.ptr(qkv), rows, heads, tokens_per_frame, _lib.ptr(pos_yx), _lib.ptr(q_gamma),
                                                  _lib.ptr(q_beta), _lib.ptr(k_gamma), _lib.ptr(k_beta), _lib.ptr(cos_t),
                                                  _lib.ptr(sin_t), eps, _lib.stream_ptr(qkv.device)), "evw_qknorm_rope_f16")
    return qkv


def bilinear_ac(src: torch.Tensor, H: int, W: int, out_dtype=torch.float16, addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src fp32 [F, h, w, C] -> [F, H, W, C]: F.interpolate(mode="bilinear", align_corners=True), channels last, plus an
    optional per-pixel addend fp32 [H*W, C] (evw_bilinear_ac_f32)."""
    _lib.require_cuda(src, "src")
    F_, h, w, C = src.shape
    assert src.dtype == torch.float32 and (addend is None or (addend.dtype == torch.float32 and addend.numel() == H * W * C))
    out = torch.empty((F_, H, W, C), dtype=out_dtype, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.lib().evw_bilinear_ac_f32(_lib.ptr(src), _lib.ptr(out), 1 if out_dtype == torch.float16 else 0,
                                                  _lib.ptr(addend), F_, h, w, H, W, C, _lib.stream_ptr(src.device)),
                   "evw_bilinear_ac_f32")
    return out


def patchify_f16(images: torch.Tensor, patch: int, Kp: int, mean: Sequence[float], std: Sequence[float]) -> torch.Tensor:
    """images fp32 [F, 3, H, W] -> fp16 [F (H/p) (W/p), Kp]: rows = patches, columns = (channel, ky, kx) of (x - mean[c]) / std[c],
    zero-padded to Kp — the operand of the patch-embedding GEMM (evw_patchify_f16)."""
    _lib.require_cuda(images, "images")
    F_, Cc, H, W = images.shape
    assert Cc == 3 and images.dtype == torch.float32 and images.is_contiguous()
    out = torch.empty((F_ * (H // patch) * (W // patch), Kp), dtype=torch.float16, device=images.device)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(images.device):
        _lib.check(_lib.lib().evw_patchify_f16(_lib.ptr(images), _lib.ptr(out), F_, H, W, patch, Kp, m, s, _lib.stream_ptr(images.device)),
                   "evw_patchify_f16")
    return out


def relu_inplace_f16(x: torch.Tensor) -> torch.Tensor:
    """x fp32 <- relu(x) in place; returns the fp16 copy (evw_relu_inplace_f16: nn.ReLU(inplace=True) feeding a convolution)."""
    _lib.require_cuda(x, "x")
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() % 4 == 0
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().evw_relu_inplace_f16(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr(x.device)), "evw_relu_inplace_f16")
    return out


def adaln_modulate(xn: torch.Tensor, mod: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """gate * (xn * (1 + scale) + shift) + x with mod fp32 [rows, 3C] = (shift | scale | gate) (evw_adaln_modulate_f32)."""
    _lib.require_cuda(x, "x")
    rows, C = x.shape
    assert xn.shape == x.shape and mod.shape == (rows, 3 * C) and all(t.dtype == torch.float32 for t in (xn, mod, x))
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().evw_adaln_modulate_f32(_lib.ptr(xn), _lib.ptr(mod), _lib.ptr(x), _lib.ptr(out), rows, C,
                                                     _lib.stream_ptr(x.device)), "evw_adaln_modulate_f32")
    return out


def dpt_activate(x: torch.Tensor, n_ch: int, mode: str):
    """x fp32 [rows, ld] -> (pts fp32 [rows, n_ch-1] through exp / inv_log, conf fp32 [rows] = 1 + exp) (evw_dpt_activate_f32)."""
    _lib.require_cuda(x, "x")
    rows, ld = x.shape
    assert x.dtype == torch.float32
    pts = torch.empty((rows, n_ch - 1), dtype=torch.float32, device=x.device)
    conf = torch.empty((rows,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().evw_dpt_activate_f32(_lib.ptr(x), rows, ld, n_ch, {"exp": 0, "inv_log": 1}[mode], _lib.ptr(pts),
                                                   _lib.ptr(conf), _lib.stream_ptr(x.device)), "evw_dpt_activate_f32")
    return pts, conf
