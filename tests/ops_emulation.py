"""TEST INFRASTRUCTURE ONLY: torch restatements of the evoworld_b200.ops entry points (same signatures, same rounding points:
fp16 operands, fp32 accumulation, fp16 / fp32 outputs) so that the HOST logic that strings the kernels together — weight
packing, tap orders, pixel shuffles, broadcast indices — can be checked against the golden vectors on a machine without a
GPU (tests/test_vggt_host.py monkeypatches them in).  Never imported by the product."""
import torch
import torch.nn.functional as F


def gemm_f16(a0, w, taps=((0, 0, 0, 0),), a1=None, bias=None, rowvec=None, rv_div=1, rv_mod=1, res1=None, s1=1.0, res2=None,
             s2=1.0, s0=1.0, geglu=False, out_dtype=torch.float16, block_n=0, out=None, out_lo=None, gn_stats=None,
             gn_rows_per_inst=0, act=None):
    assert act in (None, "gelu") and not (act and (rowvec is not None or res1 is not None))
    assert a0.dtype == torch.float16 and w.dtype == torch.float16 and not geglu and out_lo is None and gn_stats is None
    if a0.dim() == 2:
        a0 = a0[None, None, None]
        if a1 is not None:
            a1 = a1[None, None, None]
    B, T, Y, X, C0 = a0.shape
    assert C0 % 64 == 0 and w.shape[0] % 8 == 0
    srcs = [a0.float(), a1.float() if a1 is not None else None]
    wf = w.float()
    acc = torch.zeros((B, T, Y, X, w.shape[0]), dtype=torch.float32)
    k0 = 0
    for tap in taps:
        dx, dy, dt, src = tap
        a = srcs[src]
        C = a.shape[-1]
        pad = F.pad(a, (0, 0, 1, 1, 1, 1, 1, 1))          # zero halo of one element in x, y, t
        sh = pad[:, 1 + dt: 1 + dt + T, 1 + dy: 1 + dy + Y, 1 + dx: 1 + dx + X]
        acc += sh @ wf[:, k0: k0 + C].T
        k0 += C
    assert k0 == w.shape[1], (k0, w.shape)
    rows = B * T * Y * X
    r = acc.reshape(rows, -1)
    if bias is not None:
        r = r + bias
    r = s0 * r
    if act == "gelu":
        r = F.gelu(r)
    if rowvec is not None:
        idx = (torch.arange(rows) // rv_div) % rv_mod
        r = r + rowvec.reshape(-1, r.shape[1])[idx]
    if res1 is not None:
        r = r + s1 * res1.float().reshape(rows, -1)
    if res2 is not None:
        r = r + s2 * res2.reshape(rows, -1)
    return r.to(out_dtype)


def _sdpa(qkv, B, S, heads, hd, scale):
    q, k, v = qkv.float().view(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = ((q * scale) @ k.transpose(-2, -1)).softmax(-1)
    return (att @ v).transpose(1, 2).reshape(B * S, heads * hd).half()


def spatial_attention(qkv, frames, S, heads):
    return _sdpa(qkv, frames, S, heads, 64, 0.125)


def small_attention(qkv, B, S, heads, head_dim, scale):
    return _sdpa(qkv, B, S, heads, head_dim, scale)


def layer_norm(x, gamma, beta, eps=1e-5, rowvec=None, rv_div=1, rv_mod=1):
    assert rowvec is None and x.dtype == torch.float32
    return F.layer_norm(x, (x.shape[-1],), gamma, beta, eps).half()


def layer_norm_f32(x, gamma, beta, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)


def activation_f16(x, mode="gelu"):
    assert x.dtype == torch.float32
    f = {"gelu": F.gelu, "relu": F.relu, "identity": lambda t: t, "silu": F.silu, "quick_gelu": lambda t: t * torch.sigmoid(1.702 * t)}[mode]
    return f(x).half()


def patchify_f16(images, patch, Kp, mean, std):
    x = (images - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    cols = F.unfold(x, kernel_size=patch, stride=patch).transpose(1, 2).reshape(-1, 3 * patch * patch)
    out = torch.zeros((cols.shape[0], Kp), dtype=torch.float16)
    out[:, : cols.shape[1]] = cols
    return out


def relu_inplace_f16(x):
    x.clamp_(min=0)
    return x.half()


def qknorm_rope_(qkv, heads, tokens_per_frame, pos_yx, q_gamma, q_beta, k_gamma, k_beta, cos_t, sin_t, eps=1e-5):
    rows = qkv.shape[0]
    C = heads * 64
    tok = torch.arange(rows) % tokens_per_frame
    py, px = pos_yx[tok, 0].long(), pos_yx[tok, 1].long()
    for which, (g, b) in enumerate(((q_gamma, q_beta), (k_gamma, k_beta))):
        t = qkv[:, which * C: (which + 1) * C].float().view(rows, heads, 64)
        t = F.layer_norm(t, (64,), g, b, eps)

        def rot(x, p):                                   # x [rows, heads, 32]
            cos = torch.cat([cos_t[p], cos_t[p]], -1)[:, None]
            sin = torch.cat([sin_t[p], sin_t[p]], -1)[:, None]
            r = torch.cat([-x[..., 16:], x[..., :16]], -1)
            return x * cos + r * sin

        t = torch.cat([rot(t[..., :32], py), rot(t[..., 32:], px)], -1)
        qkv[:, which * C: (which + 1) * C] = t.reshape(rows, C).half()
    return qkv


def bilinear_ac(src, H, W, out_dtype=torch.float16, addend=None):
    o = F.interpolate(src.permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    if addend is not None:
        o = o + addend.view(1, H, W, -1)
    return o.contiguous().to(out_dtype)


def adaln_modulate(xn, mod, x):
    shift, scale, gate = mod.chunk(3, dim=-1)
    return gate * (xn * (1 + scale) + shift) + x


def dpt_activate(x, n_ch, mode):
    v = x[:, : n_ch - 1]
    pts = torch.exp(v) if mode == "exp" else torch.sign(v) * torch.expm1(v.abs())
    return pts.contiguous(), 1 + torch.exp(x[:, n_ch - 1])


ALL = ("patchify_f16", "gemm_f16", "spatial_attention", "small_attention", "layer_norm", "layer_norm_f32", "activation_f16", "relu_inplace_f16",
       "qknorm_rope_", "bilinear_ac", "adaln_modulate", "dpt_activate")
