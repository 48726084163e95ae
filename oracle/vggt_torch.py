"""TEST INFRASTRUCTURE ONLY — fp32 restatement of the VGGT forward pass the reference loop runs.

Only tests/, __graft_entry__.smoke() and the baseline legs of the benches may import this file; the product
(evoworld_b200/vggt.py) never does.

What it restates (plain torch functions over a state dict with the reference module's own keys), file:line under
/root/reference/third_party/vggt/vggt:
  * aggregator()      models/aggregator.py:184-306 (image normalisation :201, DINOv2 patch tokens :205-208, camera /
                      register tokens :213-217 + slice_expand_and_flatten :309-331, RoPE positions :219-228,
                      alternating frame / global blocks :236-256)
  * dinov2()          layers/vision_transformer.py:183-243 (interpolate_pos_encoding :183-213 — bicubic, antialias,
                      offset 0 → `size=`; prepare_tokens_with_masks :215-229; forward_features :258-275)
  * block()           layers/block.py:84-107 (eval branch), layers/attention.py:50-72 (qkv split :52-53, q/k LayerNorm
                      :54, RoPE :56-58, softmax(q k^T / sqrt(d)) v :60-68), layers/layer_scale.py:21, layers/mlp.py:33-39
  * rope2d()          layers/rope.py:84-188
  * camera_head()     heads/camera_head.py:75-140, heads/head_act.py:11-60
  * dpt_head()        heads/dpt_head.py:159-272 (_forward_impl; the frame chunking of :119-157 only bounds memory),
                      pos-embed heads/utils.py:11-109, fusion blocks heads/dpt_head.py:330-470, activate_head
                      heads/head_act.py:62-112
  * vggt_forward()    models/vggt.py:56-92

PINNED: tests/golden/vggt_golden.npz holds outputs of the REFERENCE modules themselves (imported from
/root/reference by tests/golden/make_vggt_golden.py) on seeded weights / inputs; tests/test_oracle_vggt.py checks this
restatement against them on the CPU.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

RESNET_MEAN = (0.485, 0.456, 0.406)
RESNET_STD = (0.229, 0.224, 0.225)

DEFAULT_CONFIG = dict(  # facebook/VGGT-1B (models/vggt.py:17-25 with every default)
    img_size=518, patch_size=14, embed_dim=1024, depth=24, num_heads=16, num_register_tokens=4, rope_freq=100.0,
    vit_depth=24, vit_heads=16,
    camera_heads=16, camera_trunk_depth=4, camera_iterations=4,
    dpt_features=256, dpt_out_channels=(256, 512, 1024, 1024), dpt_layers=(4, 11, 17, 23),
)


FUSED_ATTN = False  # True: F.scaled_dot_product_attention instead of the explicit softmax (long sequences)


def ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# ----------------------------------------------------------------------------- RoPE (layers/rope.py)
def rope_tables(dim: int, max_pos: int, freq: float, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """rope.py:84-114 for one spatial direction of `dim` features: [max_pos, dim] cos / sin (angles repeated twice)."""
    exponents = torch.arange(0, dim, 2, device=device).float() / dim
    inv_freq = 1.0 / (freq ** exponents)
    positions = torch.arange(max_pos, device=device, dtype=inv_freq.dtype)
    angles = torch.einsum("i,j->ij", positions, inv_freq)
    angles = torch.cat((angles, angles), dim=-1)
    return angles.cos(), angles.sin()


def rope2d(t: torch.Tensor, pos: torch.Tensor, freq: float) -> torch.Tensor:
    """t [B, heads, N, D], pos int64 [B, N, 2] (y, x) — rope.py:146-188."""
    half = t.shape[-1] // 2
    cos_t, sin_t = rope_tables(half, int(pos.max()) + 1, freq, t.device)

    def one(x, p):
        cos = F.embedding(p, cos_t)[:, None]
        sin = F.embedding(p, sin_t)[:, None]
        d = x.shape[-1]
        rot = torch.cat((-x[..., d // 2:], x[..., : d // 2]), dim=-1)
        return x * cos + rot * sin

    v, h = t.chunk(2, dim=-1)
    return torch.cat((one(v, pos[..., 0]), one(h, pos[..., 1])), dim=-1)


# ----------------------------------------------------------------------------- transformer block
def attention(x, sd, pre, heads, pos=None, qk_norm=False, rope_freq=0.0):
    B, N, C = x.shape
    hd = C // heads
    qkv = F.linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"]).reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.unbind(0)
    if qk_norm:
        q = ln(q, sd[pre + "q_norm.weight"], sd[pre + "q_norm.bias"], 1e-5)
        k = ln(k, sd[pre + "k_norm.weight"], sd[pre + "k_norm.bias"], 1e-5)
    if pos is not None and rope_freq > 0:
        q = rope2d(q, pos, rope_freq)
        k = rope2d(k, pos, rope_freq)
    if FUSED_ATTN:   # layers/attention.py:60-61 (fused_attn=True, the reference's default): the bench's eager baseline uses it
        o = F.scaled_dot_product_attention(q, k, v)
    else:
        att = (q * hd ** -0.5) @ k.transpose(-2, -1)
        o = att.softmax(dim=-1) @ v
    o = o.transpose(1, 2).reshape(B, N, C)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def mlp(x, sd, pre):
    return F.linear(F.gelu(F.linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"])), sd[pre + "fc2.weight"], sd[pre + "fc2.bias"])


def block(x, sd, pre, heads, eps, pos=None, qk_norm=False, rope_freq=0.0):
    a = attention(ln(x, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps), sd, pre + "attn.", heads, pos, qk_norm, rope_freq)
    x = x + a * sd[pre + "ls1.gamma"]
    m = mlp(ln(x, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps), sd, pre + "mlp.")
    return x + m * sd[pre + "ls2.gamma"]


# ----------------------------------------------------------------------------- DINOv2 patch tokens
def interpolate_pos_embed(pos_embed: torch.Tensor, h0: int, w0: int) -> torch.Tensor:
    """vision_transformer.py:183-213 with interpolate_offset = 0, antialias on (aggregator.py:150-151): [1, 1 + h0 w0, D]."""
    N = pos_embed.shape[1] - 1
    M = int(math.sqrt(N))
    assert N == M * M
    if h0 * w0 == N and h0 == w0:
        return pos_embed
    pe = pos_embed.float()
    dim = pe.shape[-1]
    patch = F.interpolate(pe[:, 1:].reshape(1, M, M, dim).permute(0, 3, 1, 2), mode="bicubic", antialias=True, size=(h0, w0))
    patch = patch.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((pe[:, :1], patch), dim=1)


def dinov2(x, sd, pre, cfg):
    """x [F, 3, H, W] normalised -> x_norm_patchtokens [F, h0 w0, D]."""
    p = cfg["patch_size"]
    Fr, _, H, W = x.shape
    t = F.conv2d(x, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"], stride=p).flatten(2).transpose(1, 2)
    t = torch.cat((sd[pre + "cls_token"].expand(Fr, -1, -1), t), dim=1)
    t = t + interpolate_pos_embed(sd[pre + "pos_embed"], H // p, W // p)
    t = torch.cat((t[:, :1], sd[pre + "register_tokens"].expand(Fr, -1, -1), t[:, 1:]), dim=1)
    for i in range(cfg["vit_depth"]):
        t = block(t, sd, f"{pre}blocks.{i}.", cfg["vit_heads"], 1e-6)
    t = ln(t, sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-6)
    return t[:, cfg["num_register_tokens"] + 1:]


def slice_expand_and_flatten(tok, B, S):
    q = tok[:, 0:1].expand(B, 1, *tok.shape[2:])
    o = tok[:, 1:].expand(B, S - 1, *tok.shape[2:])
    c = torch.cat([q, o], dim=1)
    return c.reshape(B * S, *c.shape[2:])


def aggregator(images, sd, cfg, pre="aggregator."):
    """images [B, S, 3, H, W] in [0, 1] -> (list of depth tensors [B, S, P, 2C], patch_start_idx)."""
    B, S, _, H, W = images.shape
    p = cfg["patch_size"]
    mean = torch.tensor(RESNET_MEAN, device=images.device).view(1, 1, 3, 1, 1)
    std = torch.tensor(RESNET_STD, device=images.device).view(1, 1, 3, 1, 1)
    x = ((images - mean) / std).view(B * S, 3, H, W)
    patches = dinov2(x, sd, pre + "patch_embed.", cfg)
    C = patches.shape[-1]
    cam = slice_expand_and_flatten(sd[pre + "camera_token"], B, S)
    reg = slice_expand_and_flatten(sd[pre + "register_token"], B, S)
    tokens = torch.cat([cam, reg, patches], dim=1)
    start = 1 + cfg["num_register_tokens"]
    h0, w0 = H // p, W // p
    grid = torch.cartesian_prod(torch.arange(h0, device=images.device), torch.arange(w0, device=images.device))
    pos = torch.cat([torch.zeros(start, 2, dtype=grid.dtype, device=grid.device), grid + 1], dim=0)[None].expand(B * S, -1, -1)
    P = tokens.shape[1]
    out = []
    for i in range(cfg["depth"]):
        tokens = block(tokens.reshape(B * S, P, C), sd, f"{pre}frame_blocks.{i}.", cfg["num_heads"], 1e-5, pos.reshape(B * S, P, 2),
                       True, cfg["rope_freq"])
        frame = tokens.reshape(B, S, P, C)
        tokens = block(tokens.reshape(B, S * P, C), sd, f"{pre}global_blocks.{i}.", cfg["num_heads"], 1e-5, pos.reshape(B, S * P, 2),
                       True, cfg["rope_freq"])
        out.append(torch.cat([frame, tokens.reshape(B, S, P, C)], dim=-1))
    return out, start


# ----------------------------------------------------------------------------- camera head
def camera_head(tokens_last, sd, cfg, pre="camera_head."):
    """tokens_last [B, S, P, 2C] -> list of activated pose encodings [B, S, 9] (camera_head.py:75-140)."""
    x = ln(tokens_last[:, :, 0], sd[pre + "token_norm.weight"], sd[pre + "token_norm.bias"], 1e-5)
    B, S, C = x.shape
    pred = None
    outs = []
    for _ in range(cfg["camera_iterations"]):
        inp = sd[pre + "empty_pose_tokens"].expand(B, S, -1) if pred is None else pred
        emb = F.linear(inp, sd[pre + "embed_pose.weight"], sd[pre + "embed_pose.bias"])
        mod = F.linear(F.silu(emb), sd[pre + "poseLN_modulation.1.weight"], sd[pre + "poseLN_modulation.1.bias"])
        shift, scale, gate = mod.chunk(3, dim=-1)
        t = gate * (F.layer_norm(x, (C,), None, None, 1e-6) * (1 + scale) + shift) + x
        for i in range(cfg["camera_trunk_depth"]):
            t = block(t, sd, f"{pre}trunk.{i}.", cfg["camera_heads"], 1e-5)
        delta = mlp(ln(t, sd[pre + "trunk_norm.weight"], sd[pre + "trunk_norm.bias"], 1e-5), sd, pre + "pose_branch.")
        pred = delta if pred is None else pred + delta
        outs.append(torch.cat([pred[..., :7], F.relu(pred[..., 7:])], dim=-1))   # trans / quat linear, FoV relu
    return outs


# ----------------------------------------------------------------------------- DPT head
def uv_pos_embed(w: int, h: int, C: int, aspect: float, device, ratio: float = 0.1) -> torch.Tensor:
    """heads/dpt_head.py:274-283 + heads/utils.py: sinusoidal embedding of the normalised uv grid -> [h, w, C] * ratio."""
    diag = (aspect ** 2 + 1.0) ** 0.5
    sx, sy = aspect / diag, 1.0 / diag
    xs = torch.linspace(-sx * (w - 1) / w, sx * (w - 1) / w, steps=w, dtype=torch.float32, device=device)
    ys = torch.linspace(-sy * (h - 1) / h, sy * (h - 1) / h, steps=h, dtype=torch.float32, device=device)
    uu, vv = torch.meshgrid(xs, ys, indexing="xy")        # [h, w]

    def sincos(D, pos):
        omega = torch.arange(D // 2, dtype=torch.double, device=device)
        omega /= D / 2.0
        omega = 1.0 / 100 ** omega
        out = torch.einsum("m,d->md", pos.reshape(-1), omega)   # float32 pos x float64 omega -> promoted to float64
        return torch.cat([torch.sin(out), torch.cos(out)], dim=1).float()

    emb = torch.cat([sincos(C // 2, uu), sincos(C // 2, vv)], dim=-1)
    return emb.view(h, w, C) * ratio


def _rcu(x, sd, pre):
    """ResidualConvUnit, heads/dpt_head.py:375-410.  Its activation is nn.ReLU(inplace=True) (_make_fusion_block :333), so
    `out = self.activation(x)` :397 overwrites x and the skip connection :410 adds relu(x), not x."""
    xr = F.relu(x)
    o = F.conv2d(xr, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    o = F.conv2d(F.relu(o), sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return o + xr


def _fusion(sd, pre, x0, x1=None, size=None):
    out = x0
    if x1 is not None:
        out = out + _rcu(x1, sd, pre + "resConfUnit1.")
    out = _rcu(out, sd, pre + "resConfUnit2.")
    if size is None:
        size = (out.shape[-2] * 2, out.shape[-1] * 2)
    out = F.interpolate(out, size=tuple(size), mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[pre + "out_conv.weight"], sd[pre + "out_conv.bias"])


def dpt_head(tokens_list, H, W, patch_start, sd, cfg, pre, activation="exp"):
    """-> (preds [B, S, H, W, C-1], conf [B, S, H, W]) — heads/dpt_head.py:159-272."""
    p = cfg["patch_size"]
    B, S = tokens_list[0].shape[:2]
    ph, pw = H // p, W // p
    aspect = W / H
    feats = []
    for j, li in enumerate(cfg["dpt_layers"]):
        x = tokens_list[li][:, :, patch_start:].reshape(B * S, ph * pw, -1)
        x = ln(x, sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-5)
        x = x.permute(0, 2, 1).reshape(B * S, -1, ph, pw)
        x = F.conv2d(x, sd[f"{pre}projects.{j}.weight"], sd[f"{pre}projects.{j}.bias"])
        x = x + uv_pos_embed(pw, ph, x.shape[1], aspect, x.device).permute(2, 0, 1)[None]
        if j == 0:
            x = F.conv_transpose2d(x, sd[pre + "resize_layers.0.weight"], sd[pre + "resize_layers.0.bias"], stride=4)
        elif j == 1:
            x = F.conv_transpose2d(x, sd[pre + "resize_layers.1.weight"], sd[pre + "resize_layers.1.bias"], stride=2)
        elif j == 3:
            x = F.conv2d(x, sd[pre + "resize_layers.3.weight"], sd[pre + "resize_layers.3.bias"], stride=2, padding=1)
        feats.append(x)
    l1, l2, l3, l4 = (F.conv2d(f, sd[f"{pre}scratch.layer{i + 1}_rn.weight"], None, padding=1) for i, f in enumerate(feats))
    out = _fusion(sd, pre + "scratch.refinenet4.", l4, None, l3.shape[2:])
    out = _fusion(sd, pre + "scratch.refinenet3.", out, l3, l2.shape[2:])
    out = _fusion(sd, pre + "scratch.refinenet2.", out, l2, l1.shape[2:])
    out = _fusion(sd, pre + "scratch.refinenet1.", out, l1, None)
    out = F.conv2d(out, sd[pre + "scratch.output_conv1.weight"], sd[pre + "scratch.output_conv1.bias"], padding=1)
    out = F.interpolate(out, size=(ph * p, pw * p), mode="bilinear", align_corners=True)
    out = out + uv_pos_embed(out.shape[-1], out.shape[-2], out.shape[1], aspect, out.device).permute(2, 0, 1)[None]
    out = F.conv2d(out, sd[pre + "scratch.output_conv2.0.weight"], sd[pre + "scratch.output_conv2.0.bias"], padding=1)
    out = F.conv2d(F.relu(out), sd[pre + "scratch.output_conv2.2.weight"], sd[pre + "scratch.output_conv2.2.bias"])
    fmap = out.permute(0, 2, 3, 1)
    xyz, conf = fmap[..., :-1], fmap[..., -1]
    if activation == "exp":
        pts = torch.exp(xyz)
    elif activation == "inv_log":
        pts = torch.sign(xyz) * torch.expm1(torch.abs(xyz))
    else:
        raise ValueError(activation)
    conf = 1 + conf.exp()
    return pts.reshape(B, S, *pts.shape[1:]), conf.reshape(B, S, *conf.shape[1:])


def vggt_forward(images, sd, cfg, point_head: bool = True) -> Dict[str, torch.Tensor]:
    """models/vggt.py:56-92 without the track head (query_points is None on the reference's path,
    unified_loop_consistency.py:135)."""
    if images.dim() == 4:
        images = images[None]
    H, W = images.shape[-2:]
    toks, start = aggregator(images, sd, cfg)
    out = {"pose_enc": camera_head(toks[-1], sd, cfg)[-1]}
    out["depth"], out["depth_conf"] = dpt_head(toks, H, W, start, sd, cfg, "depth_head.", "exp")
    if point_head and any(k.startswith("point_head.") for k in sd):
        out["world_points"], out["world_points_conf"] = dpt_head(toks, H, W, start, sd, cfg, "point_head.", "inv_log")
    out["images"] = images
    return out


# ----------------------------------------------------------------------------- the small configuration of the golden vectors
SMALL_TEST_CONFIG = dict(  # Aggregator(dinov2_vits14_reg) + CameraHead + two DPTHeads at 1/3 width; 70 x 98 images on a 5 x 5 position table
    img_size=70, patch_size=14, embed_dim=384, depth=4, num_heads=6, num_register_tokens=4, rope_freq=100.0,
    vit_depth=12, vit_heads=6, camera_heads=6, camera_trunk_depth=4, camera_iterations=4,
    dpt_features=128, dpt_out_channels=(64, 128, 256, 256), dpt_layers=(0, 1, 2, 3), point_head=True, seed=20251017,
)


def small_test_images(S: int = 3, H: int = 70, W: int = 98, seed: int = 7) -> torch.Tensor:
    """Seeded smooth-ish test clip [1, S, 3, H, W] in [0, 1] (CPU generator: identical wherever it is built)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    low = torch.rand((S, 3, H // 7, W // 7), generator=g)
    img = F.interpolate(low, size=(H, W), mode="bilinear", align_corners=False) * 0.8 + 0.2 * torch.rand((S, 3, H, W), generator=g)
    return img.clamp(0, 1)[None].contiguous()


# ----------------------------------------------------------------------------- the full-size golden vectors (VGGT-1B, 2 x 392 x 518)
FULL_TEST_SEED = 20251018


def full_test_images(S: int = 2, H: int = 392, W: int = 518, seed: int = 9) -> torch.Tensor:
    return small_test_images(S, H, W, seed)


def subsample_full(out: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The slices of a full-size result kept in tests/golden/vggt_1b_golden.npz (every 4th pixel of the maps, the special
    tokens and every 16th patch token of two aggregator layers)."""
    keep = {"pose_enc": out["pose_enc"]}
    for k in ("depth", "depth_conf", "world_points", "world_points_conf"):
        keep[k] = out[k][:, :, ::4, ::4].contiguous()
    for k in ("tokens_last", "tokens_4"):
        if k in out:
            keep[k] = torch.cat([out[k][:, :, :5], out[k][:, :, 5::16]], dim=2).contiguous()
    return keep
