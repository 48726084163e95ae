"""ORACLE — test infrastructure only (see oracle/__init__.py).

Plain-PyTorch fp32 restatement of the denoise step's arithmetic:

  UNetSpatioTemporalConditionModel   evoworld/trainer/unet_plucker.py:30-488 (wrapper, in the reference)
  its blocks                         diffusers==0.31.0 (requirements.txt:36) models/unets/unet_3d_blocks.py,
                                     models/resnet.py, models/attention.py, models/transformers/
                                     transformer_temporal.py, models/embeddings.py — NOT vendored in
                                     the reference and not installable here: restated from the
                                     published implementation.
  EulerDiscreteScheduler             diffusers==0.31.0 schedulers/scheduling_euler_discrete.py (SVD config)
  denoise-loop body                  evoworld/pipeline/pipeline_evoworld.py:689-725

PARITY UNPINNED: the reference holds no golden vector for the UNet or the scheduler and diffusers
cannot be run here.  Structural checks only: state-dict key set of SURVEY A.5, analytic parameter
count 1 524 623 082 + 28 800 (the ten extra conv_in channels), shape walk.  Module/attribute names
follow diffusers so that real checkpoints' keys line up.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# embeddings (diffusers/models/embeddings.py)
# ---------------------------------------------------------------------------------------------


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos(t f_k), sin(t f_k)], f_k = 10000^(-k/(dim/2))."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, out_dim: Optional[int] = None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class AlphaBlender(nn.Module):
    """merge_strategy="learned_with_images" with image_only_indicator == 0 everywhere
    (unet_plucker.py:428): alpha = sigmoid(mix_factor); out = alpha x_spatial + (1 - alpha) x_temporal."""

    def __init__(self, alpha: float = 0.5):
        super().__init__()
        self.mix_factor = nn.Parameter(torch.tensor([alpha]))

    def forward(self, x_spatial, x_temporal):
        a = torch.sigmoid(self.mix_factor).to(x_spatial.dtype)
        return a * x_spatial + (1.0 - a) * x_temporal


# ---------------------------------------------------------------------------------------------
# resnets (diffusers/models/resnet.py)
# ---------------------------------------------------------------------------------------------


class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class TemporalResnetBlock(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps)
        self.conv1 = nn.Conv3d(in_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps)
        self.conv2 = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, x, temb):  # x [B,C,T,H,W], temb [B,T,E]
        h = self.conv1(F.silu(self.norm1(x)))
        t = self.time_emb_proj(F.silu(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
        h = h + t
        h = self.conv2(F.silu(self.norm2(h)))
        return x + h


class SpatioTemporalResBlock(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, eps):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(in_channels, out_channels, temb_channels, eps)
        self.temporal_res_block = TemporalResnetBlock(out_channels, out_channels, temb_channels, eps)
        self.time_mixer = AlphaBlender(0.5)

    def forward(self, x, temb, num_frames):
        x = self.spatial_res_block(x, temb)
        bf, c, h, w = x.shape
        b = bf // num_frames
        xs = x.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        xt = self.temporal_res_block(xs, temb.reshape(b, num_frames, -1))
        x = self.time_mixer(xs, xt)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


# ---------------------------------------------------------------------------------------------
# attention (diffusers/models/attention.py, attention_processor.py)
# ---------------------------------------------------------------------------------------------


class Attention(nn.Module):
    def __init__(self, query_dim, heads, dim_head, cross_attention_dim=None):
        super().__init__()
        inner = heads * dim_head
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv, inner, bias=False)
        self.to_v = nn.Linear(kv, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        b, n, _ = x.shape
        q = self.to_q(x).view(b, n, self.heads, -1).transpose(1, 2)
        k = self.to_k(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(b, ctx.shape[1], self.heads, -1).transpose(1, 2)
        scale = q.shape[-1] ** -0.5
        attn = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(b, n, -1)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim_out or dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        return self.ff(self.norm3(x)) + x


class TemporalBasicTransformerBlock(nn.Module):
    def __init__(self, dim, time_mix_inner_dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.is_res = dim == time_mix_inner_dim
        self.norm_in = nn.LayerNorm(dim)
        self.ff_in = FeedForward(dim, dim_out=time_mix_inner_dim)
        self.norm1 = nn.LayerNorm(time_mix_inner_dim)
        self.attn1 = Attention(time_mix_inner_dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(time_mix_inner_dim)
        self.attn2 = Attention(time_mix_inner_dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(time_mix_inner_dim)
        self.ff = FeedForward(time_mix_inner_dim)

    def forward(self, x, num_frames, context):  # x [(b t), s, c]
        bf, s, c = x.shape
        b = bf // num_frames
        x = x.reshape(b, num_frames, s, c).permute(0, 2, 1, 3).reshape(b * s, num_frames, c)
        res = x
        x = self.ff_in(self.norm_in(x))
        if self.is_res:
            x = x + res
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        ff = self.ff(self.norm3(x))
        x = ff + x if self.is_res else ff
        return x.reshape(b, s, num_frames, c).permute(0, 2, 1, 3).reshape(bf, s, c)


class TransformerSpatioTemporalModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, cross_attention_dim, num_layers=1):
        super().__init__()
        inner = heads * dim_head
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.temporal_transformer_blocks = nn.ModuleList(
            [TemporalBasicTransformerBlock(inner, inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.time_pos_embed = TimestepEmbedding(in_channels, in_channels * 4, out_dim=in_channels)
        self.time_mixer = AlphaBlender(0.5)
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, x, context, num_frames):  # x [(b t), c, h, w], context [(b t), 1, d]
        bf, c, h, w = x.shape
        b = bf // num_frames
        tc = context.reshape(b, num_frames, -1, context.shape[-1])[:, 0]
        tc = tc[:, None].broadcast_to(b, h * w, tc.shape[-2], tc.shape[-1]).reshape(b * h * w, -1, tc.shape[-1])
        residual = x
        x = self.norm(x).permute(0, 2, 3, 1).reshape(bf, h * w, c)
        x = self.proj_in(x)
        frames = torch.arange(num_frames, device=x.device).repeat(b)
        emb = self.time_pos_embed(timestep_embedding(frames, self.in_channels).to(x.dtype))[:, None, :]
        for blk, tblk in zip(self.transformer_blocks, self.temporal_transformer_blocks):
            x = blk(x, context)
            x_mix = tblk(x + emb, num_frames, tc)
            x = self.time_mixer(x, x_mix)
        x = self.proj_out(x)
        return x.reshape(bf, h, w, c).permute(0, 3, 1, 2) + residual


# ---------------------------------------------------------------------------------------------
# blocks (diffusers/models/unets/unet_3d_blocks.py)
# ---------------------------------------------------------------------------------------------


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """CrossAttnDownBlockSpatioTemporal (has_attn, eps 1e-6) / DownBlockSpatioTemporal (eps 1e-5)."""

    def __init__(self, cin, cout, temb, layers, heads, cross_dim, has_attn, add_down, eps):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(cin if i == 0 else cout, cout, temb, eps) for i in range(layers)])
        if has_attn:
            self.attentions = nn.ModuleList([TransformerSpatioTemporalModel(heads, cout // heads, cout, cross_dim) for _ in range(layers)])
        self.has_attn = has_attn
        if add_down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        self.add_down = add_down

    def forward(self, x, temb, context, T):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb, T)
            if self.has_attn:
                x = self.attentions[i](x, context, T)
            outs.append(x)
        if self.add_down:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb, heads, cross_dim, eps=1e-5):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(ch, ch, temb, eps), SpatioTemporalResBlock(ch, ch, temb, eps)])
        self.attentions = nn.ModuleList([TransformerSpatioTemporalModel(heads, ch // heads, ch, cross_dim)])

    def forward(self, x, temb, context, T):
        x = self.resnets[0](x, temb, T)
        x = self.attentions[0](x, context, T)
        return self.resnets[1](x, temb, T)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, prev, temb, layers, heads, cross_dim, has_attn, add_up, eps=1e-6):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            rin = prev if i == 0 else cout
            rs.append(SpatioTemporalResBlock(rin + skip, cout, temb, eps))
        self.resnets = nn.ModuleList(rs)
        if has_attn:
            self.attentions = nn.ModuleList([TransformerSpatioTemporalModel(heads, cout // heads, cout, cross_dim) for _ in range(layers)])
        self.has_attn = has_attn
        if add_up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.add_up = add_up

    def forward(self, x, skips, temb, context, T):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb, T)
            if self.has_attn:
                x = self.attentions[i](x, context, T)
        if self.add_up:
            x = self.upsamplers[0](x)
        return x


class UNetSpatioTemporalConditionModel(nn.Module):
    """unet_plucker.py:30-488 with diffusers' blocks restated above."""

    def __init__(self, in_channels=18, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
                 cross_attention_dim=1024, num_attention_heads=(5, 10, 20, 20), addition_time_embed_dim=256,
                 projection_class_embeddings_input_dim=768, num_frames=25,
                 down_attn=(True, True, True, False), eps_cross=1e-6, eps_plain_down=1e-5, eps_mid=1e-5, eps_up=1e-6):
        super().__init__()
        boc = tuple(block_out_channels)
        self.cfg = dict(in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
                        layers_per_block=layers_per_block, cross_attention_dim=cross_attention_dim,
                        num_attention_heads=tuple(num_attention_heads), addition_time_embed_dim=addition_time_embed_dim,
                        projection_class_embeddings_input_dim=projection_class_embeddings_input_dim, num_frames=num_frames)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        temb = boc[0] * 4
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, temb)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i in range(len(boc)):
            in_ch, out_ch = out_ch, boc[i]
            last = i == len(boc) - 1
            self.down_blocks.append(DownBlock(in_ch, out_ch, temb, layers_per_block, num_attention_heads[i], cross_attention_dim,
                                              down_attn[i], not last, eps_cross if down_attn[i] else eps_plain_down))
        self.mid_block = MidBlock(boc[-1], temb, num_attention_heads[-1], cross_attention_dim, eps_mid)
        self.up_blocks = nn.ModuleList()
        rev, rev_heads, rev_attn = boc[::-1], tuple(num_attention_heads)[::-1], tuple(down_attn)[::-1]
        out_ch = rev[0]
        for i in range(len(boc)):
            last = i == len(boc) - 1
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(in_ch, out_ch, prev, temb, layers_per_block + 1, rev_heads[i], cross_attention_dim,
                                          rev_attn[i], not last, eps_up))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids):
        """sample [B,T,C,h,w]; timestep 0-d/1-elt tensor or float; ehs [B,1,D]; added_time_ids [B,3] -> [B,T,4,h,w]."""
        B, T = sample.shape[:2]
        t = torch.as_tensor(timestep, dtype=torch.float32, device=sample.device).reshape(-1).expand(B)
        boc0 = self.cfg["block_out_channels"][0]
        emb = self.time_embedding(timestep_embedding(t, boc0).to(sample.dtype))
        te = timestep_embedding(added_time_ids.flatten(), self.cfg["addition_time_embed_dim"]).reshape(B, -1).to(emb.dtype)
        emb = emb + self.add_embedding(te)
        x = sample.flatten(0, 1)
        emb = emb.repeat_interleave(T, dim=0)
        ctx = encoder_hidden_states.repeat_interleave(T, dim=0)
        x = self.conv_in(x)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, ctx, T)
            skips += outs
        x = self.mid_block(x, emb, ctx, T)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, ctx, T)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return x.reshape(B, T, *x.shape[1:])


# ---------------------------------------------------------------------------------------------
# scheduler + loop body
# ---------------------------------------------------------------------------------------------


def karras_sigmas(num_steps: int, sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0) -> torch.Tensor:
    """EulerDiscreteScheduler.set_timesteps with use_karras_sigmas (SVD config): N sigmas + final 0 (float32)."""
    ramp = torch.linspace(0, 1, num_steps, dtype=torch.float64)
    min_inv, max_inv = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (max_inv + ramp * (min_inv - max_inv)) ** rho
    return torch.cat([sig, torch.zeros(1, dtype=torch.float64)]).float()


def sigma_to_timestep(sigmas: torch.Tensor) -> torch.Tensor:
    """timestep_type == "continuous": t = 0.25 ln(sigma)."""
    return 0.25 * torch.log(sigmas)


def denoise_step(unet, latents, cond_latents, sigma, sigma_next, ehs, added_time_ids, guidance):
    """One iteration of pipeline_evoworld.py:689-725 (v-prediction Euler step, no churn).
    latents [1,T,4,h,w]; cond_latents [2,T,14,h,w]; guidance [1,T,1,1,1]."""
    x_in = torch.cat([latents] * 2) / ((sigma ** 2 + 1) ** 0.5)
    x_in = torch.cat([x_in, cond_latents], dim=2)
    v = unet(x_in, 0.25 * math.log(sigma), ehs, added_time_ids)
    vu, vc = v.chunk(2)
    v = vu + guidance * (vc - vu)
    x0 = v * (-sigma / (sigma ** 2 + 1) ** 0.5) + latents / (sigma ** 2 + 1)
    d = (latents - x0) / sigma
    return latents + d * (sigma_next - sigma)
