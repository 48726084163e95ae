"""ORACLE — test infrastructure only (see oracle/__init__.py).

Plain-PyTorch fp32 restatement of the VAE the pipeline calls around the denoise loop:

  AutoencoderKLTemporalDecoder.encode(x).latent_dist.mode()   pipeline_evoworld.py:307-328 (`_encode_vae_image`), :610-617
  AutoencoderKLTemporalDecoder.decode(z, num_frames).sample   pipeline_evoworld.py:358-385 (`decode_latents`), :731

The class lives in diffusers==0.31.0 (requirements.txt:36; models/autoencoders/autoencoder_kl_temporal_decoder.py with
vae.py `Encoder`, unet_2d_blocks.py `DownEncoderBlock2D` / `UNetMidBlock2D`, unet_3d_blocks.py `MidBlockTemporalDecoder` /
`UpBlockTemporalDecoder`, resnet.py, attention_processor.py) — NOT vendored in the reference and not installable here:
restated from the published implementation with the SVD checkpoint's config (block_out_channels 128/256/512/512,
layers_per_block 2, latent_channels 4, scaling_factor 0.18215).

PARITY UNPINNED: the reference holds no golden vector for the VAE and diffusers cannot be run here.  Structural checks
only (tests/test_vae_host.py): the encoder has the 34 163 592 parameters of the published SD / SVD VAE encoder (it is the
AutoencoderKL encoder), quant_conv 72, temporal decoder 63 579 183 (97 742 847 in total, "97.7 M"); state-dict key names as
in diffusers so that real checkpoints' keys line up; shape walk.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class ResnetBlock2D(nn.Module):
    """resnet.py ResnetBlock2D with temb_channels=None: GN, SiLU, conv3x3, GN, SiLU, conv3x3, + (1x1-shortcut) input."""

    def __init__(self, in_channels: int, out_channels: int, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class TemporalResnetBlock(nn.Module):
    """resnet.py TemporalResnetBlock with temb_channels=None on [B, C, T, H, W]: GroupNorm over (C/32, T, H, W)."""

    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, channels, eps=eps)
        self.conv1 = nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0))
        self.norm2 = nn.GroupNorm(32, channels, eps=eps)
        self.conv2 = nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return x + h


class AlphaBlender(nn.Module):
    """merge_strategy="learned", switch_spatial_to_temporal_mix=True (unet_3d_blocks.py MidBlockTemporalDecoder /
    UpBlockTemporalDecoder): alpha = 1 - sigmoid(mix_factor); out = alpha x_spatial + (1 - alpha) x_temporal."""

    def __init__(self, alpha: float = 0.0):
        super().__init__()
        self.mix_factor = nn.Parameter(torch.tensor([alpha]))

    def forward(self, x_spatial, x_temporal):
        a = 1.0 - torch.sigmoid(self.mix_factor).to(x_spatial.dtype)
        return a * x_spatial + (1.0 - a) * x_temporal


class SpatioTemporalResBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(in_channels, out_channels, eps=1e-6)
        self.temporal_res_block = TemporalResnetBlock(out_channels, eps=1e-5)
        self.time_mixer = AlphaBlender(0.0)

    def forward(self, x, num_frames: int):
        x = self.spatial_res_block(x)
        bf, c, h, w = x.shape
        b = bf // num_frames
        xs = x.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        xt = self.temporal_res_block(xs)
        x = self.time_mixer(xs, xt)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class Attention(nn.Module):
    """attention_processor.py Attention as the VAE mid blocks build it: one head of width C, GroupNorm(32, eps 1e-6) on the
    input, biased q/k/v/out projections, residual connection, rescale_output_factor 1."""

    def __init__(self, channels: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, channels, eps=1e-6)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])

    def forward(self, x, q_chunk: int = 4096):
        b, c, h, w = x.shape
        t = self.group_norm(x.reshape(b, c, h * w)).transpose(1, 2)  # [b, hw, c]
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        scale = c ** -0.5
        outs = []
        for i in range(0, h * w, q_chunk):  # query chunks only bound the size of the score matrix
            a = torch.softmax((q[:, i:i + q_chunk] @ k.transpose(-1, -2)) * scale, dim=-1)
            outs.append(a @ v)
        o = self.to_out[0](torch.cat(outs, dim=1))
        return x + o.transpose(1, 2).reshape(b, c, h, w)


class Downsample2D(nn.Module):
    """resnet/downsampling.py Downsample2D(use_conv=True, padding=0): zero-pad right and bottom by one, conv 3x3 stride 2."""

    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0.0))


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, num_layers: int, add_downsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels), ResnetBlock2D(channels, channels)])
        self.attentions = nn.ModuleList([Attention(channels)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    """vae.py Encoder (double_z): conv_in, 4 DownEncoderBlock2D, UNetMidBlock2D, GroupNorm + SiLU + conv_out -> 2 x latent."""

    def __init__(self, in_channels: int, latent_channels: int, block_out_channels: Sequence[int], layers_per_block: int):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        blocks, prev = [], boc[0]
        for i, ch in enumerate(boc):
            blocks.append(DownEncoderBlock2D(prev, ch, layers_per_block, add_downsample=i != len(boc) - 1))
            prev = ch
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = UNetMidBlock2D(boc[-1])
        self.conv_norm_out = nn.GroupNorm(32, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class MidBlockTemporalDecoder(nn.Module):
    def __init__(self, channels: int, num_layers: int):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(channels, channels) for _ in range(num_layers)])
        self.attentions = nn.ModuleList([Attention(channels)])

    def forward(self, x, num_frames: int):
        x = self.resnets[0](x, num_frames)
        for resnet, attn in zip(self.resnets[1:], self.attentions):
            x = attn(x)
            x = resnet(x, num_frames)
        return x


class UpBlockTemporalDecoder(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, num_layers: int, add_upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels)
                                      for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, x, num_frames: int):
        for r in self.resnets:
            x = r(x, num_frames)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class TemporalDecoder(nn.Module):
    """autoencoder_kl_temporal_decoder.py TemporalDecoder: conv_in, MidBlockTemporalDecoder, 4 UpBlockTemporalDecoder,
    GroupNorm + SiLU + conv_out, then a (3, 1, 1) convolution over the frames of the RGB output (time_conv_out)."""

    def __init__(self, latent_channels: int, out_channels: int, block_out_channels: Sequence[int], layers_per_block: int):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(latent_channels, boc[-1], 3, padding=1)
        self.mid_block = MidBlockTemporalDecoder(boc[-1], layers_per_block)
        rev = boc[::-1]
        blocks, ch = [], rev[0]
        for i, out_ch in enumerate(rev):
            blocks.append(UpBlockTemporalDecoder(ch, out_ch, layers_per_block + 1, add_upsample=i != len(rev) - 1))
            ch = out_ch
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.time_conv_out = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, z, num_frames: int):
        x = self.conv_in(z)
        x = self.mid_block(x, num_frames)
        for b in self.up_blocks:
            x = b(x, num_frames)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        bf, c, h, w = x.shape
        x = x.reshape(bf // num_frames, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        x = self.time_conv_out(x)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class AutoencoderKLTemporalDecoder(nn.Module):
    def __init__(self, in_channels: int = 3, out_channels: int = 3, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 layers_per_block: int = 2, latent_channels: int = 4, scaling_factor: float = 0.18215):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block)
        self.decoder = TemporalDecoder(latent_channels, out_channels, block_out_channels, layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.latent_channels = latent_channels
        self.scaling_factor = scaling_factor

    def encode_moments(self, x):
        """[N, 3, H, W] -> (mean, logvar) of DiagonalGaussianDistribution, each [N, latent, H/8, W/8]; .mode() == mean."""
        m = self.quant_conv(self.encoder(x))
        mean, logvar = torch.chunk(m, 2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)

    def decode(self, z, num_frames: int):
        """[N, latent, h, w] with N a multiple of num_frames -> [N, 3, 8h, 8w]."""
        return self.decoder(z, num_frames)
