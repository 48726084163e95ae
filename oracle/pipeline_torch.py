"""ORACLE — test infrastructure only (see oracle/__init__.py).

Plain-PyTorch restatement of the reference pipeline call
    StableVideoDiffusionPipeline.__call__        evoworld/pipeline/pipeline_evoworld.py:456-741
and of the single-clip caller
    prepare_batch_data / process_batch           evoworld/inference/forward_evoworld.py:119-211
with the oracle UNet (oracle/unet_torch.py) and STUB encoders standing in for the two models whose weights and
library (diffusers' AutoencoderKLTemporalDecoder, CLIP ViT-H) are not available here: `StubVAE` (one fixed strided
convolution each way) and `StubCLIP` (a fixed pooled projection).  The stubs implement exactly the attribute surface
the reference pipeline touches (`encode(x).latent_dist.mode()`, `decode(z, num_frames=n).sample`, `config.scaling_factor`,
`config.force_upcast`, `dtype`, `image_encoder(pv).image_embeds`, `parameters()`), so the product pipeline can be driven
with the same objects.

PARITY UNPINNED for the diffusers-side helpers restated from memory (randn_tensor, VideoProcessor.preprocess,
retrieve_timesteps, EulerDiscreteScheduler): they are not in the reference checkout.  Everything that *is* in the
checkout (conditioning assembly :570-643, guidance :677, loop :689-725, `_encode_image` :255-305) is followed line by
line; `_resize_with_antialiasing` is pinned by tests/golden/pipeline_golden.npz.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import unet_torch as O


# ---------------------------------------------------------------------------------------------
# stub encoders
# ---------------------------------------------------------------------------------------------


class _Dist:
    def __init__(self, mean):
        self._mean = mean

    def mode(self):
        return self._mean


class StubVAE(nn.Module):
    """8x down / up with one fixed convolution each way (seeded weights); not a model of the real VAE."""

    def __init__(self, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.enc = nn.Conv2d(3, 4, 8, stride=8)
        self.dec = nn.ConvTranspose2d(4, 3, 8, stride=8)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        self.config = SimpleNamespace(scaling_factor=0.18215, force_upcast=True, block_out_channels=(128, 256, 512, 512))

    @property
    def dtype(self):
        return self.enc.weight.dtype

    def encode(self, x):
        return SimpleNamespace(latent_dist=_Dist(self.enc(x)))

    def decode(self, z, num_frames=None):
        return SimpleNamespace(sample=self.dec(z))


class StubCLIP(nn.Module):
    """pixel_values [B,3,224,224] -> image_embeds [B, dim]: 8x8 average pool + one fixed linear map."""

    def __init__(self, dim: int = 64, seed: int = 1):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.proj = nn.Linear(3 * 8 * 8, dim)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)

    def forward(self, pixel_values):
        x = F.adaptive_avg_pool2d(pixel_values, 8).flatten(1)
        return SimpleNamespace(image_embeds=self.proj(x))


# ---------------------------------------------------------------------------------------------
# diffusers-side helpers, restated [from memory]
# ---------------------------------------------------------------------------------------------


def randn_tensor(shape, generator, device, dtype):
    """diffusers.utils.torch_utils.randn_tensor: a CPU generator draws on the CPU and the result is moved
    (navigator_evoworld.py:198 passes the global CPU generator); no generator draws on `device`."""
    if generator is not None and generator.device.type != torch.device(device).type:
        return torch.randn(shape, generator=generator, device=generator.device, dtype=dtype).to(device)
    return torch.randn(shape, generator=generator, device=device, dtype=dtype)


def video_preprocess(image01, height, width):
    """VideoProcessor.preprocess on a tensor in [0,1]: resize (nearest `interpolate`) when the size differs, then 2x-1."""
    if tuple(image01.shape[-2:]) != (height, width):
        image01 = F.interpolate(image01, size=(height, width))
    return 2.0 * image01 - 1.0


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def resize_with_antialiasing(x, size):
    """pipeline_evoworld.py:746-850 restated with plain loops over the two axes (pinned by pipeline_golden.npz)."""
    h, w = x.shape[-2:]
    out = x
    for axis, (n_in, n_out) in ((3, (w, size[1])), (2, (h, size[0]))):  # the reference filters W first, then H
        sigma = max((n_in / n_out - 1.0) / 2.0, 0.001)
        k = int(max(2.0 * 2 * sigma, 3))
        k += 1 - k % 2
        s = torch.tensor(sigma, dtype=x.dtype, device=x.device)
        t = torch.arange(k, dtype=x.dtype, device=x.device) - k // 2
        g = torch.exp(-t.pow(2.0) / (2 * s.pow(2.0)))
        g = g / g.sum(-1, keepdim=True)
        pad = (k - 1) // 2
        c = out.shape[1]
        if axis == 3:
            out = F.conv2d(F.pad(out, (pad, k - 1 - pad, 0, 0), mode="reflect"), g.view(1, 1, 1, k).repeat(c, 1, 1, 1), groups=c)
        else:
            out = F.conv2d(F.pad(out, (0, 0, pad, k - 1 - pad), mode="reflect"), g.view(1, 1, k, 1).repeat(c, 1, 1, 1), groups=c)
    return F.interpolate(out, size=tuple(size), mode="bicubic", align_corners=True)


def encode_image(clip, image01, device, do_cfg=True):
    """`_encode_image` (:255-305) for a tensor input."""
    x = image01 * 2.0 - 1.0
    x = resize_with_antialiasing(x, (224, 224))
    x = (x + 1.0) / 2.0
    mean = torch.tensor(CLIP_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    pv = ((x - mean) / std).to(device)
    emb = clip(pv).image_embeds.unsqueeze(1)
    return torch.cat([torch.zeros_like(emb), emb]) if do_cfg else emb


# ---------------------------------------------------------------------------------------------
# the pipeline call
# ---------------------------------------------------------------------------------------------


def prepare(unet, vae, clip, image, memorized_pixel_values, plucker_embedding, *, height, width, num_frames,
            num_inference_steps=25, min_guidance_scale=1.0, max_guidance_scale=3.0, fps=7, motion_bucket_id=127,
            noise_aug_strength=0.02, generator=None, latents=None, mask_mem=False, device="cpu"):
    """Steps 1-8 of `__call__` (:558-683).  Returns a dict with image_embeddings, conditional_latents, added_time_ids,
    sigmas (N+1), timesteps, latents (scaled initial noise) and guidance [1,T,1,1,1]."""
    image = torch.cat([image.unsqueeze(1), memorized_pixel_values], dim=1)                       # :570
    batch_size = image.shape[0]
    image = image / 2.0 + 0.5                                                                   # :579
    image_embeddings = encode_image(clip, image[:, 0].to(device), device)                       # :588
    fps = fps - 1                                                                               # :592
    num_cond = image.shape[1]
    flat = image.reshape(-1, *image.shape[2:])                                                  # :596
    flat = video_preprocess(flat, height, width).to(device)                                     # :597
    noise = randn_tensor(flat.shape, generator, device, flat.dtype)                             # :598
    flat = flat + noise_aug_strength * noise                                                    # :599
    lat = vae.encode(flat).latent_dist.mode()                                                   # :315 (no scaling factor)
    lat = torch.cat([torch.zeros_like(lat), lat])                                               # :320-326
    lat = lat.to(image_embeddings.dtype)
    # "(b f) c h w -> b f c h w": the CFG concat above stacks [uncond frames | cond frames] along (b f)
    image_latents = lat.reshape(2 * batch_size, num_cond, *lat.shape[1:])                       # :617
    if mask_mem:
        image_latents[:, 1:] = torch.zeros_like(image_latents[:, 1:])                           # :629-631
    plucker = torch.cat([plucker_embedding, plucker_embedding], dim=0).to(device)               # :635 (not zeroed)
    first = image_latents[:, 0:1].repeat(1, num_frames, 1, 1, 1)                                # :642
    conditional_latents = torch.cat([first, image_latents[:, 1:], plucker], dim=2)              # :643
    added = torch.tensor([[fps, motion_bucket_id, noise_aug_strength]], dtype=image_embeddings.dtype).repeat(batch_size, 1)
    added_time_ids = torch.cat([added, added]).to(device)                                       # :340-356
    sigmas = O.karras_sigmas(num_inference_steps)                                               # :658
    timesteps = O.sigma_to_timestep(sigmas[:-1])
    init_sigma = float((sigmas.max() ** 2 + 1) ** 0.5)
    shape = (batch_size, num_frames, 4, height // 8, width // 8)
    if latents is None:
        latents = randn_tensor(shape, generator, device, image_embeddings.dtype)                 # :663, :427
    latents = latents.to(device) * init_sigma                                                   # :432
    guidance = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames).unsqueeze(0).to(device, latents.dtype)
    guidance = guidance.repeat(batch_size, 1)[:, :, None, None, None]                           # :677-681
    return dict(image_embeddings=image_embeddings, conditional_latents=conditional_latents, added_time_ids=added_time_ids,
                sigmas=sigmas, timesteps=timesteps, latents=latents, guidance=guidance)


def denoise_loop(unet, st, callback=None):
    """Step 9 (:686-725) with the oracle UNet; callback(i, t, latents) -> latents | None."""
    x = st["latents"]
    sig = st["sigmas"]
    for i in range(len(sig) - 1):
        x = O.denoise_step(unet, x, st["conditional_latents"], float(sig[i]), float(sig[i + 1]), st["image_embeddings"],
                           st["added_time_ids"], st["guidance"])
        if callback is not None:
            new = callback(i, st["timesteps"][i], x)
            x = x if new is None else new
    return x


def decode(vae, latents, num_frames, decode_chunk_size):
    """decode_latents (:358-385) + VideoProcessor.postprocess_video(output_type="pt")."""
    z = latents.flatten(0, 1) / vae.config.scaling_factor
    frames = torch.cat([vae.decode(z[i:i + decode_chunk_size], num_frames=z[i:i + decode_chunk_size].shape[0]).sample
                        for i in range(0, z.shape[0], decode_chunk_size)])
    frames = frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()
    return (frames / 2 + 0.5).clamp(0, 1).permute(0, 2, 1, 3, 4)  # [B,T,C,H,W] in [0,1]


# ---------------------------------------------------------------------------------------------
# forward_evoworld.prepare_batch_data (:119-156) with the oracle's own camera / Plücker math
# ---------------------------------------------------------------------------------------------


def prepare_batch_data(batch, num_frames, height, width, rays, device):
    from . import reproj_np as R

    images = batch["pixel_values"]
    first_frame = images[:, 0].to(device)
    traj = batch["cam_traj"]
    plucker = torch.zeros(traj.shape[0], num_frames, 6, height // 8, width // 8)
    c2w_all = torch.zeros(traj.shape[0], num_frames, 3, 4)
    for i in range(traj.shape[0]):
        c2w = R.euler_to_matrix(traj[i].cpu().float(), relative=True, four_by_four=False)       # :147
        c2w_all[i] = c2w
        plucker[i] = R.ray_c2w_to_plucker(torch.as_tensor(rays).cpu().float(), c2w)           # :150
    return first_frame, c2w_all.to(device), plucker.to(device), batch["memorized_pixel_values"].to(device), images
