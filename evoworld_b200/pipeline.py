"""Denoise-loop driver (operator boundary 2).

Mirrors evoworld/pipeline/pipeline_evoworld.py:211-741 StableVideoDiffusionPipeline for the part on
the hot path: conditioning assembly (:570-643), guidance schedule (:677) and the denoise loop
(:689-725), which runs as one fused evw_denoise_step per iteration.  VAE and CLIP are SURVEY §8f
"next" rows (not built): pass precomputed `image_latents` / `image_embeddings`, or inject any objects
with diffusers' `vae.encode/decode` and `image_encoder` interfaces.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch

from .scheduler import EulerDiscreteScheduler
from .unet import UNetSpatioTemporalConditionModel


@dataclass
class StableVideoDiffusionPipelineOutput:
    frames: Union[List, np.ndarray, torch.Tensor]


def _append_dims(x, target_dims):
    return x[(...,) + (None,) * (target_dims - x.ndim)]


class StableVideoDiffusionPipeline:
    model_cpu_offload_seq = "image_encoder->unet->vae"
    _callback_tensor_inputs = ["latents"]

    def __init__(self, vae=None, image_encoder=None, unet: UNetSpatioTemporalConditionModel = None, scheduler=None,
                 feature_extractor=None):
        self.vae, self.image_encoder, self.unet, self.feature_extractor = vae, image_encoder, unet, feature_extractor
        self.scheduler = scheduler or EulerDiscreteScheduler()
        self.vae_scale_factor = 8
        self._device = unet.device if unet is not None else torch.device("cpu")
        self._progress = {}
        self._guidance_scale = None
        self._num_timesteps = 0

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, unet=None, vae=None, image_encoder=None, feature_extractor=None,
                        **kwargs):
        if unet is None:
            unet = UNetSpatioTemporalConditionModel.from_pretrained(pretrained_model_name_or_path, subfolder="unet")
        sched = EulerDiscreteScheduler.from_pretrained(pretrained_model_name_or_path, subfolder="scheduler")
        return cls(vae=vae, image_encoder=image_encoder, unet=unet, scheduler=sched, feature_extractor=feature_extractor)

    def to(self, device=None, dtype=None):
        if device is not None:
            self._device = torch.device(device)
            if self.unet is not None:
                self.unet.to(self._device)
            for m in (self.vae, self.image_encoder):
                if m is not None and hasattr(m, "to"):
                    m.to(self._device)
        return self

    @property
    def device(self):
        return self._device

    _execution_device = device

    def set_progress_bar_config(self, **kwargs):
        self._progress = kwargs

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def do_classifier_free_guidance(self):
        if isinstance(self.guidance_scale, (int, float)):
            return self.guidance_scale > 1
        return self.guidance_scale.max() > 1

    @property
    def num_timesteps(self):
        return self._num_timesteps

    def check_inputs(self, image, height, width):
        if not isinstance(image, torch.Tensor):
            raise ValueError(f"`image` has to be of type `torch.Tensor` here but is {type(image)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    def _get_add_time_ids(self, fps, motion_bucket_id, noise_aug_strength, dtype, batch_size, do_cfg):
        add_time_ids = [fps, motion_bucket_id, noise_aug_strength]
        passed = self.unet.config.addition_time_embed_dim * len(add_time_ids)
        expected = self.unet.add_embedding.linear_1.in_features
        if expected != passed:
            raise ValueError(f"Model expects an added time embedding vector of length {expected}, but a vector of "
                             f"{passed} was created. The model has an incorrect config.")
        ids = torch.tensor([add_time_ids], dtype=dtype).repeat(batch_size, 1)
        return torch.cat([ids, ids]) if do_cfg else ids

    def prepare_latents(self, batch_size, num_frames, num_channels_latents, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_frames, 4, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            gdev = generator.device if generator is not None else device
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def _encode_vae_image(self, image, device, do_cfg):
        if self.vae is None:
            raise RuntimeError("StableVideoDiffusionPipeline: no VAE attached (SURVEY §8f: VAE is not built); pass "
                               "`image_latents=[B, 1+T_mem, 4, h, w]` or inject a `vae`.")
        lat = self.vae.encode(image.to(device)).latent_dist.mode()
        return torch.cat([torch.zeros_like(lat), lat]) if do_cfg else lat

    def _encode_image(self, image, device, do_cfg):
        if self.image_encoder is None:
            raise RuntimeError("StableVideoDiffusionPipeline: no CLIP image encoder attached (SURVEY §8f: not built); pass "
                               "`image_embeddings=[B, 1, 1024]` or inject an `image_encoder`.")
        emb = self.image_encoder(image.to(device)).image_embeds.unsqueeze(1)
        return torch.cat([torch.zeros_like(emb), emb]) if do_cfg else emb

    @torch.no_grad()
    def __call__(self, image: torch.Tensor, height: int = 576, width: int = 1024, num_frames: Optional[int] = None,
                 num_inference_steps: int = 25, sigmas: Optional[List[float]] = None, min_guidance_scale: float = 1.0,
                 max_guidance_scale: float = 3.0, fps: int = 7, motion_bucket_id: int = 127, noise_aug_strength: float = 0.02,
                 decode_chunk_size: Optional[int] = None, num_videos_per_prompt: Optional[int] = 1, generator=None,
                 latents: Optional[torch.Tensor] = None, output_type: Optional[str] = "pil",
                 callback_on_step_end: Optional[Callable] = None, callback_on_step_end_tensor_inputs: List[str] = ["latents"],
                 return_dict: bool = True, plucker_embedding=None, memorized_plucker_embedding=None,
                 memorized_pixel_values=None, mask_mem: bool = False, image_embeddings: Optional[torch.Tensor] = None,
                 image_latents: Optional[torch.Tensor] = None):
        """Same arguments as the reference `__call__` (:456-483) plus two extensions that bypass the
        un-built encoders: `image_embeddings` [B,1,1024] (conditional row only) and `image_latents`
        [B, 1+T_mem, 4, h, w] (VAE `latent_dist.mode()` of first frame + memory frames, unscaled)."""
        num_frames = num_frames if num_frames is not None else self.unet.config.num_frames
        device = self._device
        self.check_inputs(image, height, width)
        batch_size = image.shape[0]
        if batch_size != 1 or num_videos_per_prompt != 1:
            raise NotImplementedError("the fused denoise step handles one clip per call (shard clips over ranks)")
        self._guidance_scale = max_guidance_scale
        do_cfg = self.do_classifier_free_guidance
        if not do_cfg:
            raise NotImplementedError("max_guidance_scale <= 1 (no classifier-free guidance) is not built")
        # 3./4. conditioning: CLIP embedding of the first frame, VAE latents of first frame + memory frames
        if image_embeddings is None:
            pix = torch.cat([image.unsqueeze(1), memorized_pixel_values], dim=1) / 2.0 + 0.5
            image_embeddings = self._encode_image(pix[:, 0], device, do_cfg)
        else:
            emb = image_embeddings.to(device, torch.float32)
            image_embeddings = torch.cat([torch.zeros_like(emb), emb]) if emb.shape[0] == batch_size else emb
        fps = fps - 1
        if image_latents is None:
            pix = torch.cat([image.unsqueeze(1), memorized_pixel_values], dim=1)
            n_cond = pix.shape[1]
            flat = pix.reshape(-1, *pix.shape[2:]).to(device)
            noise = torch.randn(flat.shape, generator=generator, device=generator.device if generator is not None else device,
                                dtype=flat.dtype).to(device)
            flat = flat + noise_aug_strength * noise
            lat = self._encode_vae_image(flat, device, do_cfg)
            lat = lat.reshape(2 * batch_size, n_cond, *lat.shape[1:])
        else:
            lat = image_latents.to(device, torch.float32)
            if lat.shape[0] == batch_size:
                lat = torch.cat([torch.zeros_like(lat), lat])
        lat = lat.clone()
        if mask_mem:
            lat[:, 1:] = 0
        _, _, _, h_lat, w_lat = lat.shape
        if plucker_embedding is None:
            raise ValueError("plucker_embedding [B, T, 6, h, w] is required")
        plucker = plucker_embedding.to(device, torch.float32)
        plucker = torch.cat([plucker, plucker], dim=0)  # NOT zeroed for the unconditional branch (:635)
        if lat.shape[1] - 1 != num_frames or plucker.shape[1] != num_frames:
            raise ValueError(f"expected {num_frames} memory frames and Plücker frames, got {lat.shape[1] - 1} and {plucker.shape[1]}")
        cond_first = lat[:, 0:1].repeat(1, num_frames, 1, 1, 1)
        conditional_latents = torch.cat([cond_first, lat[:, 1:], plucker], dim=2).contiguous()
        # 5. added time ids, 6. timesteps, 7. latents, 8. guidance
        added_time_ids = self._get_add_time_ids(fps, motion_bucket_id, noise_aug_strength, torch.float32, batch_size, do_cfg).to(device)
        self.scheduler.set_timesteps(num_inference_steps, device=device, sigmas=sigmas)
        timesteps = self.scheduler.timesteps
        latents = self.prepare_latents(batch_size, num_frames, self.unet.config.in_channels, height, width, torch.float32,
                                       device, generator, latents).contiguous()
        guidance = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames).unsqueeze(0).to(device, latents.dtype)
        self._guidance_scale = _append_dims(guidance.repeat(batch_size, 1), latents.ndim)
        # 9. denoising loop — one fused device call per iteration
        self._num_timesteps = len(timesteps)
        sig = self.scheduler.sigmas
        for i, t in enumerate(timesteps):
            self.unet.denoise_step(latents, conditional_latents, float(sig[i]), float(sig[i + 1]), image_embeddings,
                                   added_time_ids, min_guidance_scale, max_guidance_scale)
            if callback_on_step_end is not None:
                kw = {k: locals()[k] for k in callback_on_step_end_tensor_inputs}
                out = callback_on_step_end(self, i, t, kw)
                new = out.pop("latents", latents)
                if new is not latents:
                    latents.copy_(new)
        if output_type == "latent":
            frames = latents
        else:
            if self.vae is None:
                raise RuntimeError("decoding needs a VAE (SURVEY §8f: not built); use output_type='latent' or inject a `vae`.")
            frames = self.decode_latents(latents, num_frames, decode_chunk_size or num_frames)
            frames = self._postprocess(frames, output_type)
        if not return_dict:
            return frames
        return StableVideoDiffusionPipelineOutput(frames=frames)

    def decode_latents(self, latents, num_frames, decode_chunk_size=14):
        latents = latents.flatten(0, 1) / self.vae.config.scaling_factor
        frames = []
        for i in range(0, latents.shape[0], decode_chunk_size):
            n = latents[i:i + decode_chunk_size].shape[0]
            frames.append(self.vae.decode(latents[i:i + decode_chunk_size], num_frames=n).sample)
        frames = torch.cat(frames, dim=0)
        return frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()

    @staticmethod
    def _postprocess(video, output_type):
        video = (video / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 4, 1)  # [B,T,H,W,C]
        if output_type == "pt":
            return video.permute(0, 1, 4, 2, 3)
        arr = (video.cpu().numpy() * 255).round().astype("uint8")
        if output_type == "np":
            return arr
        from PIL import Image

        return [[Image.fromarray(f) for f in clip] for clip in arr]
