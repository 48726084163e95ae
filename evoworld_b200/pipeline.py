"""Denoise-loop driver (operator boundary 2).

Mirrors evoworld/pipeline/pipeline_evoworld.py:211-741 StableVideoDiffusionPipeline for the part on
the hot path: conditioning assembly (:570-643), guidance schedule (:677) and the denoise loop
(:689-725), which runs as one fused evw_denoise_step per iteration.  The CLIP pre-processing (:255-291: x*2-1 ->
anti-aliased resize to 224x224 -> (x+1)/2 -> CLIP mean/std) is evoworld_b200/image_ops.py.  The VAE and the CLIP
network themselves are SURVEY §8f "next" rows: `from_pretrained` loads them through transformers / diffusers when those
libraries are importable, objects with the same interfaces can be injected, or pass precomputed `image_latents` /
`image_embeddings`.
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Union

import os

import numpy as np
import torch
import torch.nn.functional as F

from .image_ops import clip_preprocess
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
                        scheduler=None, strict_components: bool = True, **kwargs):
        """Reference call pattern: `from_pretrained(ckpt, unet=unet, local_files_only=True, low_cpu_mem_usage=True)`
        (forward_evoworld.py:103, navigator_evoworld.py:112, unified_loop_consistency.py:195).  Components that are not
        passed are loaded from the checkpoint's sub-folders: `unet/`, `scheduler/`, `vae/` (evoworld_b200.vae) and
        `image_encoder/` (evoworld_b200.clip) natively, `feature_extractor/` through transformers (optional: without it the
        CLIP mean / std of `image_ops` are used).  If a sub-folder exists but its library is
        not importable the call fails here with a clear message instead of at the first pipeline call
        (`strict_components=False` leaves the component None: then pass `image_latents=` / `image_embeddings=`)."""
        root = str(pretrained_model_name_or_path)
        if unet is None:
            unet = UNetSpatioTemporalConditionModel.from_pretrained(root, subfolder="unet")
        sched = scheduler or EulerDiscreteScheduler.from_pretrained(root, subfolder="scheduler")

        def load(sub, what, loader):
            if not os.path.isdir(os.path.join(root, sub)):
                return None
            try:
                return loader(os.path.join(root, sub))
            except ImportError as exc:
                if strict_components:
                    raise ImportError(f"StableVideoDiffusionPipeline.from_pretrained: `{sub}/` needs {what}, which is not "
                                      f"importable here ({exc}); install it, pass `{sub}=` explicitly, or use "
                                      f"strict_components=False and feed precomputed latents / embeddings") from exc
                return None

        if image_encoder is None and os.path.isdir(os.path.join(root, "image_encoder")):
            from .clip import CLIPVisionModelWithProjection  # native: evoworld_b200/clip.py

            image_encoder = CLIPVisionModelWithProjection.from_pretrained(root, subfolder="image_encoder")
        if feature_extractor is None:
            def _fe(path):
                from transformers import CLIPImageProcessor

                return CLIPImageProcessor.from_pretrained(path, local_files_only=True)
            feature_extractor = load("feature_extractor", "transformers", _fe)
        if vae is None and os.path.isdir(os.path.join(root, "vae")):
            from .vae import AutoencoderKLTemporalDecoder  # native: csrc/vae_host.cu

            vae = AutoencoderKLTemporalDecoder.from_pretrained(root, subfolder="vae")
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
            latents = self._randn(shape, generator, device, dtype)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def _encode_vae_image(self, image, device, num_videos_per_prompt=1, do_classifier_free_guidance=True):
        """:307-328 — `latent_dist.mode()`, NOT multiplied by the scaling factor; zeros for the unconditional half."""
        if self.vae is None:
            raise RuntimeError("StableVideoDiffusionPipeline: no VAE attached (the checkpoint has no `vae/`); pass "
                               "`image_latents=[B, 1+T_mem, 4, h, w]` or a `vae` (evoworld_b200.vae.AutoencoderKLTemporalDecoder).")
        lat = self.vae.encode(image.to(device)).latent_dist.mode()
        lat = lat.repeat(num_videos_per_prompt, 1, 1, 1)
        return torch.cat([torch.zeros_like(lat), lat]) if do_classifier_free_guidance else lat

    def _encode_image(self, image, device, num_videos_per_prompt=1, do_classifier_free_guidance=True):
        """:255-305 — `image` is a tensor in [0,1]: x*2-1 -> `_resize_with_antialiasing`(224,224) -> (x+1)/2 -> CLIP
        normalisation (`feature_extractor` when attached, else the CLIP mean/std) -> image_encoder(...).image_embeds."""
        if self.image_encoder is None:
            raise RuntimeError("StableVideoDiffusionPipeline: no CLIP image encoder attached (the checkpoint has no `image_encoder/`); pass "
                               "`image_embeddings=[B, 1, 1024]` or an `image_encoder` (evoworld_b200.clip.CLIPVisionModelWithProjection).")
        if not isinstance(image, torch.Tensor):
            arr = np.stack([np.asarray(im, dtype=np.float32) / 255.0 for im in (image if isinstance(image, list) else [image])])
            image = torch.from_numpy(arr).permute(0, 3, 1, 2)
        pv = clip_preprocess(image.to(device, torch.float32), self.feature_extractor)
        try:
            dtype = next(self.image_encoder.parameters()).dtype
        except (StopIteration, AttributeError, TypeError):
            dtype = pv.dtype
        emb = self.image_encoder(pv.to(device=device, dtype=dtype)).image_embeds.unsqueeze(1)
        bs, seq, _ = emb.shape
        emb = emb.repeat(1, num_videos_per_prompt, 1).view(bs * num_videos_per_prompt, seq, -1)
        return torch.cat([torch.zeros_like(emb), emb]) if do_classifier_free_guidance else emb

    @staticmethod
    def _video_preprocess(image01, height, width):
        """diffusers VideoProcessor.preprocess for a tensor in [0,1] (:597): resize when the size differs, then 2x-1."""
        if tuple(image01.shape[-2:]) != (height, width):
            image01 = F.interpolate(image01, size=(height, width))
        return 2.0 * image01 - 1.0

    @staticmethod
    def _randn(shape, generator, device, dtype):
        """diffusers randn_tensor: a CPU generator (navigator_evoworld.py:198) draws on the CPU, then moves."""
        if generator is not None and generator.device.type != torch.device(device).type:
            return torch.randn(shape, generator=generator, device=generator.device, dtype=dtype).to(device)
        return torch.randn(shape, generator=generator, device=device, dtype=dtype)

    def prepare_conditioning(self, image, memorized_pixel_values, plucker_embedding, height, width, num_frames, fps,
                             motion_bucket_id, noise_aug_strength, generator, mask_mem, image_embeddings=None,
                             image_latents=None):
        """Steps 2-5 of `__call__` (:570-656): returns (image_embeddings [2,1,D], conditional_latents
        [2,T,4+4+6,h,w], added_time_ids [2,3])."""
        device = self._device
        batch_size = image.shape[0]
        pix01 = None
        if image_embeddings is None or image_latents is None:
            if memorized_pixel_values is None:
                raise ValueError("memorized_pixel_values [B, T_mem, 3, H, W] is required")
            pix01 = torch.cat([image.unsqueeze(1), memorized_pixel_values.to(image.device)], dim=1) / 2.0 + 0.5  # :570,:579
        if image_embeddings is None:
            image_embeddings = self._encode_image(pix01[:, 0], device, 1, True)                                   # :588
        else:
            emb = image_embeddings.to(device, torch.float32)
            image_embeddings = torch.cat([torch.zeros_like(emb), emb]) if emb.shape[0] == batch_size else emb
        if image_latents is None:
            n_cond = pix01.shape[1]
            flat = pix01.reshape(-1, *pix01.shape[2:])                                                            # :596
            flat = self._video_preprocess(flat, height, width).to(device)                                         # :597
            noise = self._randn(flat.shape, generator, device, flat.dtype)                                        # :598
            flat = flat + noise_aug_strength * noise                                                              # :599
            lat = self._encode_vae_image(flat, device, 1, True).to(image_embeddings.dtype)                        # :610-616
            lat = lat.reshape(2 * batch_size, n_cond, *lat.shape[1:])                                             # :617
        else:
            lat = image_latents.to(device, torch.float32)
            if lat.shape[0] == batch_size:
                lat = torch.cat([torch.zeros_like(lat), lat])
            lat = lat.clone()
        if mask_mem:
            lat[:, 1:] = 0                                                                                        # :629-631
        if plucker_embedding is None:
            raise ValueError("plucker_embedding [B, T, 6, h, w] is required")
        plucker = plucker_embedding.to(device, torch.float32)
        plucker = torch.cat([plucker, plucker], dim=0)  # NOT zeroed for the unconditional branch (:635)
        if lat.shape[1] - 1 != num_frames or plucker.shape[1] != num_frames:
            raise ValueError(f"expected {num_frames} memory frames and Plücker frames, got {lat.shape[1] - 1} and {plucker.shape[1]}")
        cond_first = lat[:, 0:1].repeat(1, num_frames, 1, 1, 1)                                                   # :642
        conditional_latents = torch.cat([cond_first, lat[:, 1:], plucker], dim=2).contiguous()                    # :643
        added_time_ids = self._get_add_time_ids(fps, motion_bucket_id, noise_aug_strength, torch.float32, batch_size, True).to(device)
        return image_embeddings, conditional_latents, added_time_ids

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
        # 3.-5. conditioning: CLIP embedding of the first frame, VAE latents of first frame + memory frames, added ids
        fps = fps - 1
        image_embeddings, conditional_latents, added_time_ids = self.prepare_conditioning(
            image, memorized_pixel_values, plucker_embedding, height, width, num_frames, fps, motion_bucket_id,
            noise_aug_strength, generator, mask_mem, image_embeddings, image_latents)
        # 6. timesteps, 7. latents, 8. guidance
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
                raise RuntimeError("decoding needs a VAE (evoworld_b200.vae.AutoencoderKLTemporalDecoder); use "
                                   "output_type='latent' or pass `vae=`.")
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
