"""Euler discrete scheduler with the Stable-Video-Diffusion configuration (host-side logic).

Mirrors the calls the reference pipeline makes on diffusers' EulerDiscreteScheduler
(evoworld/pipeline/pipeline_evoworld.py:30,433,658,692,714): v-prediction, continuous timesteps
t = 0.25 ln(sigma), Karras sigmas (rho 7, sigma_max 700, sigma_min 0.002), final sigma 0, no churn.
The per-step arithmetic on tensors is fused into evw_denoise_step; `scale_model_input` and `step`
are kept for API compatibility (and for callers that drive the UNet themselves).
"""
from __future__ import annotations

import json
import math
import os
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

SVD_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                  prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True, sigma_min=0.002,
                  sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous", steps_offset=1,
                  rescale_betas_zero_snr=False)


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, **kwargs):
        cfg = dict(SVD_CONFIG)
        cfg.update({k: v for k, v in kwargs.items() if k in cfg})
        if cfg["prediction_type"] != "v_prediction" or not cfg["use_karras_sigmas"] or cfg["timestep_type"] != "continuous":
            raise NotImplementedError("only the SVD configuration (v_prediction, Karras sigmas, continuous t) is built")
        self.config = SimpleNamespace(**cfg)
        self.sigmas: Optional[torch.Tensor] = None
        self.timesteps: Optional[torch.Tensor] = None
        self.num_inference_steps = None
        self._step_index = None
        self.set_timesteps(25)

    @classmethod
    def from_pretrained(cls, path, subfolder: Optional[str] = None, **_):
        p = os.path.join(path, subfolder or "", "scheduler_config.json")
        if os.path.isfile(p):
            with open(p) as f:
                return cls(**json.load(f))
        return cls()

    @property
    def init_noise_sigma(self) -> float:
        return float((float(self.sigmas.max()) ** 2 + 1) ** 0.5)

    def set_timesteps(self, num_inference_steps: int, device=None, sigmas=None):
        if sigmas is None:
            ramp = np.linspace(0, 1, num_inference_steps)
            lo, hi = self.config.sigma_min ** (1 / 7.0), self.config.sigma_max ** (1 / 7.0)
            sig = (hi + ramp * (lo - hi)) ** 7.0
        else:
            sig = np.asarray(sigmas, dtype=np.float64)
            num_inference_steps = len(sig)
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(np.array([0.25 * math.log(s) for s in sig])).to(torch.float32)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]])).to(torch.float32)
        if device is not None:
            self.timesteps = self.timesteps.to(device)
        self._step_index = None

    def _index_for(self, timestep) -> int:
        if self._step_index is None:
            t = float(timestep)
            self._step_index = int(torch.argmin((self.timesteps.cpu() - t).abs()))
        return self._step_index

    def scale_model_input(self, sample: torch.Tensor, timestep) -> torch.Tensor:
        sigma = float(self.sigmas[self._index_for(timestep)])
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, **_):
        i = self._index_for(timestep)
        sigma, sigma_next = float(self.sigmas[i]), float(self.sigmas[i + 1])
        x = sample.to(torch.float32)
        x0 = model_output.to(torch.float32) * (-sigma / (sigma ** 2 + 1) ** 0.5) + x / (sigma ** 2 + 1)
        prev = x + (x - x0) / sigma * (sigma_next - sigma)
        self._step_index = i + 1
        return SimpleNamespace(prev_sample=prev.to(model_output.dtype), pred_original_sample=x0)
