"""Plücker-embedding construction (operator boundary 3).

Mirrors utils/plucker_embedding.py of the reference:
  equirectangular_to_ray  :56-116  host-side, evaluated once per process (numpy, identical formula)
  ray_c2w_to_plucker      :221-255 sm_100a kernel behind evw_plucker (csrc/reproj.cu)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def equirectangular_to_ray(target_H: int = 576, target_W: int = 1024) -> np.ndarray:
    """(H, W, 3) unit ray per equirect pixel, OpenCV RDF: image centre -> +Z, top row -> -Y.
    phi = (x/W - 1/2) 2pi, theta = (y/H - 1/2) pi, d = (cos th sin ph, sin th, cos th cos ph)."""
    ys = np.arange(target_H, dtype=np.float32)
    xs = np.arange(target_W, dtype=np.float32)
    phi = (xs / target_W - 0.5) * 2.0 * np.pi
    theta = (ys / target_H - 0.5) * np.pi
    Phi, Theta = np.meshgrid(phi, theta)
    cos_t = np.cos(Theta)
    return np.stack([cos_t * np.sin(Phi), np.sin(Theta), cos_t * np.cos(Phi)], axis=-1)


def ray_c2w_to_plucker(ray: torch.Tensor, c2w: torch.Tensor) -> torch.Tensor:
    """ray (H,W,3), c2w (N,3,4) [or (N,4,4)] -> (N,6,H,W) float32 = [R d, t x (R d)] on ray.device."""
    if ray.dim() != 3 or ray.shape[-1] != 3:
        raise ValueError(f"ray must be (H, W, 3), got {tuple(ray.shape)}")
    if c2w.dim() != 3 or c2w.shape[-1] != 4 or c2w.shape[-2] < 3:
        raise ValueError(f"c2w must be (N, 3, 4), got {tuple(c2w.shape)}")
    _lib.require_cuda(ray, "ray")
    H, W, _ = ray.shape
    N = c2w.shape[0]
    ray_c = ray.float().contiguous()
    c2w_c = c2w[:, :3, :4].to(device=ray.device, dtype=torch.float32).contiguous()
    out = torch.empty((N, 6, H, W), dtype=torch.float32, device=ray.device)
    with torch.cuda.device(ray.device):
        _lib.check(
            _lib.lib().evw_plucker(_lib.ptr(ray_c), _lib.ptr(c2w_c), _lib.ptr(out), N, H, W, _lib.stream_ptr(ray.device)),
            "evw_plucker",
        )
    return out
