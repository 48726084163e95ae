"""Device-resident evolving 3D point memory (the caller loop around hot path 2).

The reference's `process_episode` (unified_loop_consistency.py:398-492) rebuilds everything from scratch each segment:
tensor -> PIL -> PNG -> tmpdir -> reload for VGGT (:339-348), numpy lift of ALL frames so far (:366), float64 points and
float32 images handed to `predictions_to_target_view` on the host (:470-485).  `PointMemory` keeps the packed cloud
(`float4 {x, y, z, rgb}` per pixel, 16 B) and its confidences in HBM across segments: a segment appends only its new
frames (fused lift + pack, evw_lift_pack_points), and the joint confidence percentile over everything accumulated so far
(reproject_vggt_open3d_utils.py:294-310) + order-preserving compaction (evw_conf_select) produce the `PointScene` the
splat consumes.  Given the same per-frame predictions the scene is bit-identical to the one-shot
`PointCloudProcessor.filter_predictions_device` over all frames (tests/test_gpu_reproj.py::test_point_memory_*).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .reprojection import PointScene, conf_select_device


class PointMemory:
    def __init__(self, height: int = 392, width: int = 518, capacity_frames: int = 74, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PointMemory lives on a CUDA device (no CPU fallback)")
        self.H, self.W, self.capacity_frames = height, width, capacity_frames
        n = capacity_frames * height * width
        self.pts4 = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        self.conf = torch.empty(n, dtype=torch.float32, device=self.device)
        self.frames = 0

    def __len__(self) -> int:
        return self.frames * self.H * self.W

    def reset(self):
        self.frames = 0

    def append(self, depth: torch.Tensor, conf: torch.Tensor, images: torch.Tensor, extrinsic: torch.Tensor,
               intrinsic: torch.Tensor) -> "PointMemory":
        """New frames: depth [S,H,W] (or [S,H,W,1]), conf [S,H,W], images [S,3,H,W] in [0,1], extrinsic [S,3,4]
        (cam-from-world), intrinsic [S,3,3] — CUDA tensors (host tensors are copied over, non-blocking)."""
        dev = self.device
        d = depth.to(dev, torch.float32, non_blocking=True)
        if d.dim() == 4:
            d = d[..., 0]
        S = d.shape[0]
        if tuple(d.shape[1:]) != (self.H, self.W):
            raise ValueError(f"depth must be [S,{self.H},{self.W}], got {tuple(depth.shape)}")
        if self.frames + S > self.capacity_frames:
            raise ValueError(f"PointMemory: {self.frames} + {S} frames exceed the capacity of {self.capacity_frames}")
        d = d.contiguous()
        img = images.to(dev, torch.float32, non_blocking=True).contiguous()
        ex = extrinsic.to(dev, torch.float32, non_blocking=True)[:, :3, :4].contiguous()
        k = intrinsic.to(dev, torch.float32, non_blocking=True).contiguous()
        if tuple(img.shape) != (S, 3, self.H, self.W) or tuple(ex.shape) != (S, 3, 4) or tuple(k.shape) != (S, 3, 3):
            raise ValueError("images must be [S,3,H,W], extrinsic [S,3,4], intrinsic [S,3,3]")
        n0, n = len(self), S * self.H * self.W
        out = self.pts4[n0:n0 + n]
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().evw_lift_pack_points(_lib.ptr(d), _lib.ptr(ex), _lib.ptr(k), _lib.ptr(img), _lib.ptr(out),
                                                       S, self.H, self.W, _lib.stream_ptr(dev)), "evw_lift_pack_points")
        self.conf[n0:n0 + n].copy_(conf.to(dev, torch.float32, non_blocking=True).reshape(-1), non_blocking=True)
        self.frames += S
        return self

    def scene(self, conf_thres: float = 50.0, first_frame: int = 0) -> PointScene:
        """Joint percentile filter over frames [first_frame, frames) -> compacted PointScene (no host sync)."""
        hw = self.H * self.W
        a, b = first_frame * hw, len(self)
        if b <= a:
            raise ValueError("PointMemory.scene: no frames")
        out, _, count, _ = conf_select_device(self.conf[a:b], self.pts4[a:b], conf_thres)
        return PointScene(out, count)
