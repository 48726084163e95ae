"""Depth -> world point lift (operator boundary 5).

Drop-in for third_party/vggt/vggt/utils/geometry.py:12-41 unproject_depth_map_to_point_map:
numpy/torch in, float64 numpy out.  `lift_depth_device` is the device-resident variant used by the
fused reprojection path (no host round trip).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def lift_depth_device(depth: torch.Tensor, extr: torch.Tensor, intr: torch.Tensor, out_dtype=torch.float64) -> torch.Tensor:
    """depth [S,H,W] (or [S,H,W,1]) f32, extr [S,3,4], intr [S,3,3] CUDA tensors -> [S,H,W,3] f64 | f32."""
    _lib.require_cuda(depth, "depth")
    if depth.dim() == 4:
        depth = depth[..., 0]
    S, H, W = depth.shape
    dev = depth.device
    depth = depth.float().contiguous()
    extr = extr.to(dev, torch.float32)[:, :3, :4].contiguous()
    intr = intr.to(dev, torch.float32).contiguous()
    if tuple(extr.shape) != (S, 3, 4) or tuple(intr.shape) != (S, 3, 3):
        raise ValueError("extrinsics must be [S,3,4] and intrinsics [S,3,3]")
    out = torch.empty((S, H, W, 3), dtype=out_dtype, device=dev)
    o64 = _lib.ptr(out) if out_dtype == torch.float64 else None
    o32 = _lib.ptr(out) if out_dtype == torch.float32 else None
    if o64 is None and o32 is None:
        raise ValueError("out_dtype must be float64 or float32")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().evw_lift_depth(_lib.ptr(depth), _lib.ptr(extr), _lib.ptr(intr), o64, o32, S, H, W,
                                             _lib.stream_ptr(dev)), "evw_lift_depth")
    return out


def unproject_depth_map_to_point_map(depth_map, extrinsics_cam, intrinsics_cam, device="cuda") -> np.ndarray:
    """(S,H,W,1)|(S,H,W) depth, (S,3,4) cam-from-world extrinsics, (S,3,3) intrinsics -> (S,H,W,3) float64."""
    def as_t(x):
        return (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))).to(device)

    intr_np = intrinsics_cam.detach().cpu().numpy() if isinstance(intrinsics_cam, torch.Tensor) else np.asarray(intrinsics_cam)
    assert intr_np.shape[-2:] == (3, 3), "Intrinsic matrix must be 3x3"
    assert (intr_np[..., 0, 1] == 0).all() and (intr_np[..., 1, 0] == 0).all(), "Intrinsic matrix must have zero skew"
    out = lift_depth_device(as_t(depth_map), as_t(extrinsics_cam), as_t(intrinsics_cam), torch.float64)
    return out.cpu().numpy()
