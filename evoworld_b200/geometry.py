"""Pose conversion helpers (host side; negligible cost).

  xyz_euler_to_four_by_four_matrix_batch    utils/geometry.py:5-88
  xyz_euler_to_three_by_four_matrix_batch   dataset/CameraTrajDataset.py:643-779
  quat_to_mat / pose_encoding_to_extri_intri  third_party/vggt/vggt/utils/{rotation.py:14-44,pose_enc.py:62-124}
  closed_form_inverse_se3                   third_party/vggt/vggt/utils/geometry.py:114-163
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _euler_rotation(rot_deg: torch.Tensor) -> torch.Tensor:
    """[B,3] degrees (rx,ry,rz) -> R = Rz Ry Rx, [B,3,3]."""
    r = rot_deg * (math.pi / 180.0)
    cx, cy, cz = torch.cos(r[:, 0]), torch.cos(r[:, 1]), torch.cos(r[:, 2])
    sx, sy, sz = torch.sin(r[:, 0]), torch.sin(r[:, 1]), torch.sin(r[:, 2])
    zero, one = torch.zeros_like(cx), torch.ones_like(cx)
    Rx = torch.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], dim=1).view(-1, 3, 3)
    Ry = torch.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], dim=1).view(-1, 3, 3)
    Rz = torch.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], dim=1).view(-1, 3, 3)
    return torch.bmm(Rz, torch.bmm(Ry, Rx))


def _pose_matrix(xyz_euler: torch.Tensor, relative: bool, rows: int) -> torch.Tensor:
    if xyz_euler.dim() != 2 or xyz_euler.shape[1] != 6:
        raise ValueError(f"xyz_euler must be [B, 6], got {tuple(xyz_euler.shape)}")
    B = xyz_euler.shape[0]
    R = _euler_rotation(xyz_euler[:, 3:6])
    t = xyz_euler[:, 0:3].reshape(B, 3, 1)
    if relative:
        R0_inv = R[0:1].transpose(1, 2).expand(B, -1, -1)
        R, t = torch.bmm(R0_inv, R), torch.bmm(R0_inv, t - t[0:1])
    F = torch.cat([R, t], dim=2)
    if rows == 4:
        bottom = torch.tensor([0, 0, 0, 1], dtype=F.dtype, device=F.device).view(1, 1, 4).expand(B, -1, -1)
        F = torch.cat([F, bottom], dim=1)
    return F


def xyz_euler_to_four_by_four_matrix_batch(xyz_euler, relative=False, flatten=False, debug=False, euler_as_rotation=False):
    """[B,6] = [x,y,z,rotx,roty,rotz] (degrees) -> [B,4,4] camera-to-world (or [B,16] if flatten)."""
    B = xyz_euler.size(0)
    F = _pose_matrix(xyz_euler, relative, 4)
    if debug:
        F = F[0].repeat(B, 1, 1)
    if flatten:
        F = F.reshape(B, 16)
    if euler_as_rotation:
        rot = xyz_euler[:, 3:6] * (math.pi / 180.0)
        if relative:
            rot = rot - rot[0:1]
        F = torch.cat([F.reshape(B, 4, 4)[:, :3, 3], rot], dim=1)
    return F


def xyz_euler_to_three_by_four_matrix_batch(xyz_euler, relative=False, flatten=False, debug=False, euler_as_rotation=False):
    """[B,6] -> [B,3,4] camera-to-world (or [B,12] if flatten)."""
    B = xyz_euler.size(0)
    F = _pose_matrix(xyz_euler, relative, 3)
    if debug:
        F = F[0].repeat(B, 1, 1)
    if flatten:
        F = F.reshape(B, 12)
    if euler_as_rotation:
        rot = xyz_euler[:, 3:6] * (math.pi / 180.0)
        if relative:
            rot = rot - rot[0:1]
        F = torch.cat([F.reshape(B, 3, 4)[:, :3, 3], rot], dim=1)
    return F


def quat_to_mat(quaternions: torch.Tensor) -> torch.Tensor:
    """Scalar-last (x,y,z,w) quaternion -> rotation matrix [...,3,3]."""
    i, j, k, r = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def pose_encoding_to_extri_intri(pose_encoding, image_size_hw=None, pose_encoding_type="absT_quaR_FoV", build_intrinsics=True):
    """[B,S,9] = (T, quat xyzw, fov_h, fov_w) -> extrinsics [B,S,3,4] (cam-from-world), intrinsics [B,S,3,3]."""
    if pose_encoding_type != "absT_quaR_FoV":
        raise NotImplementedError(pose_encoding_type)
    T = pose_encoding[..., :3]
    quat = pose_encoding[..., 3:7]
    fov_h = pose_encoding[..., 7]
    fov_w = pose_encoding[..., 8]
    extrinsics = torch.cat([quat_to_mat(quat), T[..., None]], dim=-1)
    intrinsics = None
    if build_intrinsics:
        H, W = image_size_hw
        fy = (H / 2.0) / torch.tan(fov_h / 2.0)
        fx = (W / 2.0) / torch.tan(fov_w / 2.0)
        intrinsics = torch.zeros(pose_encoding.shape[:2] + (3, 3), device=pose_encoding.device, dtype=pose_encoding.dtype)
        intrinsics[..., 0, 0] = fx
        intrinsics[..., 1, 1] = fy
        intrinsics[..., 0, 2] = W / 2
        intrinsics[..., 1, 2] = H / 2
        intrinsics[..., 2, 2] = 1.0
    return extrinsics, intrinsics


def closed_form_inverse_se3(se3, R=None, T=None):
    """Inverse of N x (4x4 | 3x4) rigid transforms: [R^T | -R^T t]."""
    is_numpy = isinstance(se3, np.ndarray)
    if se3.shape[-2:] != (4, 4) and se3.shape[-2:] != (3, 4):
        raise ValueError(f"se3 must be of shape (N,4,4), got {se3.shape}.")
    R = se3[:, :3, :3] if R is None else R
    T = se3[:, :3, 3:] if T is None else T
    if is_numpy:
        Rt = np.transpose(R, (0, 2, 1))
        top_right = -np.matmul(Rt, T)
        inv = np.tile(np.eye(4), (len(R), 1, 1))
    else:
        Rt = R.transpose(1, 2)
        top_right = -torch.bmm(Rt, T)
        inv = torch.eye(4, 4)[None].repeat(len(R), 1, 1).to(R.dtype).to(R.device)
    inv[:, :3, :3] = Rt
    inv[:, :3, 3:] = top_right
    return inv
