"""Seeded synthetic workloads of the benchmark shapes; no reference data needed.

Used by bench.py, __graft_entry__.smoke() and the GPU parity tests so that every leg (CUDA path,
oracle, CPU baseline) sees identical inputs.
"""
from __future__ import annotations

import math

import numpy as np


def curve_trajectory(n: int = 126, radius: float = 6.0, arc_deg: float = 100.0, height: float = 1.78) -> np.ndarray:
    """[n,6] = [x,y,z,rx,ry,rz] (OpenCV RDF, degrees): a camera walking along a circular arc with the
    yaw following the tangent — the shape of the reference's curve-path episodes."""
    a = np.radians(np.linspace(0.0, arc_deg, n))
    x = radius * (1.0 - np.cos(a))
    z = radius * np.sin(a)
    yaw = np.degrees(a)
    out = np.zeros((n, 6), dtype=np.float64)
    out[:, 0], out[:, 1], out[:, 2], out[:, 4] = x, -height, z, yaw
    return out


def euler_c2w(poses: np.ndarray) -> np.ndarray:
    """[n,6] -> [n,4,4] float64 camera-to-world, R = Rz Ry Rx (degrees), relative to the first frame."""
    n = poses.shape[0]
    out = np.tile(np.eye(4), (n, 1, 1))
    for i, (x, y, z, rx, ry, rz) in enumerate(poses):
        rx, ry, rz = math.radians(rx), math.radians(ry), math.radians(rz)
        Rx = np.array([[1, 0, 0], [0, math.cos(rx), -math.sin(rx)], [0, math.sin(rx), math.cos(rx)]])
        Ry = np.array([[math.cos(ry), 0, math.sin(ry)], [0, 1, 0], [-math.sin(ry), 0, math.cos(ry)]])
        Rz = np.array([[math.cos(rz), -math.sin(rz), 0], [math.sin(rz), math.cos(rz), 0], [0, 0, 1]])
        out[i, :3, :3] = Rz @ Ry @ Rx
        out[i, :3, 3] = (x, y, z)
    inv0 = np.linalg.inv(out[0])
    return inv0[None] @ out


def reprojection_predictions(S: int = 25, H: int = 392, W: int = 518, seed: int = 0, scale: float = 0.37) -> dict:
    """A VGGT-shaped `predictions` dict (numpy): depth, depth_conf, images, extrinsic, intrinsic and
    camera_pose (GT, [126,4,4]) such that the VGGT frame is a similarity transform of the GT frame."""
    rng = np.random.default_rng(seed)
    depth = np.exp(rng.normal(0.7, 0.5, size=(S, H, W, 1))).astype(np.float32)
    conf = (1.0 + np.exp(rng.normal(0.0, 1.0, size=(S, H, W)))).astype(np.float32)
    images = rng.random((S, 3, H, W), dtype=np.float32)
    cam = euler_c2w(curve_trajectory())
    c2w = cam[:S].copy()
    c2w[:, :3, 3] *= scale
    extr = np.linalg.inv(c2w)[:, :3, :4].astype(np.float32)
    intr = np.tile(np.array([[W / 2.0, 0, W / 2.0], [0, W / 2.0, H / 2.0], [0, 0, 1]], dtype=np.float32), (S, 1, 1))
    return {"depth": depth, "depth_conf": conf, "images": images, "extrinsic": extr, "intrinsic": intr,
            "camera_pose": cam}


def random_cloud(n: int, seed: int = 0, radius: float = 4.0):
    """xyz float64 [n,3] roughly uniform in a ball shell around the origin + uint8 colours."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = radius * (0.3 + 0.7 * rng.random(n)) ** (1 / 3)
    xyz = d * r[:, None]
    rgb = rng.integers(0, 256, size=(n, 3), dtype=np.uint8)
    return xyz, rgb
