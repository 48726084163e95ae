"""Equirectangular -> perspective warp (operator boundary 4).

Drop-in for `equilib.Equi2Pers` (pyequilib 0.5.8) as the reference constructs and calls it
(unified_loop_consistency.py:178-183,329; pano_to_pers.py:176; pano_to_pers_per_segment.py:203):
    Equi2Pers(height=384, width=512, fov_x=90, mode="bilinear")(equi=uint8[3,He,We], rots={...})
The sampling runs in evw_equi2pers_u8 (csrc/reproj.cu); only the 3x3 pixel->direction matrix is
composed on the host (float64).  pyequilib is not vendored in the reference; the conventions are
restated in DESIGN.md and oracle/reproj_np.py (parity unpinned).
"""
from __future__ import annotations

import math
from typing import Dict, List, Union

import numpy as np
import torch

from . import _lib


def pix2dir_matrix(yaw: float, pitch: float, roll: float, height: int, width: int, fov_x: float, skew: float = 0.0,
                   z_down: bool = False) -> np.ndarray:
    """M = R G K^-1: homogeneous output pixel (x, y, 1) -> direction in the global frame
    (x forward, y right, z down).  R = Rz(yaw) Ry(pitch) Rx(roll); pitch and yaw change sign when
    z_down is False (pyequilib's default)."""
    f = width / (2.0 * math.tan(math.radians(fov_x) / 2.0))
    K = np.array([[f, skew, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)
    G = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=np.float64)
    if not z_down:
        pitch, yaw = -pitch, -yaw
    cr, sr = math.cos(roll), math.sin(roll)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cy, sy = math.cos(yaw), math.sin(yaw)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]], dtype=np.float64)
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]], dtype=np.float64)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]], dtype=np.float64)
    return (Rz @ Ry @ Rx) @ G @ np.linalg.inv(K)


def pix2dir_matrices(rots, height: int, width: int, fov_x: float, skew: float = 0.0, z_down: bool = False) -> np.ndarray:
    """Batched pix2dir_matrix: one [B,3,3] float64 array for a list of rotation dicts (same expression order per
    element, so every matrix equals the scalar function's bit for bit); avoids B small numpy products per call."""
    B = len(rots)
    yaw = np.array([r.get("yaw", 0.0) for r in rots], dtype=np.float64)
    pitch = np.array([r.get("pitch", 0.0) for r in rots], dtype=np.float64)
    roll = np.array([r.get("roll", 0.0) for r in rots], dtype=np.float64)
    if not z_down:
        pitch, yaw = -pitch, -yaw
    f = width / (2.0 * math.tan(math.radians(fov_x) / 2.0))
    K = np.array([[f, skew, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]], dtype=np.float64)
    G = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=np.float64)
    Kinv = np.linalg.inv(K)
    one, zero = np.ones(B), np.zeros(B)
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    Rx = np.stack([one, zero, zero, zero, cr, -sr, zero, sr, cr], axis=1).reshape(B, 3, 3)
    Ry = np.stack([cp, zero, sp, zero, one, zero, -sp, zero, cp], axis=1).reshape(B, 3, 3)
    Rz = np.stack([cy, -sy, zero, sy, cy, zero, zero, zero, one], axis=1).reshape(B, 3, 3)
    return (((Rz @ Ry) @ Rx) @ G) @ Kinv  # same association as pix2dir_matrix


class Equi2Pers:
    """Callable with pyequilib's constructor/`__call__` signature; uint8 bilinear only (the
    configuration EvoWorld uses).  Accepts numpy (host) or torch CUDA `equi`; returns the same kind."""

    def __init__(self, height: int, width: int, fov_x: float, skew: float = 0.0, z_down: bool = False,
                 mode: str = "bilinear", clip_output: bool = True, device: Union[str, torch.device] = "cuda",
                 fast_yaw: bool = True):
        if mode != "bilinear":
            raise NotImplementedError(f"Equi2Pers mode {mode!r}: only 'bilinear' is built (EvoWorld's setting)")
        self.height, self.width, self.fov_x, self.skew, self.z_down = height, width, fov_x, skew, z_down
        self.mode = mode
        self.device = torch.device(device)
        self.fast_yaw = fast_yaw  # pitch = roll = 0 -> tabulated yaw = 0 camera + longitude shift (False: always the general kernel)
        self._tables = {}  # (device, He, We) -> float2 table of the yaw = 0 camera (pure-yaw fast path)

    def _table(self, dev, He: int, We: int) -> torch.Tensor:
        key = (str(dev), He, We)
        if key not in self._tables:
            m0 = pix2dir_matrix(0.0, 0.0, 0.0, self.height, self.width, self.fov_x, self.skew, self.z_down).astype(np.float32)
            m0 = torch.from_numpy(m0.reshape(9)).to(dev)
            table = torch.empty((self.height * self.width, 2), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().evw_equi2pers_table(_lib.ptr(m0), _lib.ptr(table), He, We, self.height, self.width,
                                                          _lib.stream_ptr(dev)), "evw_equi2pers_table")
            self._tables[key] = table
        return self._tables[key]

    def yaw_shift_px(self, yaw: float, We: int) -> float:
        """Longitude shift of a yaw rotation in source pixels: R = Rz(-yaw) for z_down = False (pyequilib's default),
        reduced to [-We/2, We/2] so that ui0 + shift stays within one wrap."""
        a = yaw if self.z_down else -yaw
        a = math.remainder(a, 2.0 * math.pi)
        return a * We / (2.0 * math.pi)

    def __call__(self, equi, rots: Union[Dict[str, float], List[Dict[str, float]]], **_):
        is_numpy = isinstance(equi, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(equi)) if is_numpy else equi
        if t.dtype != torch.uint8:
            raise NotImplementedError("Equi2Pers: only uint8 images are supported (EvoWorld's usage)")
        single = t.dim() == 3
        if single:
            t = t[None]
            rots = [rots] if isinstance(rots, dict) else rots
        if t.dim() != 4 or len(rots) != t.shape[0]:
            raise ValueError("Equi2Pers: equi must be [C,H,W] with one rots dict or [B,C,H,W] with B dicts")
        dev = t.device if t.is_cuda else self.device
        t = t.to(dev, non_blocking=True).contiguous()
        B, C, He, We = t.shape
        out = torch.empty((B, C, self.height, self.width), dtype=torch.uint8, device=dev)
        pure_yaw = self.fast_yaw and all(r.get("pitch", 0.0) == 0.0 and r.get("roll", 0.0) == 0.0 for r in rots)
        if pure_yaw:  # EvoWorld's only usage: a table look-up + longitude shift per frame
            shift = torch.tensor([self.yaw_shift_px(float(r.get("yaw", 0.0)), We) for r in rots], dtype=torch.float32).to(dev)
            table = self._table(dev, He, We)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().evw_equi2pers_yaw_u8(_lib.ptr(t), _lib.ptr(table), _lib.ptr(shift), _lib.ptr(out), B, C, He,
                                                           We, self.height, self.width, _lib.stream_ptr(dev)), "evw_equi2pers_yaw_u8")
        else:
            mats = pix2dir_matrices(rots, self.height, self.width, self.fov_x, self.skew, self.z_down).astype(np.float32)
            m = torch.from_numpy(mats.reshape(B, 9)).to(dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().evw_equi2pers_u8(_lib.ptr(t), _lib.ptr(m), _lib.ptr(out), B, C, He, We, self.height,
                                                       self.width, _lib.stream_ptr(dev)), "evw_equi2pers_u8")
        if single:
            out = out[0]
        return out.cpu().numpy() if is_numpy else out
