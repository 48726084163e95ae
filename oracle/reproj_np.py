"""ORACLE — test infrastructure only (see oracle/__init__.py).

NumPy / torch-CPU restatement of EvoWorld's reprojection path.  Every function cites the reference
file:line it follows.  Pinning status (details in DESIGN.md §Oracle):

  pinned by golden vectors generated from the *imported reference function* (tests/golden/):
      equirectangular_to_ray, ray_c2w_to_plucker, xyz_euler_to_*_matrix_batch,
      unproject_depth_map_to_point_map, pose_encoding_to_extri_intri,
      apply_confidence_filter / extract_colors, cube_to_equirectangular, align_first_and_last_points
  parity UNPINNED (un-vendored third-party dependency; algorithm restated from its published
  behaviour, conventions fixed here):
      equi2pers   (pyequilib==0.5.8, requirements.txt:164)
      splat       (open3d==0.18.0 OffscreenRenderer, requirements.txt:135) -> oracle/reproj_oracle.c
"""
from __future__ import annotations

import ctypes as C
import math
from pathlib import Path

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------
# camera math
# ---------------------------------------------------------------------------------------------


def equirectangular_to_ray(target_H: int = 576, target_W: int = 1024) -> np.ndarray:
    """utils/plucker_embedding.py:56-116."""
    ys = np.arange(target_H, dtype=np.float32)
    xs = np.arange(target_W, dtype=np.float32)
    phi = (xs / target_W - 0.5) * 2.0 * np.pi
    theta = (ys / target_H - 0.5) * np.pi
    Phi, Theta = np.meshgrid(phi, theta)
    cosT, sinT, sinP, cosP = np.cos(Theta), np.sin(Theta), np.sin(Phi), np.cos(Phi)
    return np.stack([cosT * sinP, sinT, cosT * cosP], axis=-1)


def ray_c2w_to_plucker(ray: torch.Tensor, c2w: torch.Tensor) -> torch.Tensor:
    """utils/plucker_embedding.py:221-255 ([direction, moment] channel order, :250)."""
    R = c2w[:, :3, :3].float()
    t = c2w[:, :3, 3].float()
    rays_world = torch.einsum("nij,hwj->nhwi", R, ray.float())
    moment = torch.cross(t[:, None, None, :].expand_as(rays_world), rays_world, dim=-1)
    return torch.cat([rays_world, moment], dim=-1).permute(0, 3, 1, 2)


def euler_to_matrix(xyz_euler: torch.Tensor, relative: bool, four_by_four: bool) -> torch.Tensor:
    """utils/geometry.py:5-88 (4x4) and dataset/CameraTrajDataset.py:643-779 (3x4).
    R = Rz Ry Rx, degrees; relative: R_rel = R0^T R_i, t_rel = R0^T (t_i - t0)."""
    x, y, z, rx, ry, rz = [xyz_euler[:, i].double() for i in range(6)]
    rx, ry, rz = rx * math.pi / 180, ry * math.pi / 180, rz * math.pi / 180
    n = xyz_euler.shape[0]
    Rx = torch.zeros(n, 3, 3, dtype=torch.float64)
    Ry = torch.zeros(n, 3, 3, dtype=torch.float64)
    Rz = torch.zeros(n, 3, 3, dtype=torch.float64)
    Rx[:, 0, 0] = 1; Rx[:, 1, 1] = torch.cos(rx); Rx[:, 1, 2] = -torch.sin(rx); Rx[:, 2, 1] = torch.sin(rx); Rx[:, 2, 2] = torch.cos(rx)
    Ry[:, 1, 1] = 1; Ry[:, 0, 0] = torch.cos(ry); Ry[:, 0, 2] = torch.sin(ry); Ry[:, 2, 0] = -torch.sin(ry); Ry[:, 2, 2] = torch.cos(ry)
    Rz[:, 2, 2] = 1; Rz[:, 0, 0] = torch.cos(rz); Rz[:, 0, 1] = -torch.sin(rz); Rz[:, 1, 0] = torch.sin(rz); Rz[:, 1, 1] = torch.cos(rz)
    R = Rz @ Ry @ Rx
    t = torch.stack([x, y, z], dim=1)[:, :, None]
    if relative:
        R0t = R[0].T
        R, t = R0t[None] @ R, R0t[None] @ (t - t[0:1])
    F = torch.cat([R, t], dim=2)
    if four_by_four:
        bottom = torch.tensor([0, 0, 0, 1], dtype=F.dtype).view(1, 1, 4).expand(n, -1, -1)
        F = torch.cat([F, bottom], dim=1)
    return F.to(xyz_euler.dtype)


def quat_to_mat(q: torch.Tensor) -> torch.Tensor:
    """third_party/vggt/vggt/utils/rotation.py:14-44 (scalar-last xyzw)."""
    i, j, k, r = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


def pose_encoding_to_extri_intri(pose_enc: torch.Tensor, image_size_hw):
    """third_party/vggt/vggt/utils/pose_enc.py:62-124 (absT_quaR_FoV)."""
    T, quat, fov_h, fov_w = pose_enc[..., :3], pose_enc[..., 3:7], pose_enc[..., 7], pose_enc[..., 8]
    R = quat_to_mat(quat)
    extr = torch.cat([R, T[..., None]], dim=-1)
    H, W = image_size_hw
    fy = (H / 2.0) / torch.tan(fov_h / 2.0)
    fx = (W / 2.0) / torch.tan(fov_w / 2.0)
    K = torch.zeros(pose_enc.shape[:2] + (3, 3), dtype=pose_enc.dtype)
    K[..., 0, 0] = fx
    K[..., 1, 1] = fy
    K[..., 0, 2] = W / 2
    K[..., 1, 2] = H / 2
    K[..., 2, 2] = 1.0
    return extr, K


# ---------------------------------------------------------------------------------------------
# equirect -> perspective (pyequilib 0.5.8 Equi2Pers, numpy path) — parity UNPINNED
# ---------------------------------------------------------------------------------------------


def equi2pers_matrix(yaw: float, pitch: float, roll: float, Hp: int, Wp: int, fov_x: float, z_down: bool = False):
    """pix2dir = R * G * K^-1 in float64.  K: f = Wp / (2 tan(fov_x/2)), principal point (Wp/2, Hp/2).
    G maps camera axes (x right, y down, z fwd) to equilib's global axes (x fwd, y right, z down).
    R = Rz(yaw) Ry(pitch) Rx(roll) with pitch, yaw negated when z_down is False."""
    f = Wp / (2.0 * math.tan(math.radians(fov_x) / 2.0))
    K = np.array([[f, 0, Wp / 2.0], [0, f, Hp / 2.0], [0, 0, 1.0]])
    G = np.array([[0, 0, 1.0], [1.0, 0, 0], [0, 1.0, 0]])
    if not z_down:
        pitch, yaw = -pitch, -yaw
    Rx = np.array([[1, 0, 0], [0, math.cos(roll), -math.sin(roll)], [0, math.sin(roll), math.cos(roll)]])
    Ry = np.array([[math.cos(pitch), 0, math.sin(pitch)], [0, 1, 0], [-math.sin(pitch), 0, math.cos(pitch)]])
    Rz = np.array([[math.cos(yaw), -math.sin(yaw), 0], [math.sin(yaw), math.cos(yaw), 0], [0, 0, 1]])
    return (Rz @ Ry @ Rx) @ G @ np.linalg.inv(K)


def equi2pers(equi: np.ndarray, yaw: float, pitch: float = 0.0, roll: float = 0.0, Hp: int = 384, Wp: int = 512,
              fov_x: float = 90.0) -> np.ndarray:
    """equi uint8 [C,He,We] -> uint8 [C,Hp,Wp]; float32 arithmetic in the order the CUDA kernel uses."""
    f32 = np.float32
    C_, He, We = equi.shape
    A = equi2pers_matrix(yaw, pitch, roll, Hp, Wp, fov_x).astype(f32)
    xs = np.arange(Wp, dtype=f32)[None, :]
    ys = np.arange(Hp, dtype=f32)[:, None]
    mx = (A[0, 0] * xs + A[0, 1] * ys) + A[0, 2]
    my = (A[1, 0] * xs + A[1, 1] * ys) + A[1, 2]
    mz = (A[2, 0] * xs + A[2, 1] * ys) + A[2, 2]
    nrm = np.sqrt((mx * mx + my * my) + mz * mz)
    phi = np.arcsin(mz / nrm)
    theta = np.arctan2(my, mx)
    PI = f32(np.pi)
    ui = ((theta - PI) * f32(We)) / (f32(2.0) * PI) + f32(0.5)
    uj = ((phi - f32(0.5) * PI) * f32(He)) / PI + f32(0.5)
    ui = np.fmod(ui, f32(We)); ui = np.where(ui < 0, ui + f32(We), ui).astype(f32)
    uj = np.fmod(uj, f32(He)); uj = np.where(uj < 0, uj + f32(He), uj).astype(f32)
    x0f, y0f = np.floor(ui), np.floor(uj)
    dx, dy = (ui - x0f).astype(f32), (uj - y0f).astype(f32)
    x0 = x0f.astype(np.int64) % We
    y0 = y0f.astype(np.int64) % He
    x1, y1 = (x0 + 1) % We, (y0 + 1) % He
    img = equi.astype(f32)
    wx0, wy0 = f32(1.0) - dx, f32(1.0) - dy
    top = img[:, y0, x0] * wx0 + img[:, y0, x1] * dx
    bot = img[:, y1, x0] * wx0 + img[:, y1, x1] * dx
    v = top * wy0 + bot * dy
    return np.clip(v, 0, 255).astype(np.uint8)


def equi2pers_yaw(equi: np.ndarray, yaw: float, Hp: int = 384, Wp: int = 512, fov_x: float = 90.0) -> np.ndarray:
    """Pure-yaw form of `equi2pers` (pitch = roll = 0, EvoWorld's only call): the (ui, uj) of the yaw = 0 camera, then the
    yaw as a longitude shift -yaw We / 2pi in source pixels (reduced to [-We/2, We/2]); float32 in the kernel's order."""
    f32 = np.float32
    C_, He, We = equi.shape
    A = equi2pers_matrix(0.0, 0.0, 0.0, Hp, Wp, fov_x).astype(f32)
    xs = np.arange(Wp, dtype=f32)[None, :]
    ys = np.arange(Hp, dtype=f32)[:, None]
    mx = (A[0, 0] * xs + A[0, 1] * ys) + A[0, 2]
    my = (A[1, 0] * xs + A[1, 1] * ys) + A[1, 2]
    mz = (A[2, 0] * xs + A[2, 1] * ys) + A[2, 2]
    nrm = np.sqrt((mx * mx + my * my) + mz * mz)
    phi = np.arcsin(mz / nrm)
    theta = np.arctan2(my, mx)
    PI = f32(np.pi)
    ui0 = ((theta - PI) * f32(We)) / (f32(2.0) * PI) + f32(0.5)
    uj = ((phi - f32(0.5) * PI) * f32(He)) / PI + f32(0.5)
    uj = np.fmod(uj, f32(He)); uj = np.where(uj < 0, uj + f32(He), uj).astype(f32)
    shift = f32(math.remainder(-yaw, 2.0 * math.pi) * We / (2.0 * math.pi))
    ui = (ui0 + shift).astype(f32)
    ui = np.fmod(ui, f32(We)); ui = np.where(ui < 0, ui + f32(We), ui).astype(f32)
    x0f, y0f = np.floor(ui), np.floor(uj)
    dx, dy = (ui - x0f).astype(f32), (uj - y0f).astype(f32)
    x0 = x0f.astype(np.int64) % We
    y0 = y0f.astype(np.int64) % He
    x1, y1 = (x0 + 1) % We, (y0 + 1) % He
    img = equi.astype(f32)
    wx0, wy0 = f32(1.0) - dx, f32(1.0) - dy
    top = img[:, y0, x0] * wx0 + img[:, y0, x1] * dx
    bot = img[:, y1, x0] * wx0 + img[:, y1, x1] * dx
    v = top * wy0 + bot * dy
    return np.clip(v, 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# lift / filter
# ---------------------------------------------------------------------------------------------


def unproject_depth_map_to_point_map(depth: np.ndarray, extr: np.ndarray, intr: np.ndarray) -> np.ndarray:
    """third_party/vggt/vggt/utils/geometry.py:12-111 (float64 output; cam coords rounded to f32 :109)."""
    if depth.ndim == 4:
        depth = depth[..., 0]
    S, H, W = depth.shape
    out = np.empty((S, H, W, 3), dtype=np.float64)
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    for s in range(S):
        K, E = intr[s], extr[s]
        x = (u - K[0, 2]) * depth[s] / K[0, 0]
        y = (v - K[1, 2]) * depth[s] / K[1, 1]
        cam = np.stack((x, y, depth[s]), axis=-1).astype(np.float32)
        Rt = np.transpose(E[:3, :3])
        tr = -np.matmul(Rt, E[:3, 3:])  # float32 when extr is float32 (geometry.py:151)
        c2w = np.eye(4)
        c2w[:3, :3] = Rt
        c2w[:3, 3:] = tr
        out[s] = np.dot(cam, c2w[:3, :3].T) + c2w[:3, 3]
    return out


def extract_colors(images: np.ndarray) -> np.ndarray:
    """reproject_vggt_open3d_utils.py:286-292 (truncating cast)."""
    if images.ndim == 4 and images.shape[1] == 3:
        images = np.transpose(images, (0, 2, 3, 1))
    return (images.reshape(-1, 3) * 255).astype(np.uint8)


def apply_confidence_filter(points: np.ndarray, conf: np.ndarray, colors: np.ndarray, conf_thres: float):
    """reproject_vggt_open3d_utils.py:294-310."""
    conf_flat = conf.reshape(-1)
    thr = 0.0 if conf_thres == 0.0 else np.percentile(conf_flat, conf_thres)
    mask = conf_flat >= thr
    if not np.any(mask):
        return np.array([[1, 0, 0]]), np.array([[255, 255, 255]])
    return points.reshape(-1, 3)[mask], colors[mask]


# ---------------------------------------------------------------------------------------------
# target cameras
# ---------------------------------------------------------------------------------------------

CUBEMAP_TRANSFORMS = {  # reproject_vggt_open3d_utils.py:29-36
    "front": np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]),
    "right": np.array([[0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 0, 1]]),
    "back": np.array([[-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]]),
    "left": np.array([[0, 0, -1, 0], [0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1]]),
    "top": np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]),
    "bottom": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]]),
}
FACE_ORDER = list(CUBEMAP_TRANSFORMS.keys())


def rotation_from_vectors(u, v):
    """reproject_vggt_open3d_utils.py:1126-1174."""
    nu, nv = np.linalg.norm(u), np.linalg.norm(v)
    if nu < 1e-15 or nv < 1e-15:
        return np.eye(3)
    uh, vh = u / nu, v / nv
    dot = np.clip(np.dot(uh, vh), -1.0, 1.0)
    if np.isclose(dot, 1.0):
        return np.eye(3)
    if np.isclose(dot, -1.0):
        tmp = np.array([1.0, 0.0, 0.0])
        if np.abs(np.dot(uh, tmp)) > 0.9:
            tmp = np.array([0.0, 1.0, 0.0])
        w = np.cross(uh, tmp)
        w /= np.linalg.norm(w)
        return np.eye(3) - 2.0 * np.outer(w, w)
    angle = np.arccos(dot)
    w = np.cross(uh, vh)
    wh = w / np.linalg.norm(w)
    K = np.array([[0, -wh[2], wh[1]], [wh[2], 0, -wh[0]], [-wh[1], wh[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1.0 - np.cos(angle)) * (K @ K)


def align_first_and_last_points(A, B):
    """reproject_vggt_open3d_utils.py:1176-1213."""
    A0, A1, B0, B1 = A[0], A[-1], B[0], B[-1]
    vA, vB = A1 - A0, B1 - B0
    lenA, lenB = np.linalg.norm(vA), np.linalg.norm(vB)
    if lenA < 1e-15:
        return 1.0, np.eye(3), B0 - A0
    s = lenB / lenA
    R = rotation_from_vectors(vA, vB)
    return s, R, B0 - s * R @ A0


def align_extrinsics(camera_pose, predictions_extrinsic, num_target_view, outdir, only_render_last_24_frame=False):
    """SceneBuilder.align_extrinsics, reproject_vggt_open3d_utils.py:472-519."""
    n = len(predictions_extrinsic)
    E = np.zeros((n, 4, 4))
    E[:, :3, :4] = predictions_extrinsic
    E[:, 3, 3] = 1
    Einv = np.stack([np.linalg.inv(e) for e in E])
    try:
        segment_id = int(outdir.rstrip("/").split("_")[-1])
    except Exception:
        segment_id = 1
    start = (segment_id + 1) * num_target_view + 1 if not only_render_last_24_frame else -num_target_view
    gt = np.asarray(camera_pose[:start])
    tgt = np.asarray(camera_pose[start:start + num_target_view] if not only_render_last_24_frame else camera_pose[start:])
    s, R, t = align_first_and_last_points(gt[:, :3, 3], Einv[:, :3, 3])
    Tm = np.eye(4)
    Tm[:3, :3] = s * R
    Tm[:3, 3] = t
    return np.einsum("ij, bjk -> bik", Tm, tgt)


def face_w2c(target_c2w: np.ndarray) -> np.ndarray:
    """render_cubemap / render_face, reproject_vggt_open3d_utils.py:617-666:
    cam = c2w @ T_face, top/bottom additionally @ Rz(180 deg); extrinsic = inv(cam).  -> [V,6,3,4] f64"""
    Fz = np.eye(4)
    Fz[:3, :3] = np.array([[-1.0, 0, 0], [0, -1.0, 0], [0, 0, 1.0]])
    V = target_c2w.shape[0]
    out = np.zeros((V, 6, 3, 4))
    for v in range(V):
        for fi, name in enumerate(FACE_ORDER):
            cam = target_c2w[v] @ CUBEMAP_TRANSFORMS[name]
            if name in ("top", "bottom"):
                cam = cam @ Fz
            out[v, fi] = np.linalg.inv(cam)[:3, :4]
    return out


# ---------------------------------------------------------------------------------------------
# cube -> equirect (reproject_vggt_open3d_utils.py:542-614), torch CPU float32 like the reference
# ---------------------------------------------------------------------------------------------


def cube_to_equirect_lut(width: int, height: int, face_res: int) -> np.ndarray:
    """Index table equivalent to cube_to_equirectangular_cuda: uint32 (face<<28 | row<<14 | col),
    face numbered in FACE_ORDER; 0xFFFFFFFF where no mask matches (never, except NaN)."""
    x = torch.linspace(0, width - 1, width)
    y = torch.linspace(0, height - 1, height)
    xv, yv = torch.meshgrid(y, x, indexing="ij")
    lon = (-yv / width) * 2 * torch.pi - torch.pi + torch.pi / 2
    lat = (xv / height) * torch.pi - torch.pi / 2
    X = torch.cos(lat) * torch.cos(lon)
    Y = torch.sin(lat)
    Z = torch.cos(lat) * torch.sin(lon)
    aX, aY, aZ = X.abs(), Y.abs(), Z.abs()
    masks = {
        "right": (aX >= aY) & (aX >= aZ) & (X > 0),
        "left": (aX >= aY) & (aX >= aZ) & (X < 0),
        "bottom": (aY >= aX) & (aY >= aZ) & (Y > 0),
        "top": (aY >= aX) & (aY >= aZ) & (Y < 0),
        "front": (aZ >= aX) & (aZ >= aY) & (Z > 0),
        "back": (aZ >= aX) & (aZ >= aY) & (Z < 0),
    }
    face = torch.full((height, width), -1, dtype=torch.int64)
    u = torch.zeros_like(X)
    v = torch.zeros_like(Y)
    for f, m in masks.items():  # later masks overwrite earlier ones, as in the reference
        face[m] = FACE_ORDER.index(f)
        if f in ("right", "left"):
            u[m] = -Z[m] / aX[m] if f == "right" else Z[m] / aX[m]
            v[m] = -Y[m] / aX[m]
        elif f in ("bottom", "top"):
            u[m] = -X[m] / aY[m]
            v[m] = -Z[m] / aY[m] if f == "bottom" else Z[m] / aY[m]
        else:
            u[m] = X[m] / aZ[m] if f == "front" else -X[m] / aZ[m]
            v[m] = -Y[m] / aZ[m]
    u = (u + 1) / 2
    v = (v + 1) / 2
    u_px = (u * (face_res - 1)).long()
    v_px = ((1 - v) * (face_res - 1)).long()
    lut = (face << 28) | (v_px << 14) | u_px
    lut[face < 0] = 0xFFFFFFFF
    return lut.numpy().astype(np.uint32)


def cube_to_equirectangular(cube_faces: dict, width: int, height: int) -> np.ndarray:
    """cube_faces[name] uint8 [B,3,res,res] -> uint8 [B,height,width,3] through the table above."""
    res = cube_faces["front"].shape[-1]
    lut = cube_to_equirect_lut(width, height, res).astype(np.int64)
    stack = np.stack([np.asarray(cube_faces[n]) for n in FACE_ORDER], axis=1)  # [B,6,3,res,res]
    f, r, c = lut >> 28, (lut >> 14) & 0x3FFF, lut & 0x3FFF
    return stack[:, f, :, r, c].transpose(2, 0, 1, 3)  # advanced dims first: [H,W,B,3] -> [B,H,W,3]


# ---------------------------------------------------------------------------------------------
# splat (C oracle wrapper)
# ---------------------------------------------------------------------------------------------

_ORACLE_LIB = None


def _clib():
    global _ORACLE_LIB
    if _ORACLE_LIB is None:
        p = Path(__file__).resolve().parent / "_build" / "libreproj_oracle.so"
        if not p.exists():
            from evoworld_b200.build import build_oracle

            build_oracle()
        _ORACLE_LIB = C.CDLL(str(p))
    return _ORACLE_LIB


def pack_points(xyz: np.ndarray, rgb: np.ndarray) -> np.ndarray:
    """float4 {x,y,z,bits(r|g<<8|b<<16)}; xyz is rounded to float32 as Open3D does on upload."""
    n = xyz.shape[0]
    out = np.empty((n, 4), dtype=np.float32)
    out[:, :3] = xyz.astype(np.float32)
    rgb = rgb.astype(np.uint32)
    out[:, 3] = (rgb[:, 0] | (rgb[:, 1] << 8) | (rgb[:, 2] << 16)).astype(np.uint32).view(np.float32)
    return out


def splat_keys(pts4: np.ndarray, w2c: np.ndarray, res: int, focal: float, z_near: float) -> np.ndarray:
    pts4 = np.ascontiguousarray(pts4, dtype=np.float32)
    w2c = np.ascontiguousarray(w2c, dtype=np.float32)
    V = w2c.shape[0]
    keys = np.empty((V, 6, res, res), dtype=np.uint64)
    _clib().oracle_splat_keys(
        pts4.ctypes.data_as(C.c_void_p), C.c_int64(pts4.shape[0]), w2c.ctypes.data_as(C.c_void_p), C.c_int(V),
        C.c_int(res), C.c_float(focal), C.c_float(z_near), keys.ctypes.data_as(C.c_void_p))
    return keys


def front_w2c(target_c2w: np.ndarray) -> np.ndarray:
    """[V,4,4] target camera-to-world -> [V,3,4] float32 cam-from-world of the FRONT cube face (T_front = I)."""
    return np.stack([np.linalg.inv(c)[:3, :4] for c in np.asarray(target_c2w, dtype=np.float64)]).astype(np.float32)


def splat_keys_cube(pts4: np.ndarray, w2c_front: np.ndarray, res: int, focal: float, z_near: float,
                    color_keys: bool = False) -> np.ndarray:
    """Cube formulation (oracle_splat_keys_cube): one transform per (point, view), face = major axis.
    color_keys: the low key word carries the packed colour instead of the point index (oracle_splat_keys_cube_colorkey)."""
    pts4 = np.ascontiguousarray(pts4, dtype=np.float32)
    w2c_front = np.ascontiguousarray(w2c_front, dtype=np.float32)
    V = w2c_front.shape[0]
    keys = np.empty((V, 6, res, res), dtype=np.uint64)
    fn = _clib().oracle_splat_keys_cube_colorkey if color_keys else _clib().oracle_splat_keys_cube
    fn(
        pts4.ctypes.data_as(C.c_void_p), C.c_int64(pts4.shape[0]), w2c_front.ctypes.data_as(C.c_void_p), C.c_int(V),
        C.c_int(res), C.c_float(focal), C.c_float(z_near), keys.ctypes.data_as(C.c_void_p))
    return keys


def keys_to_index(keys: np.ndarray) -> np.ndarray:
    idx = (keys & np.uint64(0xFFFFFFFF)).astype(np.int64)
    idx[keys == np.uint64(0xFFFFFFFFFFFFFFFF)] = -1
    return idx


def resolve(keys: np.ndarray, pts4: np.ndarray, lut: np.ndarray, color_keys: bool = False) -> np.ndarray:
    V, _, res, _ = keys.shape
    H, W = lut.shape
    out = np.empty((V, H, W, 3), dtype=np.uint8)
    lut = np.ascontiguousarray(lut, dtype=np.uint32)
    pts4 = np.ascontiguousarray(pts4, dtype=np.float32)
    if color_keys:
        _clib().oracle_resolve_colorkey(keys.ctypes.data_as(C.c_void_p), lut.ctypes.data_as(C.c_void_p), C.c_int(V), C.c_int(res),
                                        C.c_int64(H * W), out.ctypes.data_as(C.c_void_p))
        return out
    _clib().oracle_resolve(
        keys.ctypes.data_as(C.c_void_p), pts4.ctypes.data_as(C.c_void_p), lut.ctypes.data_as(C.c_void_p), C.c_int(V),
        C.c_int(res), C.c_int64(H * W), out.ctypes.data_as(C.c_void_p))
    return out


def render_panoramas(xyz: np.ndarray, rgb: np.ndarray, target_c2w: np.ndarray, res: int = 512, width: int = 2000,
                     height: int = 1000, z_near: float = 1e-6) -> np.ndarray:
    """render_cubemaps_to_panoramas (reproject_vggt_open3d_utils.py:668-711) on the oracle."""
    pts4 = pack_points(xyz, rgb)
    w2c = face_w2c(target_c2w).astype(np.float32)
    keys = splat_keys(pts4, w2c, res, res / 2.0, z_near)
    return resolve(keys, pts4, cube_to_equirect_lut(width, height, res))


def render_panoramas_cube(xyz: np.ndarray, rgb: np.ndarray, target_c2w: np.ndarray, res: int = 512, width: int = 2000,
                          height: int = 1000, z_near: float = 1e-6, color_keys: bool = False) -> np.ndarray:
    """render_cubemaps_to_panoramas on the oracle, cube formulation (what the product's fast path computes).
    color_keys selects the colour-key tie rule of the optional EVW_SPLAT_COLOR_KEYS mode."""
    pts4 = pack_points(xyz, rgb)
    keys = splat_keys_cube(pts4, front_w2c(target_c2w), res, res / 2.0, z_near, color_keys)
    return resolve(keys, pts4, cube_to_equirect_lut(width, height, res), color_keys)
