"""3D-memory reprojection (operator boundary 6).

Drop-in for evoworld/reprojection/reproject_vggt_open3d_utils.py of the reference — same class and
function names, argument meaning and error behaviour for the depth-unproject / Open3D branch:

    PointCloudProcessor.filter_predictions   :174-222  -> evw_pack_points + evw_conf_select
    SceneBuilder.build_open3d_scene          :457-470  -> device-resident PointScene (no GL context)
    SceneBuilder.align_extrinsics            :472-519  -> host numpy float64 (negligible, kept exact)
    CubemapRenderer.render_face/_cubemap     :617-666  -> evw_splat_faces_u8
    CubemapRenderer.cube_to_equirectangular_cuda :542-614 -> evw_cube_to_equirect_u8 (lookup table)
    CubemapRenderer.render_cubemaps_to_panoramas :668-711 -> evw_splat_cubemap_equirect (fused)
    predictions_to_target_view               :1216-1282

Unlike the reference, constructing the classes (and importing this module) triggers no ONNX model
download and no GL context.  Sky segmentation, background masks and the trimesh/GLB export are not
on the hot path and raise NotImplementedError (out of scope).
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib

logger = logging.getLogger(__name__)

CUBEMAP_TRANSFORMS = {
    "front": np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]),
    "right": np.array([[0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 0, 1]]),
    "back": np.array([[-1, 0, 0, 0], [0, 1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]]),
    "left": np.array([[0, 0, -1, 0], [0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1]]),
    "top": np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]),
    "bottom": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1]]),
}
CUBEMAP = CUBEMAP_TRANSFORMS
FACE_ORDER = tuple(CUBEMAP_TRANSFORMS.keys())  # z-buffer / lookup-table face numbering

FACE_RES = 512          # reproject_vggt_open3d_utils.py:617,636
PANO_W, PANO_H = 2000, 1000  # :705
Z_NEAR = 1e-6           # near plane of the point pass
DEFAULT_VIEWS_PER_PASS = 2  # 0.54 ms per 24-view set vs 0.57 (G=4) and 0.65 (G=8) with the two-stream pipeline (profiles/r01g_reproj_bench.log)
DEFAULT_PRETEST = os.environ.get("EVW_SPLAT_PRETEST", "0") != "0"  # read the cell before the 64-bit atomic min (slower on B200)
DEFAULT_OVERLAP = os.environ.get("EVW_SPLAT_OVERLAP", "1") != "0"  # two-stream pass pipeline (include/evoworld_b200.h)
DEFAULT_BY_ROLE = os.environ.get("EVW_SPLAT_BY_ROLE", "0") != "0"  # alternative overlap scheme (splat stream / resolve stream)
DEFAULT_COLOR_KEYS = os.environ.get("EVW_SPLAT_COLOR_KEYS", "0") != "0"  # optional tie rule (include/evoworld_b200.h); off
SPLAT_PRETEST, SPLAT_OVERLAP, SPLAT_OVERLAP_BY_ROLE, SPLAT_COLOR_KEYS = 1, 2, 8, 16  # bit 4 was the round-1 v1-kernel switch


def splat_flags(pretest: Optional[bool] = None, overlap: Optional[bool] = None, by_role: Optional[bool] = None,
                color_keys: Optional[bool] = None) -> int:
    f = SPLAT_PRETEST if (DEFAULT_PRETEST if pretest is None else pretest) else 0
    f |= SPLAT_OVERLAP if (DEFAULT_OVERLAP if overlap is None else overlap) else 0
    f |= SPLAT_OVERLAP_BY_ROLE if (DEFAULT_BY_ROLE if by_role is None else by_role) else 0
    f |= SPLAT_COLOR_KEYS if (DEFAULT_COLOR_KEYS if color_keys is None else color_keys) else 0
    return f


# ---------------------------------------------------------------------------------------------
# host-side scalar helpers
# ---------------------------------------------------------------------------------------------


def percentile_rank_params(n: int, q: float, dtype=np.float32) -> Tuple[int, int, float]:
    """(k_lo, k_hi, gamma) of numpy's 'linear' percentile for an n-element array of `dtype`,
    evaluated with numpy's own scalar arithmetic (numpy/lib/_function_base_impl.py: the quantile is
    divided by dtype(100), the virtual index (n-1)*q stays in that dtype).  The device then only has
    to find two order statistics and lerp them."""
    dt = np.dtype(dtype).type
    qf = np.true_divide(q, dt(100))          # weak python float -> dtype (NEP 50)
    vi = (n - 1) * qf                        # virtual index in dtype
    prev = int(np.floor(vi))
    nxt = prev + 1
    if vi >= n - 1:
        prev = nxt = n - 1
    if vi < 0:
        prev = nxt = 0
    prev = min(max(prev, 0), n - 1)
    nxt = min(max(nxt, 0), n - 1)
    gamma = float(dt(vi - np.floor(vi)))
    return prev, nxt, gamma


_LUT_CACHE: Dict[Tuple[int, int, int, str], torch.Tensor] = {}


def build_cube_lut(width: int, height: int, face_res: int) -> np.ndarray:
    """Cube->equirect lookup table, uint32 [height,width] = face<<28 | row<<14 | col.

    Evaluates the reference's own expression sequence (reproject_vggt_open3d_utils.py:542-614) with
    torch on the host in float32 so that face choice (`>=` ties, last mask wins) and the truncated
    pixel indices are the reference's; it depends only on (width, height, face_res) and is cached."""
    x = torch.linspace(0, width - 1, width)
    y = torch.linspace(0, height - 1, height)
    rows, cols = torch.meshgrid(y, x, indexing="ij")  # the reference calls these xv, yv
    lon = (-cols / width) * 2 * torch.pi - torch.pi + torch.pi / 2
    lat = (rows / height) * torch.pi - torch.pi / 2
    X = torch.cos(lat) * torch.cos(lon)
    Y = torch.sin(lat)
    Z = torch.cos(lat) * torch.sin(lon)
    aX, aY, aZ = X.abs(), Y.abs(), Z.abs()
    order = [  # mask evaluation order of the reference; later entries overwrite earlier ones
        ("right", (aX >= aY) & (aX >= aZ) & (X > 0), -Z / aX, -Y / aX),
        ("left", (aX >= aY) & (aX >= aZ) & (X < 0), Z / aX, -Y / aX),
        ("bottom", (aY >= aX) & (aY >= aZ) & (Y > 0), -X / aY, -Z / aY),
        ("top", (aY >= aX) & (aY >= aZ) & (Y < 0), -X / aY, Z / aY),
        ("front", (aZ >= aX) & (aZ >= aY) & (Z > 0), X / aZ, -Y / aZ),
        ("back", (aZ >= aX) & (aZ >= aY) & (Z < 0), -X / aZ, -Y / aZ),
    ]
    face = torch.full((height, width), -1, dtype=torch.int64)
    u = torch.zeros_like(X)
    v = torch.zeros_like(X)
    for name, mask, uu, vv in order:
        face[mask] = FACE_ORDER.index(name)
        u[mask] = uu[mask]
        v[mask] = vv[mask]
    u = (u + 1) / 2
    v = (v + 1) / 2
    u_px = (u * (face_res - 1)).long()
    v_px = ((1 - v) * (face_res - 1)).long()
    lut = (face << 28) | (v_px << 14) | u_px
    lut[face < 0] = 0xFFFFFFFF
    return lut.numpy().astype(np.uint32)


def cube_lut_device(width: int, height: int, face_res: int, device) -> torch.Tensor:
    key = (width, height, face_res, str(device))
    if key not in _LUT_CACHE:
        lut = build_cube_lut(width, height, face_res)
        _LUT_CACHE[key] = torch.from_numpy(lut.view(np.int32)).to(device)
    return _LUT_CACHE[key]


def face_w2c_matrices(target_c2w: np.ndarray) -> np.ndarray:
    """[V,4,4] target camera-to-world -> [V,6,3,4] float32 cam-from-world per cube face:
    cam = c2w @ T_face (top/bottom additionally @ Rz(180 deg)), w2c = inv(cam)  (:617-623,652-658)."""
    Fz = np.eye(4)
    Fz[0, 0] = Fz[1, 1] = -1.0
    target_c2w = np.asarray(target_c2w, dtype=np.float64)
    out = np.empty((target_c2w.shape[0], 6, 3, 4), dtype=np.float64)
    for v in range(target_c2w.shape[0]):
        for fi, name in enumerate(FACE_ORDER):
            cam = target_c2w[v] @ CUBEMAP_TRANSFORMS[name]
            if name in ("top", "bottom"):
                cam = cam @ Fz
            out[v, fi] = np.linalg.inv(cam)[:3, :4]
    return out.astype(np.float32)


def front_w2c_matrices(target_c2w: np.ndarray) -> np.ndarray:
    """[V,4,4] target camera-to-world -> [V,3,4] float32 cam-from-world of the front cube face; the other five
    faces are exact signed axis permutations of it (cube formulation of the splat, csrc/reproj.cu)."""
    target_c2w = np.asarray(target_c2w, dtype=np.float64)
    return np.stack([np.linalg.inv(c)[:3, :4] for c in target_c2w]).astype(np.float32)


def rotation_from_vectors(u, v):
    """3x3 rotation taking u onto v (Rodrigues; identity / 180 deg special cases)."""
    norm_u, norm_v = np.linalg.norm(u), np.linalg.norm(v)
    if norm_u < 1e-15 or norm_v < 1e-15:
        return np.eye(3)
    u_hat, v_hat = u / norm_u, v / norm_v
    dot = np.clip(np.dot(u_hat, v_hat), -1.0, 1.0)
    if np.isclose(dot, 1.0):
        return np.eye(3)
    if np.isclose(dot, -1.0):
        temp = np.array([1.0, 0.0, 0.0])
        if np.abs(np.dot(u_hat, temp)) > 0.9:
            temp = np.array([0.0, 1.0, 0.0])
        w = np.cross(u_hat, temp)
        w /= np.linalg.norm(w)
        return np.eye(3) - 2.0 * np.outer(w, w)
    angle = np.arccos(dot)
    w = np.cross(u_hat, v_hat)
    w_hat = w / np.linalg.norm(w)
    K = np.array([[0, -w_hat[2], w_hat[1]], [w_hat[2], 0, -w_hat[0]], [-w_hat[1], w_hat[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1.0 - np.cos(angle)) * (K @ K)


def align_first_and_last_points(A, B):
    """(s, R, t) with B0 = s R A0 + t and B_last = s R A_last + t (first/last point only)."""
    A0, A1, B0, B1 = A[0], A[-1], B[0], B[-1]
    vA, vB = A1 - A0, B1 - B0
    lenA, lenB = np.linalg.norm(vA), np.linalg.norm(vB)
    if lenA < 1e-15:
        R = np.eye(3)
        return 1.0, R, B0 - R @ A0
    s = lenB / lenA
    R = rotation_from_vectors(vA, vB)
    return s, R, B0 - s * R @ A0


# ---------------------------------------------------------------------------------------------
# device-resident point cloud
# ---------------------------------------------------------------------------------------------


@dataclass
class PointScene:
    """What `build_open3d_scene` returns here: the packed point cloud in HBM.
    pts4: [cap,4] float32 {x,y,z, bits(r|g<<8|b<<16)};  count: device int64[1] (valid prefix length)."""
    pts4: torch.Tensor
    count: Optional[torch.Tensor] = None

    @property
    def device(self):
        return self.pts4.device

    def num_points(self) -> int:
        return int(self.count.item()) if self.count is not None else self.pts4.shape[0]


def pack_points_device(xyz: torch.Tensor, rgb_u8: Optional[torch.Tensor] = None,
                       images_nchw: Optional[torch.Tensor] = None) -> torch.Tensor:
    """xyz [N,3] f64|f32 (CUDA) + colours (uint8 [N,3] or float NCHW images in [0,1]) -> pts4 [N,4] f32."""
    _lib.require_cuda(xyz, "xyz")
    dev = xyz.device
    xyz = xyz.reshape(-1, 3).contiguous()
    N = xyz.shape[0]
    S = HW = 0
    if images_nchw is not None:
        images_nchw = images_nchw.to(dev, torch.float32).contiguous()
        S, HW = images_nchw.shape[0], images_nchw.shape[2] * images_nchw.shape[3]
    else:
        rgb_u8 = rgb_u8.to(dev).reshape(-1, 3).contiguous()
    out = torch.empty((N, 4), dtype=torch.float32, device=dev)
    x64 = _lib.ptr(xyz) if xyz.dtype == torch.float64 else None
    x32 = _lib.ptr(xyz) if xyz.dtype == torch.float32 else None
    if x64 is None and x32 is None:
        raise ValueError("xyz must be float64 or float32")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().evw_pack_points(x64, x32, _lib.ptr(rgb_u8) if images_nchw is None else None,
                                              _lib.ptr(images_nchw) if images_nchw is not None else None, S, HW,
                                              _lib.ptr(out), N, _lib.stream_ptr(dev)), "evw_pack_points")
    return out


def conf_select_device(conf: torch.Tensor, pts4: Optional[torch.Tensor], conf_thres: float, want_index: bool = False):
    """Percentile threshold + order-preserving compaction on the device.
    Returns (pts4_out | None, keep_idx | None, count int64[1], thr f32[1]); nothing is synchronised."""
    _lib.require_cuda(conf, "conf")
    dev = conf.device
    conf = conf.reshape(-1).float().contiguous()
    n = conf.numel()
    use_thr = 0 if conf_thres == 0.0 else 1
    k_lo, k_hi, gamma = percentile_rank_params(n, conf_thres, np.float32) if use_thr else (0, 0, 0.0)
    L = _lib.lib()
    ws_bytes = L.evw_conf_select_workspace(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out = torch.empty_like(pts4) if pts4 is not None else None
    keep = torch.empty(n, dtype=torch.int64, device=dev) if want_index else None
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    thr = torch.zeros(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.evw_conf_select(_lib.ptr(conf), _lib.ptr(pts4), n, k_lo, k_hi, gamma, use_thr, _lib.ptr(out),
                                     _lib.ptr(keep), _lib.ptr(count), _lib.ptr(thr), _lib.ptr(ws), ws_bytes,
                                     _lib.stream_ptr(dev)), "evw_conf_select")
    return out, keep, count, thr


def splat_workspace_bytes(views_per_pass: int = DEFAULT_VIEWS_PER_PASS, face_res: int = FACE_RES, flags: Optional[int] = None) -> int:
    return int(_lib.lib().evw_splat_workspace_flags(views_per_pass, face_res, splat_flags() if flags is None else flags))


def splat_to_panoramas_device(scene: PointScene, w2c: torch.Tensor, width: int = PANO_W, height: int = PANO_H,
                              face_res: int = FACE_RES, views_per_pass: int = DEFAULT_VIEWS_PER_PASS,
                              z_near: float = Z_NEAR, out: Optional[torch.Tensor] = None,
                              zbuf: Optional[torch.Tensor] = None, pretest: Optional[bool] = None,
                              overlap: Optional[bool] = None,
                              by_role: Optional[bool] = None, color_keys: Optional[bool] = None) -> torch.Tensor:
    """Fused splat + resolve -> uint8 [V,height,width,3] (CUDA).
    w2c [V,3,4]: cube formulation (front-face camera, one transform per point-view; the fast path);
    w2c [V,6,3,4]: six independent per-face cameras (the literal restatement of render_cubemap)."""
    dev = scene.device
    V = w2c.shape[0]
    flags = splat_flags(pretest, overlap, by_role, color_keys)
    w2c = w2c.to(dev, torch.float32).contiguous()
    lut = cube_lut_device(width, height, face_res, dev)
    L = _lib.lib()
    ws_bytes = L.evw_splat_workspace_flags(views_per_pass, face_res, flags if w2c.dim() == 3 else 0)
    if zbuf is None or zbuf.numel() * zbuf.element_size() < ws_bytes:
        zbuf = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((V, height, width, 3), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        if w2c.dim() == 3:
            _lib.check(L.evw_splat_cube_equirect(
                _lib.ptr(scene.pts4), scene.pts4.shape[0], _lib.ptr(scene.count), _lib.ptr(w2c), V, face_res,
                face_res / 2.0, z_near, _lib.ptr(lut), height, width, _lib.ptr(out), _lib.ptr(zbuf),
                zbuf.numel() * zbuf.element_size(), views_per_pass, flags, _lib.stream_ptr(dev)),
                "evw_splat_cube_equirect")
        else:
            _lib.check(L.evw_splat_cubemap_equirect(
                _lib.ptr(scene.pts4), scene.pts4.shape[0], _lib.ptr(scene.count), _lib.ptr(w2c), V, face_res,
                face_res / 2.0, z_near, _lib.ptr(lut), height, width, _lib.ptr(out), _lib.ptr(zbuf),
                zbuf.numel() * zbuf.element_size(), views_per_pass, _lib.stream_ptr(dev)), "evw_splat_cubemap_equirect")
    return out


# ---------------------------------------------------------------------------------------------
# reference-facing classes
# ---------------------------------------------------------------------------------------------


class PointCloudProcessor:
    """Handles point cloud processing and filtering (confidence percentile + compaction on the GPU)."""

    def __init__(self, device="cuda"):
        self.logger = logging.getLogger(self.__class__.__name__)
        self.device = torch.device(device)

    def filter_predictions(self, predictions: Dict, conf_thres: float = 50.0, filter_by_frames: str = "all",
                           mask_black_bg: bool = False, mask_white_bg: bool = False, mask_sky: bool = False,
                           target_dir: Optional[str] = None, image_subdir: Optional[str] = None,
                           prediction_mode: str = "Predicted Pointmap",
                           only_render_last_24_frame: bool = False) -> Tuple[np.ndarray, np.ndarray, float]:
        """Returns (vertices_3d float64 [N,3], colors_rgb uint8 [N,3], scene_scale) like the reference."""
        scene, keep = self.filter_predictions_device(predictions, conf_thres, filter_by_frames, mask_black_bg,
                                                     mask_white_bg, mask_sky, target_dir, image_subdir,
                                                     prediction_mode, only_render_last_24_frame, want_index=True)
        n = scene.num_points()
        points, _ = self._extract_point_data(predictions, prediction_mode)
        if n == 0:
            vertices, colors = np.array([[1, 0, 0]]), np.array([[255, 255, 255]])
        else:
            idx = keep[:n].cpu().numpy()
            vertices = np.asarray(points).reshape(-1, 3)[idx]
            bits = scene.pts4[:n, 3].contiguous().view(torch.int32).cpu().numpy()
            colors = np.stack([bits & 0xFF, (bits >> 8) & 0xFF, (bits >> 16) & 0xFF], axis=1).astype(np.uint8)
        return vertices, colors, self._calculate_scene_scale(vertices)

    def filter_predictions_device(self, predictions: Dict, conf_thres: float = 50.0, filter_by_frames: str = "all",
                                  mask_black_bg: bool = False, mask_white_bg: bool = False, mask_sky: bool = False,
                                  target_dir=None, image_subdir=None, prediction_mode: str = "Predicted Pointmap",
                                  only_render_last_24_frame: bool = False, want_index: bool = False):
        """Device-resident variant: returns (PointScene, keep_idx | None) without any host sync."""
        if mask_sky and target_dir and image_subdir:
            raise NotImplementedError("sky segmentation (ONNX) is outside the hot path (SURVEY §2.1 #3)")
        if mask_black_bg or mask_white_bg:
            raise NotImplementedError("background colour masks are outside the hot path (SURVEY §2.1 #3)")
        points, conf = self._extract_point_data(predictions, prediction_mode)
        images = predictions["images"]
        if filter_by_frames not in ("all", "All"):
            sel = self._parse_frame_filter(filter_by_frames)
            if sel is not None:
                # the reference slices the predictions dict IN PLACE here (:199-203): later stages — align_extrinsics in
                # predictions_to_target_view — then see only the selected frame's extrinsic
                points, conf = points[sel:sel + 1], conf[sel:sel + 1]
                predictions["images"] = images = images[sel:sel + 1]
                if "extrinsic" in predictions:
                    predictions["extrinsic"] = predictions["extrinsic"][sel:sel + 1]

        def dev(x, dtype=None):
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            return t.to(self.device, dtype=dtype, non_blocking=True)

        pts = dev(points)
        if pts.dtype not in (torch.float32, torch.float64):
            pts = pts.double()
        img = dev(images, torch.float32)
        if img.dim() == 4 and img.shape[1] == 3:
            pts4 = pack_points_device(pts, images_nchw=img)
        else:  # NHWC colours: the reference reshapes them directly (:290)
            rgb = (img.reshape(-1, 3) * 255).to(torch.uint8)
            pts4 = pack_points_device(pts, rgb_u8=rgb)
        out, keep, count, _ = conf_select_device(dev(conf, torch.float32), pts4, conf_thres, want_index)
        return PointScene(out, count), keep

    def _extract_point_data(self, predictions: Dict, prediction_mode: str):
        def conf_or_ones(key, pts):  # the default is only materialised when the key is missing
            if key in predictions:
                return predictions[key]
            shape = tuple(pts.shape[:-1])
            return torch.ones(shape, dtype=torch.float32) if isinstance(pts, torch.Tensor) else np.ones(shape, dtype=np.float32)

        if "Pointmap" in prediction_mode and "world_points" in predictions:
            pts = predictions["world_points"]
            conf = conf_or_ones("world_points_conf", pts)
        else:
            if "Pointmap" in prediction_mode:
                self.logger.info("Warning: world_points not found, falling back to depth-based points")
            pts = predictions["world_points_from_depth"]
            conf = conf_or_ones("depth_conf", pts)
        return pts, conf

    def _parse_frame_filter(self, filter_by_frames: str) -> Optional[int]:
        try:
            return int(filter_by_frames.split(":")[0])
        except (ValueError, IndexError):
            return None

    def _calculate_scene_scale(self, vertices: np.ndarray) -> float:
        """5/95-percentile extent; the Open3D branch of the reference never uses it (:330-337)."""
        if vertices is None or len(vertices) == 0:
            return 1.0
        lo = np.percentile(vertices, 5, axis=0)
        hi = np.percentile(vertices, 95, axis=0)
        return float(np.linalg.norm(hi - lo))


class SceneBuilder:
    """Handles 3D scene construction and camera integration."""

    def __init__(self, device="cuda"):
        self.logger = logging.getLogger(self.__class__.__name__)
        self.device = torch.device(device)

    def build_open3d_scene(self, vertices, colors=None) -> PointScene:
        """vertices [N,3] + colors uint8 [N,3] (numpy or torch) -> PointScene; a PointScene passes through."""
        if isinstance(vertices, PointScene):
            return vertices
        v = vertices if isinstance(vertices, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(vertices))
        c = colors if isinstance(colors, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(colors))
        v = v.to(self.device)
        if v.dtype not in (torch.float32, torch.float64):
            v = v.double()
        return PointScene(pack_points_device(v, rgb_u8=c.to(self.device).to(torch.uint8)))

    def align_extrinsics(self, camera_pose, predictions_extrinsic, num_target_view, outdir, only_render_last_24_frame):
        """Similarity-align the GT trajectory to the VGGT frame from the first/last camera centre and
        return the next `num_target_view` target camera-to-world matrices (float64 [V,4,4])."""
        predictions_extrinsic = np.asarray(
            predictions_extrinsic.detach().cpu().numpy() if isinstance(predictions_extrinsic, torch.Tensor)
            else predictions_extrinsic)
        camera_pose = np.asarray(camera_pose.detach().cpu().numpy() if isinstance(camera_pose, torch.Tensor) else camera_pose)
        num_cameras = len(predictions_extrinsic)
        E = np.zeros((num_cameras, 4, 4))
        E[:, :3, :4] = predictions_extrinsic
        E[:, 3, 3] = 1
        E_inv = np.stack([np.linalg.inv(e) for e in E])
        try:
            segment_id = int(outdir.rstrip("/").split("_")[-1])
        except Exception:
            segment_id = 1
        if not only_render_last_24_frame:
            start = (segment_id + 1) * num_target_view + 1
            target_gt = camera_pose[start:start + num_target_view]
        else:
            start = -num_target_view
            target_gt = camera_pose[start:]
        gt = camera_pose[:start]
        s, R, t = align_first_and_last_points(gt[:, :3, 3], E_inv[:, :3, 3])
        transform = np.eye(4)
        transform[:3, :3] = s * R
        transform[:3, 3] = t
        return np.einsum("ij, bjk -> bik", transform, target_gt)

    def remove_opend3d_scene(self, renderer):
        """Release the point buffers (the reference clears the Filament scene here)."""
        if isinstance(renderer, PointScene):
            renderer.pts4 = renderer.pts4[:0]
            renderer.count = None


class CubemapRenderer:
    """Handles cubemap rendering and equirectangular conversion."""

    def __init__(self, views_per_pass: int = DEFAULT_VIEWS_PER_PASS, pinned_output: bool = False):
        """pinned_output: return the panoramas as a numpy view of a page-locked buffer owned by this renderer (one
        asynchronous device-to-host copy at PCIe rate; the array is overwritten by the next call).  The default
        returns a fresh array like the reference does."""
        self.logger = logging.getLogger(self.__class__.__name__)
        self.start = True
        self.views_per_pass = views_per_pass
        self.pinned_output = pinned_output
        self._pinned: Optional[torch.Tensor] = None

    def cube_to_equirectangular_cuda(self, cube_faces_batch, width, height, device="cuda"):
        """dict face -> uint8 [B,3,res,res]  ->  numpy uint8 [B,height,width,3]."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("evoworld_b200 has no CPU path for cube_to_equirectangular")
        faces = torch.stack([torch.as_tensor(cube_faces_batch[n]).to(dev) for n in FACE_ORDER], dim=1).contiguous()
        if faces.dtype != torch.uint8:
            raise ValueError("cube faces must be uint8")
        B, _, _, res, _ = faces.shape
        lut = cube_lut_device(width, height, res, dev)
        out = torch.empty((B, height, width, 3), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().evw_cube_to_equirect_u8(_lib.ptr(faces), _lib.ptr(lut), B, res, height, width,
                                                          _lib.ptr(out), _lib.stream_ptr(dev)), "evw_cube_to_equirect_u8")
        return out.cpu().numpy()

    def _faces(self, scene: PointScene, cams_c2w: np.ndarray, res: int) -> torch.Tensor:
        w2c = torch.from_numpy(face_w2c_matrices(cams_c2w)).to(scene.device)
        V = w2c.shape[0]
        L = _lib.lib()
        ws_bytes = L.evw_splat_workspace(1, res)
        zbuf = torch.empty(ws_bytes, dtype=torch.uint8, device=scene.device)
        out = torch.empty((V, 6, res, res, 3), dtype=torch.uint8, device=scene.device)
        with torch.cuda.device(scene.device):
            _lib.check(L.evw_splat_faces_u8(_lib.ptr(scene.pts4), scene.num_points(), _lib.ptr(w2c), V, res, res / 2.0,
                                            Z_NEAR, _lib.ptr(out), _lib.ptr(zbuf), ws_bytes,
                                            _lib.stream_ptr(scene.device)), "evw_splat_faces_u8")
        return out

    def render_face(self, scene_3d: PointScene, cam, fov=90, res=(512, 512), outdir="demo_trimesh_render.png",
                    do_flip=False, savefig=False):
        """One pinhole point render (uint8 [res,res,3]); `cam` is the face camera-to-world."""
        if fov != 90 or res[0] != res[1]:
            raise NotImplementedError("render_face: only the 90 degree square cube-face camera is built")
        cam = np.asarray(cam, dtype=np.float64)
        if do_flip:
            Fz = np.eye(4)
            Fz[0, 0] = Fz[1, 1] = -1.0
            cam = cam @ Fz
        w2c = np.repeat(np.linalg.inv(cam)[None, None, :3, :4], 6, axis=1).astype(np.float32)
        w2c_t = torch.from_numpy(w2c).to(scene_3d.device)
        L = _lib.lib()
        ws_bytes = L.evw_splat_workspace(1, res[0])
        zbuf = torch.empty(ws_bytes, dtype=torch.uint8, device=scene_3d.device)
        out = torch.empty((1, 6, res[0], res[0], 3), dtype=torch.uint8, device=scene_3d.device)
        with torch.cuda.device(scene_3d.device):
            _lib.check(L.evw_splat_faces_u8(_lib.ptr(scene_3d.pts4), scene_3d.num_points(), _lib.ptr(w2c_t), 1, res[0],
                                            res[0] / 2.0, Z_NEAR, _lib.ptr(out), _lib.ptr(zbuf), ws_bytes,
                                            _lib.stream_ptr(scene_3d.device)), "evw_splat_faces_u8")
        img = out[0, 0].cpu().numpy()
        if savefig:
            import cv2

            cv2.imwrite(outdir, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
        return img

    def render_cubemap(self, scene_3d: PointScene, cam, res=(512, 512), outdir="demo_trimesh_render.png", savefig=False,
                       initial_transformation=None):
        """dict face -> uint8 [res,res,3] for one target camera-to-world."""
        faces = self._faces(scene_3d, np.asarray(cam, dtype=np.float64)[None], res[0])[0].cpu().numpy()
        cubemap = {name: faces[i] for i, name in enumerate(FACE_ORDER)}
        if savefig:
            import cv2

            os.makedirs(outdir, exist_ok=True)
            for name, img in cubemap.items():
                cv2.imwrite(os.path.join(outdir, f"{name}.png"), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
        return cubemap

    def render_cubemaps_to_panoramas_device(self, scene_3d: PointScene, target_extrinsic: np.ndarray,
                                            width: int = PANO_W, height: int = PANO_H) -> torch.Tensor:
        w2c = torch.from_numpy(front_w2c_matrices(target_extrinsic)).to(scene_3d.device)
        return splat_to_panoramas_device(scene_3d, w2c, width, height, FACE_RES, self.views_per_pass)

    def render_cubemaps_to_panoramas(self, scene_3d: PointScene, target_extrinsic: np.ndarray, orig_extrinsic=None,
                                     num_target_view: int = 24, outdir: str = "demo_pyrender_render",
                                     only_render_last_24_frame: bool = False, write_png: bool = True):
        """Render every target view to a 2000x1000 panorama; returns numpy uint8 [V,1000,2000,3] and
        writes outdir/{idx:02}.png (RGB->BGR for cv2, as the reference :707-710)."""
        dev_panos = self.render_cubemaps_to_panoramas_device(scene_3d, target_extrinsic)
        if self.pinned_output:
            if self._pinned is None or self._pinned.shape != dev_panos.shape:
                self._pinned = torch.empty(dev_panos.shape, dtype=torch.uint8, pin_memory=True)
            self._pinned.copy_(dev_panos, non_blocking=True)
            torch.cuda.current_stream(dev_panos.device).synchronize()
            panos = self._pinned.numpy()
        else:
            panos = dev_panos.cpu().numpy()
        if write_png:
            import cv2

            os.makedirs(outdir, exist_ok=True)
            for idx, pano in enumerate(panos):
                cv2.imwrite(os.path.join(outdir, f"{idx:02}.png"), cv2.cvtColor(pano, cv2.COLOR_RGB2BGR))
        return panos


def predictions_to_target_view(predictions: Dict, camera_pose: np.ndarray, conf_thres: float = 50.0,
                               filter_by_frames: str = "all", point_processor=None, scene_builder=None,
                               cubemap_renderer=None, mask_black_bg: bool = False, mask_white_bg: bool = False,
                               show_cam: bool = True, mask_sky: bool = False, target_dir: Optional[str] = None,
                               image_subdir: Optional[str] = None, prediction_mode: str = "Predicted Pointmap",
                               num_target_view: int = 24, outdir: str = "demo_pyrender_render",
                               only_render_last_24_frame: bool = False):
    """VGGT predictions -> filtered point cloud -> `num_target_view` reprojected panoramas.

    Same arguments as the reference (:1216-1282); the processors default to fresh instances created
    at call time (the reference instantiates them — and downloads an ONNX model — at import time)."""
    if not isinstance(predictions, dict):
        raise ValueError("predictions must be a dictionary")
    point_processor = point_processor or PointCloudProcessor()
    scene_builder = scene_builder or SceneBuilder()
    cubemap_renderer = cubemap_renderer or CubemapRenderer()
    scene, _ = point_processor.filter_predictions_device(
        predictions, conf_thres, filter_by_frames, mask_black_bg, mask_white_bg, mask_sky, target_dir, image_subdir,
        prediction_mode, only_render_last_24_frame)
    if scene.num_points() == 0:  # reference fallback: a single white point at (1,0,0) (:306-308)
        one = torch.tensor([[1.0, 0.0, 0.0]], dtype=torch.float64, device=scene.device)
        scene = PointScene(pack_points_device(one, rgb_u8=torch.full((1, 3), 255, dtype=torch.uint8, device=scene.device)))
    scene_3d = scene_builder.build_open3d_scene(scene)
    target_extrinsic = scene_builder.align_extrinsics(camera_pose, predictions["extrinsic"], num_target_view, outdir,
                                                      only_render_last_24_frame)
    panoramas = cubemap_renderer.render_cubemaps_to_panoramas(scene_3d, target_extrinsic, predictions["extrinsic"],
                                                              num_target_view, outdir, only_render_last_24_frame)
    scene_builder.remove_opend3d_scene(scene_3d)
    return panoramas
