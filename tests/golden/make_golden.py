"""Generate tests/golden/reproj_golden.npz from the REFERENCE's own functions.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Imports the reference modules that import cleanly (SURVEY §8c list A) and, with stub modules for
matplotlib/onnxruntime/open3d/requests/trimesh, reproject_vggt_open3d_utils (list B).  The outputs
are the golden vectors that pin oracle/reproj_np.py and the host-side mirrors in evoworld_b200/.
"""
import hashlib
import os
import sys
import tempfile
from unittest.mock import MagicMock

import numpy as np
import torch

REF = "/root/reference"
sys.path[:0] = [REF, os.path.join(REF, "third_party/vggt")]
import importlib.machinery

for m in ("matplotlib", "matplotlib.pyplot", "onnxruntime", "open3d", "requests", "trimesh"):
    if m not in sys.modules:
        stub = MagicMock()
        stub.__spec__ = importlib.machinery.ModuleSpec(m, None)
        sys.modules[m] = stub

out = {}
rng = np.random.default_rng(1234)

# ---- camera trajectory of the only fixture the reference ships
lines = open(os.path.join(REF, "example/case_000/camera_poses.txt")).read().strip().splitlines()[1:]
poses = np.array([[float(v) for v in l.split(",")[1:]] for l in lines]) * np.array([1, -1, 1, -1, 1, -1])
out["poses_rdf"] = poses.astype(np.float32)

from utils.plucker_embedding import equirectangular_to_ray, ray_c2w_to_plucker
from utils.geometry import xyz_euler_to_four_by_four_matrix_batch
from dataset.CameraTrajDataset import xyz_euler_to_three_by_four_matrix_batch
from evoworld.reprojection.pano_to_pers_utils import calculate_segment_indices

ray = equirectangular_to_ray(12, 24)
out["ray_12x24"] = ray
traj = torch.tensor(poses[101:115], dtype=torch.float32)
traj[:, :3] *= 0.1
c2w = xyz_euler_to_three_by_four_matrix_batch(traj, relative=True)
out["c2w_3x4_rel"] = c2w.numpy()
out["c2w_4x4_rel"] = xyz_euler_to_four_by_four_matrix_batch(torch.tensor(poses[:30], dtype=torch.float32), relative=True).numpy()
out["c2w_4x4_abs"] = xyz_euler_to_four_by_four_matrix_batch(torch.tensor(poses[:30], dtype=torch.float32), relative=False).numpy()
out["plucker"] = ray_c2w_to_plucker(torch.from_numpy(ray), c2w).contiguous().numpy()
out["segment_indices"] = np.array([calculate_segment_indices(s) for s in range(4)])

# ---- VGGT utils
from vggt.utils.geometry import unproject_depth_map_to_point_map
from vggt.utils.pose_enc import pose_encoding_to_extri_intri

pose_enc = torch.tensor(rng.normal(size=(1, 3, 9)), dtype=torch.float32)
pose_enc[..., 7:] = pose_enc[..., 7:].abs() * 0.2 + 0.8
extr, intr = pose_encoding_to_extri_intri(pose_enc, (14, 18))
out["pose_enc"], out["pe_extr"], out["pe_intr"] = pose_enc.numpy(), extr.numpy(), intr.numpy()
depth = np.exp(rng.normal(0.7, 0.5, size=(3, 14, 18, 1))).astype(np.float32)
out["lift_depth"] = depth
out["lift_points"] = unproject_depth_map_to_point_map(depth, extr[0].numpy(), intr[0].numpy())

# ---- reprojection utils (stubbed import; an empty skyseg.onnx stops the import-time download)
cwd = os.getcwd()
tmp = tempfile.mkdtemp()
open(os.path.join(tmp, "skyseg.onnx"), "wb").close()
os.chdir(tmp)
import evoworld.reprojection.reproject_vggt_open3d_utils as ru
os.chdir(cwd)

pp = ru.PointCloudProcessor.__new__(ru.PointCloudProcessor)
n = 4001
conf = (1 + np.exp(rng.normal(size=n))).astype(np.float32)
conf[rng.integers(0, n, 400)] = conf[rng.integers(0, n, 400)]  # ties
pts = rng.normal(size=(n, 3))
imgs = rng.random((1, 3, 1, n)).astype(np.float32)
cols = pp._extract_colors(imgs)
out["cf_conf"], out["cf_pts"], out["cf_imgs"], out["cf_cols"] = conf, pts, imgs, cols
for q in (50.0, 30.0, 0.0, 99.5):
    v, c = pp._apply_confidence_filter(pts, conf, cols, q)
    out[f"cf_v_{q}"], out[f"cf_c_{q}"] = v, c

A = poses[:49, :3] * 1.0
B = (A @ ru.rotation_from_vectors(np.array([1.0, 0.2, 0.1]), np.array([0.3, 1.0, -0.5])).T) * 1.7 + np.array([0.5, -1.0, 2.0])
s, R, t = ru.align_first_and_last_points(A, B)
out["align_A"], out["align_B"], out["align_s"], out["align_R"], out["align_t"] = A, B, np.array(s), R, t

cr = ru.CubemapRenderer.__new__(ru.CubemapRenderer)


def ref_lut(width, height, res):
    """Index map of the reference's cube_to_equirectangular_cuda(device='cpu'), obtained by feeding it
    faces whose RGB encodes (face, row, col)."""
    rows, cols_ = np.meshgrid(np.arange(res), np.arange(res), indexing="ij")
    faces = {}
    for fi, name in enumerate(ru.CUBEMAP_TRANSFORMS.keys()):
        code = (fi << 20) | (rows << 10) | cols_
        img = np.stack([code & 0xFF, (code >> 8) & 0xFF, (code >> 16) & 0xFF], 0).astype(np.uint8)
        faces[name] = torch.from_numpy(img[None])
    pano = cr.cube_to_equirectangular_cuda(faces, width, height, device="cpu")[0].astype(np.uint32)
    code = pano[..., 0] | (pano[..., 1] << 8) | (pano[..., 2] << 16)
    return ((code >> 20) << 28) | (((code >> 10) & 0x3FF) << 14) | (code & 0x3FF)


out["lut_64x32_r16"] = ref_lut(64, 32, 16).astype(np.uint32)
full = ref_lut(2000, 1000, 512).astype(np.uint32)
out["lut_full_sha256"] = np.frombuffer(hashlib.sha256(full.tobytes()).digest(), dtype=np.uint8)
out["lut_full_rows"] = full[::37]  # every 37th row, for diagnostics when the hash differs
faces = {k: torch.from_numpy(rng.integers(0, 256, size=(2, 3, 16, 16), dtype=np.uint8)) for k in ru.CUBEMAP_TRANSFORMS}
out["c2e_faces"] = np.stack([faces[k].numpy() for k in ru.CUBEMAP_TRANSFORMS], 1)
out["c2e_pano"] = cr.cube_to_equirectangular_cuda(faces, 64, 32, device="cpu")
out["cubemap_transforms"] = np.stack([ru.CUBEMAP_TRANSFORMS[k] for k in ru.CUBEMAP_TRANSFORMS])

dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reproj_golden.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes;", len(out), "arrays")
