"""Host-side mirrors in evoworld_b200/ (pose math, lookup table, percentile rank parameters,
alignment) against the golden vectors and numpy.  CPU only; no compute call into the CUDA library."""
import hashlib

import numpy as np
import pytest
import torch

from evoworld_b200 import geometry as G
from evoworld_b200 import plucker as P
from evoworld_b200 import reprojection as R
from evoworld_b200.equi2pers import pix2dir_matrix


def test_equirectangular_to_ray_bit_exact(golden):
    np.testing.assert_array_equal(P.equirectangular_to_ray(12, 24), golden["ray_12x24"])
    r = P.equirectangular_to_ray(72, 128)
    assert r.shape == (72, 128, 3) and r.dtype == np.float32
    np.testing.assert_allclose(np.linalg.norm(r, axis=-1), 1.0, atol=1e-6)
    np.testing.assert_allclose(r[36, 64], [0, 0, 1], atol=1e-6)  # centre pixel -> +Z


def test_pose_matrices(golden):
    poses = torch.from_numpy(golden["poses_rdf"])
    traj = poses[101:115].clone()
    traj[:, :3] *= 0.1
    np.testing.assert_allclose(G.xyz_euler_to_three_by_four_matrix_batch(traj, relative=True).numpy(),
                               golden["c2w_3x4_rel"], atol=1e-6)
    np.testing.assert_allclose(G.xyz_euler_to_four_by_four_matrix_batch(poses[:30], relative=True).numpy(),
                               golden["c2w_4x4_rel"], atol=1e-5)
    np.testing.assert_allclose(G.xyz_euler_to_four_by_four_matrix_batch(poses[:30]).numpy(), golden["c2w_4x4_abs"], atol=1e-5)
    f = G.xyz_euler_to_four_by_four_matrix_batch(poses[:5], relative=True, flatten=True)
    assert f.shape == (5, 16)
    np.testing.assert_allclose(f[0].view(4, 4).numpy(), np.eye(4), atol=1e-6)


def test_pose_encoding(golden):
    extr, intr = G.pose_encoding_to_extri_intri(torch.from_numpy(golden["pose_enc"]), (14, 18))
    np.testing.assert_allclose(extr.numpy(), golden["pe_extr"], atol=1e-6)
    np.testing.assert_allclose(intr.numpy(), golden["pe_intr"], rtol=1e-6)
    inv = G.closed_form_inverse_se3(golden["pe_extr"][0])
    full = np.tile(np.eye(4), (3, 1, 1))
    full[:, :3] = golden["pe_extr"][0]
    np.testing.assert_allclose(inv @ full, np.tile(np.eye(4), (3, 1, 1)), atol=1e-5)


def test_cube_lut_matches_reference(golden):
    np.testing.assert_array_equal(R.build_cube_lut(64, 32, 16), golden["lut_64x32_r16"])
    full = R.build_cube_lut(2000, 1000, 512)
    np.testing.assert_array_equal(full[::37], golden["lut_full_rows"])
    assert hashlib.sha256(full.tobytes()).digest() == golden["lut_full_sha256"].tobytes()
    face = full >> 28
    assert set(np.unique(face)) == set(range(6))


@pytest.mark.parametrize("n", [1, 2, 3, 10, 1001, 4001, 65537, 203056])
@pytest.mark.parametrize("q", [0.5, 5.0, 30.0, 50.0, 77.7, 99.5, 100.0])
def test_percentile_rank_params_match_numpy(n, q):
    rng = np.random.default_rng(n)
    a = (1 + np.exp(rng.normal(size=n))).astype(np.float32)
    k_lo, k_hi, gamma = R.percentile_rank_params(n, q, np.float32)
    s = np.sort(a)
    lo, hi, t = s[k_lo], s[k_hi], np.float32(gamma)
    d = hi - lo
    thr = lo + d * t
    if t >= 0.5:
        thr = hi - d * (np.float32(1) - t)
    want = np.percentile(a, q)
    assert want.dtype == np.float32
    assert np.float32(thr) == want, (n, q, thr, want)


def test_alignment_and_face_matrices(golden):
    s, Rm, t = R.align_first_and_last_points(golden["align_A"], golden["align_B"])
    np.testing.assert_allclose(s, golden["align_s"], rtol=1e-14)
    np.testing.assert_allclose(Rm, golden["align_R"], atol=1e-14)
    np.testing.assert_allclose(t, golden["align_t"], atol=1e-12)
    np.testing.assert_array_equal(np.stack([R.CUBEMAP_TRANSFORMS[k] for k in R.FACE_ORDER]), golden["cubemap_transforms"])
    w2c = R.face_w2c_matrices(np.eye(4)[None])
    assert w2c.shape == (1, 6, 3, 4) and w2c.dtype == np.float32
    # each face camera looks along the expected world axis: third row of w2c is the view direction
    np.testing.assert_allclose(w2c[0, :, 2, :3], [[0, 0, 1], [1, 0, 0], [0, 0, -1], [-1, 0, 0], [0, -1, 0], [0, 1, 0]], atol=1e-7)


def test_align_extrinsics_matches_oracle(golden):
    from oracle import reproj_np as O

    poses = torch.from_numpy(golden["poses_rdf"]).double()
    cam = G.xyz_euler_to_four_by_four_matrix_batch(poses, relative=True).numpy()
    rng = np.random.default_rng(3)
    # VGGT-frame cameras: a similarity transform of the first 49 GT cameras, as w2c 3x4
    c2w = cam[:49].copy()
    c2w[:, :3, 3] = c2w[:, :3, 3] * 0.37 + rng.normal(size=3)
    extr = np.linalg.inv(c2w)[:, :3, :4]
    got = R.SceneBuilder().align_extrinsics(cam, extr, 24, "out/rendered_1", False)
    want = O.align_extrinsics(cam, extr, 24, "out/rendered_1", False)
    assert got.shape == (24, 4, 4)
    np.testing.assert_allclose(got, want, atol=1e-12)


def test_pix2dir_matrix():
    from oracle import reproj_np as O

    for yaw in (0.0, 0.3, -2.0):
        np.testing.assert_allclose(pix2dir_matrix(yaw, 0.0, 0.0, 384, 512, 90.0), O.equi2pers_matrix(yaw, 0, 0, 384, 512, 90.0), atol=1e-15)
    M = pix2dir_matrix(0.0, 0.0, 0.0, 384, 512, 90.0)
    np.testing.assert_allclose(M @ np.array([256.0, 192.0, 1.0]), [1, 0, 0], atol=1e-12)  # centre pixel -> forward


def test_no_cpu_fallback():
    ray = torch.zeros(4, 8, 3)
    with pytest.raises(RuntimeError):
        P.ray_c2w_to_plucker(ray, torch.zeros(2, 3, 4))
    with pytest.raises(RuntimeError):
        R.CubemapRenderer().cube_to_equirectangular_cuda({}, 64, 32, device="cpu")


def test_splat_flags_and_workspace(built_lib):
    """Flag word of evw_splat_cube_equirect (include/evoworld_b200.h) and the workspace it implies."""
    assert R.splat_flags(False, False, False, False) == 0
    assert R.splat_flags(False, False, False, True) == R.SPLAT_COLOR_KEYS == 16
    assert R.splat_flags(True, True, True, False) == R.SPLAT_PRETEST | R.SPLAT_OVERLAP | R.SPLAT_OVERLAP_BY_ROLE
    one = 4 * 6 * 512 * 512 * 8
    assert R.splat_workspace_bytes(4, 512, 0) == one
    assert R.splat_workspace_bytes(4, 512, R.SPLAT_OVERLAP) == 2 * one  # two passes in flight
    assert R.splat_workspace_bytes(4, 512, R.SPLAT_PRETEST) == one


def test_missing_confidence_defaults_to_ones():
    """reproject_vggt_open3d_utils.py:224-247: a prediction dict without a confidence map filters with conf == 1."""
    pp = R.PointCloudProcessor.__new__(R.PointCloudProcessor)
    import logging

    pp.logger = logging.getLogger("t")
    pts = np.zeros((2, 3, 4, 3), dtype=np.float64)
    got_pts, conf = pp._extract_point_data({"world_points_from_depth": pts}, "Depthmap and Camera Branch")
    assert got_pts is pts and conf.shape == (2, 3, 4) and conf.dtype == np.float32 and (conf == 1).all()
    tp = torch.zeros(2, 3, 4, 3, dtype=torch.float64)
    _, conf_t = pp._extract_point_data({"world_points": tp}, "Predicted Pointmap")
    assert isinstance(conf_t, torch.Tensor) and tuple(conf_t.shape) == (2, 3, 4) and bool((conf_t == 1).all())
    c = np.full((2, 3, 4), 2.0, dtype=np.float32)
    _, conf2 = pp._extract_point_data({"world_points_from_depth": pts, "depth_conf": c}, "x")
    assert conf2 is c


def test_pix2dir_matrices_batched_equals_scalar():
    """The batched matrix builder used by Equi2Pers.__call__ is the scalar restatement, element for element."""
    from evoworld_b200.equi2pers import pix2dir_matrices

    rng = np.random.default_rng(0)
    rots = [{"yaw": float(rng.uniform(-4, 4)), "pitch": float(rng.uniform(-1, 1)), "roll": float(rng.uniform(-1, 1))}
            for _ in range(200)] + [{"yaw": 0.3}, {}]
    got = pix2dir_matrices(rots, 384, 512, 90.0)
    want = np.stack([pix2dir_matrix(r.get("yaw", 0.0), r.get("pitch", 0.0), r.get("roll", 0.0), 384, 512, 90.0) for r in rots])
    np.testing.assert_array_equal(got, want)
    got_zd = pix2dir_matrices(rots[:5], 96, 128, 70.0, skew=0.1, z_down=True)
    want_zd = np.stack([pix2dir_matrix(r.get("yaw", 0.0), r.get("pitch", 0.0), r.get("roll", 0.0), 96, 128, 70.0, 0.1, True)
                        for r in rots[:5]])
    np.testing.assert_array_equal(got_zd, want_zd)
