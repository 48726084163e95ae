"""The oracle (oracle/reproj_np.py, oracle/reproj_oracle.c) against the golden vectors produced by
the reference's own functions (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import torch

from oracle import reproj_np as O


def test_ray_and_plucker(golden):
    ray = O.equirectangular_to_ray(12, 24)
    assert ray.dtype == golden["ray_12x24"].dtype
    np.testing.assert_array_equal(ray, golden["ray_12x24"])
    pl = O.ray_c2w_to_plucker(torch.from_numpy(ray), torch.from_numpy(golden["c2w_3x4_rel"]))
    np.testing.assert_allclose(pl.numpy(), golden["plucker"], rtol=0, atol=1e-6)


def test_pose_matrices(golden):
    poses = torch.from_numpy(golden["poses_rdf"])
    traj = poses[101:115].clone()
    traj[:, :3] *= 0.1
    np.testing.assert_allclose(O.euler_to_matrix(traj, True, False).numpy(), golden["c2w_3x4_rel"], atol=2e-6)
    np.testing.assert_allclose(O.euler_to_matrix(poses[:30], True, True).numpy(), golden["c2w_4x4_rel"], atol=2e-5)
    np.testing.assert_allclose(O.euler_to_matrix(poses[:30], False, True).numpy(), golden["c2w_4x4_abs"], atol=2e-5)


def test_pose_encoding(golden):
    extr, intr = O.pose_encoding_to_extri_intri(torch.from_numpy(golden["pose_enc"]), (14, 18))
    np.testing.assert_allclose(extr.numpy(), golden["pe_extr"], atol=1e-6)
    np.testing.assert_allclose(intr.numpy(), golden["pe_intr"], rtol=1e-6)


def test_lift(golden):
    pts = O.unproject_depth_map_to_point_map(golden["lift_depth"], golden["pe_extr"][0], golden["pe_intr"][0])
    assert pts.dtype == np.float64
    np.testing.assert_allclose(pts, golden["lift_points"], rtol=0, atol=1e-12)


def test_confidence_filter(golden):
    cols = O.extract_colors(golden["cf_imgs"])
    np.testing.assert_array_equal(cols, golden["cf_cols"])
    for q in (50.0, 30.0, 0.0, 99.5):
        v, c = O.apply_confidence_filter(golden["cf_pts"], golden["cf_conf"], cols, q)
        np.testing.assert_array_equal(v, golden[f"cf_v_{q}"])
        np.testing.assert_array_equal(c, golden[f"cf_c_{q}"])


def test_alignment(golden):
    s, R, t = O.align_first_and_last_points(golden["align_A"], golden["align_B"])
    np.testing.assert_allclose(s, golden["align_s"], rtol=1e-14)
    np.testing.assert_allclose(R, golden["align_R"], atol=1e-14)
    np.testing.assert_allclose(t, golden["align_t"], atol=1e-12)
    np.testing.assert_array_equal(np.stack([O.CUBEMAP_TRANSFORMS[k] for k in O.FACE_ORDER]), golden["cubemap_transforms"])


def test_cube_to_equirect_lut(golden):
    lut = O.cube_to_equirect_lut(64, 32, 16)
    np.testing.assert_array_equal(lut, golden["lut_64x32_r16"])
    pano = O.cube_to_equirectangular({n: golden["c2e_faces"][:, i] for i, n in enumerate(O.FACE_ORDER)}, 64, 32)
    np.testing.assert_array_equal(pano, golden["c2e_pano"])
    full = O.cube_to_equirect_lut(2000, 1000, 512)
    np.testing.assert_array_equal(full[::37], golden["lut_full_rows"])
    assert hashlib.sha256(full.tobytes()).digest() == golden["lut_full_sha256"].tobytes()


def test_equi2pers_invariants():
    """pyequilib is not vendored (parity unpinned): assert the geometric invariants instead."""
    rng = np.random.default_rng(0)
    He, We = 64, 128
    equi = rng.integers(0, 256, size=(3, He, We), dtype=np.uint8)
    # smooth image so that bilinear sampling is meaningful
    yy, xx = np.meshgrid(np.arange(He), np.arange(We), indexing="ij")
    equi[0] = (xx * 255 // (We - 1)).astype(np.uint8)
    equi[1] = (yy * 255 // (He - 1)).astype(np.uint8)
    out = O.equi2pers(equi, yaw=0.0, Hp=48, Wp=64)
    assert out.shape == (3, 48, 64) and out.dtype == np.uint8
    # yaw = 0: perspective centre looks at the panorama centre (He/2, We/2)
    assert abs(int(out[0, 24, 32]) - int(equi[0, He // 2, We // 2])) <= 2
    assert abs(int(out[1, 24, 32]) - int(equi[1, He // 2, We // 2])) <= 2
    # a quarter turn moves the centre by We/4 columns (sign: z_down=False negates yaw)
    q = O.equi2pers(equi, yaw=np.pi / 2, Hp=48, Wp=64)
    col = int(q[0, 24, 32]) * (We - 1) / 255
    assert min(abs(col - (We / 2 - We / 4)), abs(col - (We / 2 + We / 4))) <= 1.5
    # horizontal shift of the panorama by k columns == yaw of 2 pi k / We
    k = 16
    rolled = np.roll(equi, k, axis=2)
    a = O.equi2pers(rolled, yaw=0.0, Hp=48, Wp=64)[1]
    b = O.equi2pers(equi, yaw=2 * np.pi * k / We, Hp=48, Wp=64)[1]
    c = O.equi2pers(equi, yaw=-2 * np.pi * k / We, Hp=48, Wp=64)[1]
    assert min(np.abs(a.astype(int) - b).max(), np.abs(a.astype(int) - c).max()) <= 1


def test_splat_oracle_basics():
    """Unit-sphere sanity of the specified point pass: one point per face centre lands on the face
    centre pixel, nearest wins, ties go to the lowest index."""
    res = 8
    c2w = np.eye(4)[None]
    w2c = O.face_w2c(c2w).astype(np.float32)
    dirs = np.array([[0, 0, 1], [1, 0, 0], [0, 0, -1], [-1, 0, 0], [0, -1, 0], [0, 1, 0]], dtype=np.float64)
    # two points per direction: far (index first) and near (index second) + duplicate of the near one
    xyz = np.concatenate([dirs * 2.0, dirs * 1.0, dirs * 1.0])
    rgb = np.arange(18 * 3, dtype=np.uint8).reshape(18, 3)
    pts4 = O.pack_points(xyz, rgb)
    keys = O.splat_keys(pts4, w2c, res, res / 2.0, 1e-6)
    idx = O.keys_to_index(keys)[0]
    for f in range(6):
        hit = np.argwhere(idx[f] >= 0)
        assert len(hit) == 1 and tuple(hit[0]) == (res // 2, res // 2)
        assert idx[f, res // 2, res // 2] == 6 + f  # near beats far; lower index wins the tie
    lut = O.cube_to_equirect_lut(32, 16, res)
    pano = O.resolve(keys, pts4, lut)
    assert pano.shape == (1, 16, 32, 3)
    assert set(map(tuple, pano.reshape(-1, 3))) <= {(0, 0, 0)} | {tuple(rgb[6 + f]) for f in range(6)}
