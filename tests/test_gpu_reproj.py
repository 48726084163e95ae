"""GPU parity tests of the reprojection path: the CUDA kernels (through the C ABI / the
reference-named shims) against the oracle on identical seeded inputs.  Integer/index work is
bit-exact; floating-point kernels state their tolerance here."""
import numpy as np
import pytest
import torch

from evoworld_b200 import reprojection as R
from evoworld_b200 import synthetic
from evoworld_b200.equi2pers import Equi2Pers
from evoworld_b200.lift import lift_depth_device, unproject_depth_map_to_point_map
from evoworld_b200.plucker import equirectangular_to_ray, ray_c2w_to_plucker
from oracle import reproj_np as O

pytestmark = pytest.mark.gpu


def test_plucker_golden_and_benchmark_shape(golden, cuda_device, built_lib):
    ray = torch.from_numpy(golden["ray_12x24"]).to(cuda_device)
    out = ray_c2w_to_plucker(ray, torch.from_numpy(golden["c2w_3x4_rel"]).to(cuda_device))
    assert out.shape == (14, 6, 12, 24) and out.dtype == torch.float32
    np.testing.assert_allclose(out.cpu().numpy(), golden["plucker"], rtol=0, atol=1e-6)  # fp tolerance: 1e-6 abs
    ray = torch.from_numpy(equirectangular_to_ray(72, 128))
    c2w = torch.from_numpy(synthetic.euler_c2w(synthetic.curve_trajectory())[:25, :3, :4]).float()
    want = O.ray_c2w_to_plucker(ray, c2w)
    got = ray_c2w_to_plucker(ray.to(cuda_device), c2w.to(cuda_device))
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=0, atol=2e-6)


def test_plucker_accepts_4x4_and_single_frame(cuda_device, built_lib):
    ray = torch.from_numpy(equirectangular_to_ray(5, 7)).to(cuda_device)
    c2w = torch.eye(4, device=cuda_device)[None]
    out = ray_c2w_to_plucker(ray, c2w)
    np.testing.assert_allclose(out[0, :3].permute(1, 2, 0).cpu().numpy(), ray.cpu().numpy(), atol=0)
    assert float(out[0, 3:].abs().max()) == 0.0


@pytest.mark.parametrize("yaw", [0.0, 0.7, -2.4, 3.1])
def test_equi2pers_vs_oracle(yaw, cuda_device, built_lib):
    rng = np.random.default_rng(5)
    He, We = 144, 256
    yy, xx = np.meshgrid(np.arange(He), np.arange(We), indexing="ij")
    equi = np.stack([(xx * 255 // (We - 1)), (yy * 255 // (He - 1)), rng.integers(0, 256, (He, We))]).astype(np.uint8)
    rots = {"pitch": 0, "roll": 0, "yaw": yaw}
    # general kernel (any rotation) vs its restatement, and the pure-yaw fast path (the default) vs ITS restatement
    for fast, want in ((False, O.equi2pers(equi, yaw, Hp=96, Wp=128)), (True, O.equi2pers_yaw(equi, yaw, Hp=96, Wp=128))):
        got = Equi2Pers(height=96, width=128, fov_x=90, mode="bilinear", fast_yaw=fast)(equi=equi, rots=rots)
        assert got.shape == want.shape and got.dtype == np.uint8
        diff = np.abs(got.astype(int) - want.astype(int))
        # smooth channels: asinf/atan2f differ by ulps between libm and CUDA -> at most 1 LSB, rarely
        assert diff[:2].max() <= 1 and (diff[:2] > 0).mean() < 0.02, fast
        # noise channel: a coordinate ulp moves the bilinear weights by ~1e-5 -> still <= 1 LSB
        assert diff[2].max() <= 1 and (diff[2] > 0).mean() < 0.02, fast
    # the two forms describe the same warp: they differ by float32 rounding of the longitude (~1e-4 px)
    a = Equi2Pers(height=96, width=128, fov_x=90, fast_yaw=True)(equi=equi, rots=rots).astype(int)
    b = Equi2Pers(height=96, width=128, fov_x=90, fast_yaw=False)(equi=equi, rots=rots).astype(int)
    assert np.abs(a - b).max() <= 1 and (np.abs(a - b)[:2] > 0).mean() < 0.02 and (np.abs(a - b)[2] > 0).mean() < 0.08
    # any pitch / roll falls back to the general kernel
    c = Equi2Pers(height=96, width=128, fov_x=90)(equi=equi, rots={"pitch": 0.1, "roll": 0, "yaw": yaw})
    wc = O.equi2pers(equi, yaw, pitch=0.1, Hp=96, Wp=128)
    assert np.abs(c.astype(int) - wc.astype(int)).max() <= 1


def test_equi2pers_full_size_properties(cuda_device, built_lib):
    """576x1024 -> 384x512 (reference sizes): a horizontal roll of the panorama equals a yaw change."""
    rng = np.random.default_rng(6)
    He, We = 576, 1024
    base = rng.integers(0, 256, size=(3, He // 8, We // 8), dtype=np.uint8)
    equi = np.repeat(np.repeat(base, 8, axis=1), 8, axis=2)
    e2p = Equi2Pers(height=384, width=512, fov_x=90, mode="bilinear")
    k = 128
    a = e2p(equi=np.roll(equi, k, axis=2), rots={"pitch": 0, "roll": 0, "yaw": 0.0})
    b = e2p(equi=equi, rots={"pitch": 0, "roll": 0, "yaw": 2 * np.pi * k / We})
    c = e2p(equi=equi, rots={"pitch": 0, "roll": 0, "yaw": -2 * np.pi * k / We})
    assert a.shape == (3, 384, 512)
    d = min((np.abs(a.astype(int) - b) > 1).mean(), (np.abs(a.astype(int) - c) > 1).mean())
    assert d < 0.01
    # batched torch input == per-image numpy input
    t = torch.from_numpy(np.stack([equi, np.roll(equi, 5, axis=1)])).to(cuda_device)
    out = e2p(equi=t, rots=[{"yaw": 0.3}, {"yaw": -1.0}])
    np.testing.assert_array_equal(out[0].cpu().numpy(), e2p(equi=equi, rots={"yaw": 0.3}))


def test_lift_golden(golden, cuda_device, built_lib):
    pts = unproject_depth_map_to_point_map(golden["lift_depth"], golden["pe_extr"][0], golden["pe_intr"][0])
    assert pts.dtype == np.float64 and pts.shape == (3, 14, 18, 3)
    # fp tolerance: the reference evaluates -R^T t in float32 matmul (order unspecified) -> 2e-6 abs
    np.testing.assert_allclose(pts, golden["lift_points"], rtol=0, atol=2e-6)


def test_lift_benchmark_shape(cuda_device, built_lib):
    p = synthetic.reprojection_predictions(S=4)
    want = O.unproject_depth_map_to_point_map(p["depth"], p["extrinsic"], p["intrinsic"])
    got = unproject_depth_map_to_point_map(torch.from_numpy(p["depth"]), p["extrinsic"], p["intrinsic"])
    np.testing.assert_allclose(got, want, rtol=0, atol=5e-6)
    g32 = lift_depth_device(torch.from_numpy(p["depth"]).to(cuda_device), torch.from_numpy(p["extrinsic"]).to(cuda_device),
                            torch.from_numpy(p["intrinsic"]).to(cuda_device), torch.float32)
    np.testing.assert_allclose(g32.cpu().numpy(), want.astype(np.float32), rtol=0, atol=1e-5)


def _select_oracle(conf, q):
    thr = 0.0 if q == 0.0 else np.percentile(conf, q)
    return np.nonzero(conf >= thr)[0], thr


@pytest.mark.parametrize("n", [1, 2, 7, 257, 4001, 100_003, 1_000_000])
@pytest.mark.parametrize("q", [50.0, 30.0, 0.0, 99.5, 100.0])
def test_conf_select_bit_exact(n, q, cuda_device, built_lib):
    rng = np.random.default_rng(n + int(q * 10))
    conf = (1 + np.exp(rng.normal(size=n))).astype(np.float32)
    if n > 4:
        conf[rng.integers(0, n, n // 3)] = conf[rng.integers(0, n, n // 3)]  # many ties
    idx_want, thr_want = _select_oracle(conf, q)
    pts4 = torch.arange(n * 4, dtype=torch.float32, device=cuda_device).reshape(n, 4)
    out, keep, count, thr = R.conf_select_device(torch.from_numpy(conf).to(cuda_device), pts4, q, want_index=True)
    c = int(count.item())
    assert c == len(idx_want)
    assert np.float32(thr.item()) == np.float32(thr_want)
    np.testing.assert_array_equal(keep[:c].cpu().numpy(), idx_want)
    np.testing.assert_array_equal(out[:c].cpu().numpy(), pts4.cpu().numpy()[idx_want])


def test_conf_select_edge_cases(cuda_device, built_lib):
    # all equal -> everything passes; negative values and zeros; NaN poisons the percentile
    for conf in (np.full(1000, 2.5, np.float32), np.linspace(-3, 3, 999).astype(np.float32),
                 np.zeros(10, np.float32)):
        idx_want, _ = _select_oracle(conf, 50.0)
        _, keep, count, _ = R.conf_select_device(torch.from_numpy(conf).to(cuda_device), None, 50.0, want_index=True)
        np.testing.assert_array_equal(keep[: int(count.item())].cpu().numpy(), idx_want)
    conf = np.ones(100, np.float32)
    conf[17] = np.nan
    _, _, count, thr = R.conf_select_device(torch.from_numpy(conf).to(cuda_device), None, 50.0)
    assert int(count.item()) == 0 and np.isnan(thr.item())  # numpy: percentile -> nan, mask all False


def test_pack_points_colours_truncate(golden, cuda_device, built_lib):
    imgs = torch.from_numpy(golden["cf_imgs"]).to(cuda_device)
    pts = torch.from_numpy(golden["cf_pts"]).to(cuda_device)
    pts4 = R.pack_points_device(pts, images_nchw=imgs)
    bits = pts4[:, 3].contiguous().view(torch.int32).cpu().numpy()
    cols = np.stack([bits & 0xFF, (bits >> 8) & 0xFF, (bits >> 16) & 0xFF], 1).astype(np.uint8)
    np.testing.assert_array_equal(cols, golden["cf_cols"])
    np.testing.assert_array_equal(pts4[:, :3].cpu().numpy(), golden["cf_pts"].astype(np.float32))


def _faces_gpu(pts4_np, w2c_np, res, dev):
    L = R._lib.lib()
    pts4 = torch.from_numpy(pts4_np).to(dev)
    w2c = torch.from_numpy(w2c_np).to(dev)
    V = w2c.shape[0]
    ws = torch.empty(L.evw_splat_workspace(1, res), dtype=torch.uint8, device=dev)
    win = torch.empty((V, 6, res, res), dtype=torch.int64, device=dev)
    R._lib.check(L.evw_splat_faces_debug(pts4.data_ptr(), pts4.shape[0], w2c.data_ptr(), V, res, res / 2.0, R.Z_NEAR,
                                         win.data_ptr(), ws.data_ptr(), ws.numel(), R._lib.stream_ptr(dev)))
    return win.cpu().numpy()


@pytest.mark.parametrize("n,res,V", [(0, 16, 1), (1, 16, 1), (5000, 32, 3), (200_000, 128, 2), (1_000_000, 512, 2)])
def test_splat_winner_index_bit_exact(n, res, V, cuda_device, built_lib):
    xyz, rgb = synthetic.random_cloud(max(n, 1), seed=n)
    xyz, rgb = xyz[:n], rgb[:n]
    if n > 10:  # exact duplicates -> z ties -> lowest index must win
        xyz[n // 2:n // 2 + n // 10] = xyz[: n // 10]
    pts4 = O.pack_points(xyz, rgb) if n else np.zeros((0, 4), np.float32)
    cam = synthetic.euler_c2w(synthetic.curve_trajectory())[30:30 + V]
    cam[:, :3, :3] *= 1.3  # the aligned target cameras carry a uniform scale (a12)
    w2c = O.face_w2c(cam).astype(np.float32)
    want = O.keys_to_index(O.splat_keys(pts4, w2c, res, res / 2.0, R.Z_NEAR))
    if n == 0:
        got = np.full_like(want, -1)
        lut = R.cube_lut_device(64, 32, res, cuda_device)
        scene = R.PointScene(torch.zeros((0, 4), device=cuda_device))
        pano = R.splat_to_panoramas_device(scene, torch.from_numpy(w2c).to(cuda_device), 64, 32, res, 1)
        assert int(pano.max()) == 0
    else:
        got = _faces_gpu(pts4, w2c, res, cuda_device)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("G", [1, 2, 3, 4, 6, 8])
def test_fused_panoramas_bit_exact(G, cuda_device, built_lib):
    n, res, V = 300_000, 128, 7
    xyz, rgb = synthetic.random_cloud(n, seed=11)
    cam = synthetic.euler_c2w(synthetic.curve_trajectory())[40:40 + V]
    want = O.render_panoramas(xyz, rgb, cam, res=res, width=400, height=200, z_near=R.Z_NEAR)
    scene = R.SceneBuilder().build_open3d_scene(xyz, rgb)
    w2c = torch.from_numpy(R.face_w2c_matrices(cam)).to(cuda_device)
    got = R.splat_to_panoramas_device(scene, w2c, 400, 200, res, G)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # device-side count: only the first n/2 points are live
    scene2 = R.PointScene(scene.pts4, torch.tensor([n // 2], dtype=torch.int64, device=cuda_device))
    got2 = R.splat_to_panoramas_device(scene2, w2c, 400, 200, res, G)
    want2 = O.render_panoramas(xyz[: n // 2], rgb[: n // 2], cam, res=res, width=400, height=200, z_near=R.Z_NEAR)
    np.testing.assert_array_equal(got2.cpu().numpy(), want2)


def test_cubemap_renderer_api(golden, cuda_device, built_lib):
    cr = R.CubemapRenderer()
    faces = {n: torch.from_numpy(golden["c2e_faces"][:, i]) for i, n in enumerate(R.FACE_ORDER)}
    pano = cr.cube_to_equirectangular_cuda(faces, 64, 32)
    np.testing.assert_array_equal(pano, golden["c2e_pano"])  # golden = the reference function itself
    xyz, rgb = synthetic.random_cloud(50_000, seed=2)
    scene = R.SceneBuilder().build_open3d_scene(xyz, rgb)
    cam = synthetic.euler_c2w(synthetic.curve_trajectory())[60]
    cube = cr.render_cubemap(scene, cam, res=(64, 64))
    pts4 = O.pack_points(xyz, rgb)
    idx = O.keys_to_index(O.splat_keys(pts4, O.face_w2c(cam[None]).astype(np.float32), 64, 32.0, R.Z_NEAR))[0]
    for fi, name in enumerate(R.FACE_ORDER):
        want = np.where(idx[fi][..., None] >= 0, rgb[np.maximum(idx[fi], 0)], 0).astype(np.uint8)
        np.testing.assert_array_equal(cube[name], want)
    face = cr.render_face(scene, cam @ R.CUBEMAP_TRANSFORMS["top"], res=(64, 64), do_flip=True)
    np.testing.assert_array_equal(face, cube["top"])


def test_predictions_to_target_view_end_to_end(tmp_path, cuda_device, built_lib):
    """The reference-named entry point on a reduced synthetic episode == the oracle composition."""
    p = synthetic.reprojection_predictions(S=25, H=56, W=74, seed=3)
    p["world_points_from_depth"] = O.unproject_depth_map_to_point_map(p["depth"], p["extrinsic"], p["intrinsic"])
    outdir = str(tmp_path / "rendered_panorama_vggt_open3d_0")
    panos = R.predictions_to_target_view(p, p["camera_pose"], conf_thres=50.0, prediction_mode="depth_unproject",
                                         num_target_view=24, outdir=outdir)
    assert panos.shape == (24, 1000, 2000, 3) and panos.dtype == np.uint8
    cols = O.extract_colors(p["images"])
    v, c = O.apply_confidence_filter(p["world_points_from_depth"], p["depth_conf"], cols, 50.0)
    tgt = O.align_extrinsics(p["camera_pose"], p["extrinsic"], 24, outdir)
    want = O.render_panoramas_cube(v, c, tgt, z_near=R.Z_NEAR)
    np.testing.assert_array_equal(panos, want)
    import cv2

    img = cv2.cvtColor(cv2.imread(outdir + "/07.png"), cv2.COLOR_BGR2RGB)
    np.testing.assert_array_equal(img, want[7])
    # numpy-returning filter API (reference signature)
    vv, cc, scale = R.PointCloudProcessor().filter_predictions(p, 50.0, prediction_mode="depth_unproject")
    np.testing.assert_array_equal(vv, v)
    np.testing.assert_array_equal(cc, c)
    assert scale > 0


def test_pinned_host_buffers_give_the_same_panoramas(tmp_path, cuda_device, built_lib):
    """The end-to-end bench path: page-locked torch predictions in, the renderer's page-locked buffer out
    (CubemapRenderer(pinned_output=True)) — byte-identical to the numpy-in / fresh-array-out default."""
    p = synthetic.reprojection_predictions(S=25, H=28, W=42, seed=5)
    pts = O.unproject_depth_map_to_point_map(p["depth"], p["extrinsic"], p["intrinsic"])
    preds = dict(world_points_from_depth=pts, depth_conf=p["depth_conf"], images=p["images"], extrinsic=p["extrinsic"])
    want = R.predictions_to_target_view(preds, p["camera_pose"], conf_thres=50.0, prediction_mode="Depthmap and Camera Branch",
                                        num_target_view=3, outdir=str(tmp_path / "a_0"))
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in preds.items() if k != "extrinsic"}
    pinned["extrinsic"] = p["extrinsic"]
    cr = R.CubemapRenderer(pinned_output=True)
    got = R.predictions_to_target_view(pinned, p["camera_pose"], conf_thres=50.0, prediction_mode="Depthmap and Camera Branch",
                                       num_target_view=3, outdir=str(tmp_path / "b_0"), cubemap_renderer=cr)
    np.testing.assert_array_equal(got, want)
    again = R.predictions_to_target_view(pinned, p["camera_pose"], conf_thres=50.0, prediction_mode="Depthmap and Camera Branch",
                                         num_target_view=3, outdir=str(tmp_path / "c_0"), cubemap_renderer=cr)
    assert again.ctypes.data == got.ctypes.data  # the page-locked buffer is reused (documented aliasing)
    np.testing.assert_array_equal(again, want)


def test_full_size_properties(cuda_device, built_lib):
    """BASELINE config-3 size (S=25 x 392x518 -> ~2.54 M points, 24 views, 512^2 faces, 2000x1000):
    size-independent properties — views-per-pass invariance, idempotence, subset consistency, and a
    checksum-of-checksums against the oracle on 2 of the 24 views."""
    p = synthetic.reprojection_predictions(S=25)
    dev = cuda_device
    pts = lift_depth_device(torch.from_numpy(p["depth"]).to(dev), torch.from_numpy(p["extrinsic"]).to(dev),
                            torch.from_numpy(p["intrinsic"]).to(dev), torch.float64)
    pts4 = R.pack_points_device(pts.reshape(-1, 3), images_nchw=torch.from_numpy(p["images"]).to(dev))
    out, keep, count, thr = R.conf_select_device(torch.from_numpy(p["depth_conf"]).to(dev), pts4, 50.0, want_index=True)
    n = int(count.item())
    assert abs(n - 25 * 392 * 518 // 2) <= 2
    assert bool((keep[1:n] > keep[: n - 1]).all())  # order-preserving compaction
    tgt = R.SceneBuilder().align_extrinsics(p["camera_pose"], p["extrinsic"], 24, "x_0", False)
    w2c = torch.from_numpy(R.front_w2c_matrices(tgt)).to(dev)
    scene = R.PointScene(out, count)
    a = R.splat_to_panoramas_device(scene, w2c, views_per_pass=4)
    b = R.splat_to_panoramas_device(scene, w2c, views_per_pass=8)
    c = R.splat_to_panoramas_device(R.PointScene(out[:n].clone()), w2c, views_per_pass=1, pretest=False)
    assert torch.equal(a, b) and torch.equal(a, c)
    assert torch.equal(a, R.splat_to_panoramas_device(scene, w2c, views_per_pass=4))
    assert float((a.sum(dim=-1) > 0).float().mean()) > 0.2  # panoramas are populated
    sel = [0, 23]
    want = O.resolve(O.splat_keys_cube(out[:n].cpu().numpy(), w2c[sel].cpu().numpy(), 512, 256.0, R.Z_NEAR),
                     out[:n].cpu().numpy(), O.cube_to_equirect_lut(2000, 1000, 512))
    np.testing.assert_array_equal(a[sel].cpu().numpy(), want)
    # the six-independent-cameras formulation agrees except on face-boundary / rounding pixels
    w6 = torch.from_numpy(R.face_w2c_matrices(tgt[sel])).to(dev)
    d = R.splat_to_panoramas_device(scene, w6, views_per_pass=2)
    assert float((d != a[sel]).any(dim=-1).float().mean()) < 2e-3


@pytest.mark.parametrize("n,res,V", [(1, 16, 1), (5000, 32, 3), (200_000, 128, 2), (1_000_000, 512, 2)])
def test_cube_splat_winner_index_bit_exact(n, res, V, cuda_device, built_lib):
    xyz, rgb = synthetic.random_cloud(n, seed=n + 1)
    if n > 10:
        xyz[n // 2:n // 2 + n // 10] = xyz[: n // 10]  # z ties -> lowest index wins
    if n >= 5000:  # points exactly on face boundaries and on the axes
        xyz[:6] = np.array([[1, 1, 1], [-1, 1, 1], [1, -1, 0], [0, 0, 2], [0, 3, 0], [-2, 0, 0]], dtype=np.float64)
    pts4 = O.pack_points(xyz, rgb)
    cam = synthetic.euler_c2w(synthetic.curve_trajectory())[30:30 + V]
    cam[:, :3, :3] *= 1.3
    if n >= 5000:
        cam[0] = np.eye(4)
    w2c = O.front_w2c(cam)
    want = O.keys_to_index(O.splat_keys_cube(pts4, w2c, res, res / 2.0, R.Z_NEAR))
    L = R._lib.lib()
    p = torch.from_numpy(pts4).to(cuda_device)
    w = torch.from_numpy(w2c).to(cuda_device)
    ws = torch.empty(L.evw_splat_workspace(1, res), dtype=torch.uint8, device=cuda_device)
    win = torch.empty((V, 6, res, res), dtype=torch.int64, device=cuda_device)
    R._lib.check(L.evw_splat_cube_faces_debug(p.data_ptr(), n, w.data_ptr(), V, res, res / 2.0, R.Z_NEAR, win.data_ptr(),
                                              ws.data_ptr(), ws.numel(), R._lib.stream_ptr(cuda_device)))
    np.testing.assert_array_equal(win.cpu().numpy(), want)


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_cube_fused_panoramas_bit_exact(G, cuda_device, built_lib):
    n, res, V = 300_000, 128, 7
    xyz, rgb = synthetic.random_cloud(n, seed=12)
    cam = synthetic.euler_c2w(synthetic.curve_trajectory())[40:40 + V]
    want = O.render_panoramas_cube(xyz, rgb, cam, res=res, width=400, height=200, z_near=R.Z_NEAR)
    scene = R.SceneBuilder().build_open3d_scene(xyz, rgb)
    w2c = torch.from_numpy(R.front_w2c_matrices(cam)).to(cuda_device)
    # every flag combination (pre-test, two-stream pass pipeline, role-split streams) must give the same bytes
    for pretest, overlap, by_role in [(True, False, False), (False, False, False), (False, True, False), (True, True, False),
                                      (False, True, True), (True, True, True)]:
        got = R.splat_to_panoramas_device(scene, w2c, 400, 200, res, G, pretest=pretest, overlap=overlap, by_role=by_role)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    # optional colour-key tie rule (EVW_SPLAT_COLOR_KEYS): bit-exact against the oracle evaluated with the same rule
    want_ck = O.render_panoramas_cube(xyz, rgb, cam, res=res, width=400, height=200, z_near=R.Z_NEAR, color_keys=True)
    for overlap in (False, True):
        got = R.splat_to_panoramas_device(scene, w2c, 400, 200, res, G, overlap=overlap, color_keys=True)
        np.testing.assert_array_equal(got.cpu().numpy(), want_ck)
    assert int((want_ck != want).any(-1).sum()) <= 8  # the two rules differ only where float32 depths tie exactly
    # a caller-provided workspace is reused across calls and passes: stale keys must never leak into a later view
    zb = torch.empty(R.splat_workspace_bytes(G, res, R.SPLAT_OVERLAP), dtype=torch.uint8, device=cuda_device)
    for _ in range(2):
        got = R.splat_to_panoramas_device(scene, w2c, 400, 200, res, G, zbuf=zb, overlap=True)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    scene2 = R.PointScene(scene.pts4, torch.tensor([n // 3], dtype=torch.int64, device=cuda_device))
    got2 = R.splat_to_panoramas_device(scene2, w2c, 400, 200, res, G)
    want2 = O.render_panoramas_cube(xyz[: n // 3], rgb[: n // 3], cam, res=res, width=400, height=200, z_near=R.Z_NEAR)
    np.testing.assert_array_equal(got2.cpu().numpy(), want2)


def test_lift_pack_fused_matches_lift_then_pack(cuda_device, built_lib):
    """evw_lift_pack_points == evw_lift_depth (f64) -> evw_pack_points, bit for bit (same camera / colour arithmetic)."""
    from evoworld_b200.lift import lift_depth_device
    from evoworld_b200.memory import PointMemory

    p = synthetic.reprojection_predictions(S=5, H=28, W=36, seed=3)
    dev = cuda_device
    depth, conf, images = (torch.from_numpy(p[k]).to(dev) for k in ("depth", "depth_conf", "images"))
    extr, intr = torch.from_numpy(p["extrinsic"]).to(dev), torch.from_numpy(p["intrinsic"]).to(dev)
    want = R.pack_points_device(lift_depth_device(depth, extr, intr, torch.float64).reshape(-1, 3), images_nchw=images)
    mem = PointMemory(28, 36, capacity_frames=8, device=dev).append(depth, conf, images, extr, intr)
    assert len(mem) == 5 * 28 * 36
    assert torch.equal(mem.pts4[:len(mem)].view(torch.int32), want.view(torch.int32))
    assert torch.equal(mem.conf[:len(mem)], conf.reshape(-1))


def test_point_memory_incremental_equals_one_shot(cuda_device, built_lib):
    """Appending a segment's new frames to the device-resident memory and filtering jointly gives the same PointScene
    (count, order, bits) as the reference-shaped one-shot path over all frames (filter_predictions on host arrays)."""
    from evoworld_b200.memory import PointMemory
    from oracle import reproj_np as O

    S, H, W = 7, 28, 36
    p = synthetic.reprojection_predictions(S=S, H=H, W=W, seed=5)
    dev = cuda_device
    preds = dict(p)
    # the caller's lift (unified_loop_consistency.py:366) through the drop-in: float64 host array, as the reference passes it
    preds["world_points_from_depth"] = unproject_depth_map_to_point_map(p["depth"], p["extrinsic"], p["intrinsic"])
    assert preds["world_points_from_depth"].dtype == np.float64
    one_shot, _ = R.PointCloudProcessor(dev).filter_predictions_device(preds, 50.0, prediction_mode="depth_unproject")
    mem = PointMemory(H, W, capacity_frames=S, device=dev)
    for a, b in ((0, 3), (3, 4), (4, 7)):  # three "segments"
        mem.append(*(torch.from_numpy(p[k][a:b]) for k in ("depth", "depth_conf", "images", "extrinsic", "intrinsic")))
    scene = mem.scene(50.0)
    n = one_shot.num_points()
    assert scene.num_points() == n and 0 < n < S * H * W
    assert torch.equal(scene.pts4[:n].view(torch.int32), one_shot.pts4[:n].view(torch.int32))
    # a later window only (frames 3..7) == one shot over those frames
    sub = {k: (v[3:] if k != "camera_pose" else v) for k, v in preds.items()}
    want, _ = R.PointCloudProcessor(dev).filter_predictions_device(sub, 30.0, prediction_mode="depth_unproject")
    got = mem.scene(30.0, first_frame=3)
    m = want.num_points()
    assert got.num_points() == m and torch.equal(got.pts4[:m].view(torch.int32), want.pts4[:m].view(torch.int32))
    with pytest.raises(ValueError, match="capacity"):
        mem.append(*(torch.from_numpy(p[k][:1]) for k in ("depth", "depth_conf", "images", "extrinsic", "intrinsic")))
    mem.reset()
    assert len(mem) == 0
    with pytest.raises(ValueError, match="no frames"):
        mem.scene()
