"""Micro-benchmark of the 24-view splat + resolve set (config 3 segment 1) over views-per-pass / flag combinations.
Every combination is checked byte-for-byte against the first one."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoworld_b200 import reprojection as R, synthetic
from evoworld_b200.lift import lift_depth_device
import bench

dev = torch.device("cuda:0")
c = bench.REPROJ_CFG
p = synthetic.reprojection_predictions(S=c["S"], H=c["H"], W=c["W"], seed=0)
depth, extr, intr = (torch.from_numpy(p[k]).to(dev) for k in ("depth", "extrinsic", "intrinsic"))
conf = torch.from_numpy(p["depth_conf"]).to(dev)
images = torch.from_numpy(p["images"]).to(dev)
pts64 = lift_depth_device(depth, extr, intr, torch.float64)
pts4_all = R.pack_points_device(pts64.reshape(-1, 3), images_nchw=images)
tgt = R.SceneBuilder(dev).align_extrinsics(p["camera_pose"], p["extrinsic"], c["V"], "bench_0", False)
w2c = torch.from_numpy(R.front_w2c_matrices(tgt)).to(dev)
sel, _, count, _ = R.conf_select_device(conf, pts4_all, c["conf_thres"])
scene = R.PointScene(sel, count)
n = scene.num_points()
algo = bench.reproj_algorithmic_bytes(n, c["V"], c["face_res"], c["pano"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty((c["V"], c["pano"][0], c["pano"][1], 3), dtype=torch.uint8, device=dev)
ref = None
print(f"points {n}, algorithmic bytes {algo / 1e9:.3f} GB")
L = R._lib.lib()
ref_ck = None
for G, overlap, ctas, by_role, ck in [(4, False, 0, False, False), (4, True, 8, False, False), (2, True, 8, False, False),
                                      (4, True, 8, True, False), (8, True, 8, False, False), (4, False, 0, False, True),
                                      (4, True, 8, False, True), (2, True, 8, False, True)]:
    pretest = False
    L.evw_set_splat_ctas_per_sm(ctas)
    zb = torch.empty(R.splat_workspace_bytes(G, c["face_res"], R.splat_flags(pretest, overlap, by_role, ck)), dtype=torch.uint8, device=dev)
    run = lambda: R.splat_to_panoramas_device(scene, w2c, c["pano"][1], c["pano"][0], c["face_res"], G, out=out, zbuf=zb,
                                              pretest=pretest, overlap=overlap, by_role=by_role, color_keys=ck)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if ck:
        if ref_ck is None:
            ref_ck = out.clone()
        same = f"{bool(torch.equal(out, ref_ck))} (colour keys; {int((out != ref).any(-1).sum())} pixels differ from the index rule)"
    else:
        if ref is None:
            ref = out.clone()
        same = bool(torch.equal(out, ref))
    ms = []
    for _ in range(7):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    m = sorted(ms)[len(ms) // 2]
    print(f"G={G} overlap={int(overlap)} by_role={int(by_role)} colour_keys={int(ck)} ctas/SM={ctas}: median {m:.4f} ms (min {min(ms):.4f})  {algo / m / 1e6:8.1f} GB/s algorithmic "
          f"{n * c['V'] / m / 1e6:8.1f} G pv/s  identical={same}", flush=True)
L.evw_set_splat_ctas_per_sm(0)
