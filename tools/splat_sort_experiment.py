"""Experiment (round 2): how much of the splat's L2-atomic cost is point ORDER?  The benchmark cloud is in source-pixel
order with i.i.d. depths, so neighbouring threads hit unrelated z-buffer cells (one RED sector per point-view).  Re-order
the cloud (3D Morton / cell order of one target view / random) with torch and time the unchanged 24-view splat set."""
import os, sys
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
n = int(count.item())
pts = sel[:n].clone()
algo = bench.reproj_algorithmic_bytes(n, c["V"], c["face_res"], c["pano"])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = torch.empty((c["V"], c["pano"][0], c["pano"][1], 3), dtype=torch.uint8, device=dev)


def part1by2(x):
    x = x & 0x3FF
    x = (x | (x << 16)) & 0x30000FF
    x = (x | (x << 8)) & 0x300F00F
    x = (x | (x << 4)) & 0x30C30C3
    x = (x | (x << 2)) & 0x9249249
    return x


def morton3(xyz):
    lo, hi = xyz.amin(0), xyz.amax(0)
    q = ((xyz - lo) / (hi - lo + 1e-9) * 1023).long().clamp_(0, 1023)
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << 1) | (part1by2(q[:, 2]) << 2)


def cell_of_view(xyz, m):  # (face, v, u) of the cube cell in view m's frame -> coarse 8x8-tile order
    Xc = xyz @ m[:, :3].T + m[:, 3]
    ax = Xc.abs().argmax(1)
    major = Xc.gather(1, ax[:, None])[:, 0]
    face = ax * 2 + (major < 0).long()
    others = torch.stack([Xc[:, (1, 0, 0)].gather(1, ax[:, None])[:, 0], Xc[:, (2, 2, 1)].gather(1, ax[:, None])[:, 0]], 1)
    uv = ((others / major.abs()[:, None]) * 0.5 + 0.5).clamp(0, 0.9999)
    u, v = (uv[:, 0] * 512).long(), (uv[:, 1] * 512).long()
    return (face << 18) | ((v >> 3) << 12) | ((u >> 3) << 6) | ((v & 7) << 3) | (u & 7)


def bench_order(name, order):
    q = pts if order is None else pts[order].contiguous()
    scene = R.PointScene(q, torch.tensor([n], dtype=torch.int64, device=dev))
    for G in (2, 4):
        zb = torch.empty(R.splat_workspace_bytes(G, c["face_res"]), dtype=torch.uint8, device=dev)
        run = lambda: R.splat_to_panoramas_device(scene, w2c, c["pano"][1], c["pano"][0], c["face_res"], G, out=out, zbuf=zb)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ms = []
        for _ in range(7):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        m = sorted(ms)[len(ms) // 2]
        print(f"{name:28s} G={G}: median {m:.4f} ms (min {min(ms):.4f})  {algo / m / 1e6:8.1f} GB/s algorithmic = "
              f"{algo / m / 1e6 / 6546.6:.3f} of HBM peak, nonzero pixels {int((out != 0).any(-1).sum())}", flush=True)


xyz = pts[:, :3].float()
print(f"points {n}, algorithmic bytes {algo / 1e9:.3f} GB")
bench_order("source-pixel order (as is)", None)
bench_order("3D Morton (10 bit/axis)", torch.argsort(morton3(xyz)))
bench_order("cell order of target view 0", torch.argsort(cell_of_view(xyz, w2c[0])))
bench_order("cell order of target view 12", torch.argsort(cell_of_view(xyz, w2c[12])))
bench_order("random permutation", torch.randperm(n, device=dev))
# how long does a torch sort of the keys take (upper bound for a hand-written radix sort)
k = morton3(xyz)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); o = torch.argsort(k); q = pts[o]; e1.record(); torch.cuda.synchronize()
print(f"torch argsort + gather of {n} points: {e0.elapsed_time(e1):.3f} ms")
