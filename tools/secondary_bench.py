"""Achieved HBM bandwidth of the secondary reprojection kernels: Plücker, Equi2Pers warp, depth lift,
point packing, confidence select.  Algorithmic bytes as stated in DESIGN.md §2; CUDA events, L2 flushed per trial."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from evoworld_b200 import reprojection as R, synthetic
from evoworld_b200.equi2pers import Equi2Pers
from evoworld_b200.lift import lift_depth_device
from evoworld_b200.plucker import equirectangular_to_ray, ray_c2w_to_plucker

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, n=7):
    for _ in range(3):
        fn()
    ms = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return sorted(ms)[n // 2]


def report(name, nbytes, ms):
    print(f"{name:34s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {nbytes / ms / 1e6:8.1f} GB/s  {nbytes / ms / 1e6 / peak:6.1%} of measured HBM peak", flush=True)


# Plücker: T=25 frames of 72x128 latents -> 24 B per latent pixel per frame (tiny: launch bound)
ray = torch.from_numpy(equirectangular_to_ray(72, 128)).to(dev)
c2w = torch.from_numpy(synthetic.euler_c2w(synthetic.curve_trajectory())[:25, :3, :4]).float().to(dev)
report("plucker T=25 72x128", 25 * 72 * 128 * 24 + 72 * 128 * 12, timed(lambda: ray_c2w_to_plucker(ray, c2w)))
ray_big = torch.from_numpy(equirectangular_to_ray(576, 1024)).to(dev)
report("plucker T=25 576x1024", 25 * 576 * 1024 * 24 + 576 * 1024 * 12, timed(lambda: ray_c2w_to_plucker(ray_big, c2w)))

# Equi2Pers: 49 frames 576x1024 -> 384x512 (segment 3 re-warps every frame generated so far)
B = 49
equi = torch.randint(0, 256, (B, 3, 576, 1024), dtype=torch.uint8, device=dev)
e2p = Equi2Pers(384, 512, 90.0, mode="bilinear", device=dev)
rots = [{"yaw": 0.1 * i, "pitch": 0.0, "roll": 0.0} for i in range(B)]
report(f"equi2pers B={B} via Equi2Pers() pure-yaw fast path (table + H2D of {B} shifts)", B * 3 * (576 * 1024 + 384 * 512), timed(lambda: e2p(equi, rots)))
e2p_gen = Equi2Pers(384, 512, 90.0, mode="bilinear", device=dev, fast_yaw=False)
report(f"equi2pers B={B} via Equi2Pers(fast_yaw=False) (host matrices + H2D)", B * 3 * (576 * 1024 + 384 * 512), timed(lambda: e2p_gen(equi, rots)))
shift = torch.tensor([e2p.yaw_shift_px(r["yaw"], 1024) for r in rots], dtype=torch.float32, device=dev)
table = e2p._table(dev, 576, 1024)
yaw_out = torch.empty((B, 3, 384, 512), dtype=torch.uint8, device=dev)
from evoworld_b200 import _lib as _l2
report(f"equi2pers B={B} kernel only (evw_equi2pers_yaw_u8)", B * 3 * (576 * 1024 + 384 * 512),
       timed(lambda: _l2.check(_l2.lib().evw_equi2pers_yaw_u8(equi.data_ptr(), table.data_ptr(), shift.data_ptr(), yaw_out.data_ptr(), B, 3,
                                                             576, 1024, 384, 512, _l2.stream_ptr(dev)))))
from evoworld_b200.equi2pers import pix2dir_matrix
from evoworld_b200 import _lib
mats = torch.from_numpy(np.stack([pix2dir_matrix(r["yaw"], 0.0, 0.0, 384, 512, 90.0) for r in rots]).astype(np.float32).reshape(B, 9)).to(dev)
e2p_out = torch.empty((B, 3, 384, 512), dtype=torch.uint8, device=dev)
L = _lib.lib()
report(f"equi2pers B={B} kernel only (evw_equi2pers_u8)", B * 3 * (576 * 1024 + 384 * 512),
       timed(lambda: _lib.check(L.evw_equi2pers_u8(equi.data_ptr(), mats.data_ptr(), e2p_out.data_ptr(), B, 3, 576, 1024, 384, 512,
                                                   _lib.stream_ptr(dev)))))

# depth lift + pack + select at S=25 and S=49
for S in (25, 49):
    p = synthetic.reprojection_predictions(S=S, H=392, W=518, seed=0)
    depth, extr, intr = (torch.from_numpy(p[k]).to(dev) for k in ("depth", "extrinsic", "intrinsic"))
    conf = torch.from_numpy(p["depth_conf"]).to(dev)
    images = torch.from_numpy(p["images"]).to(dev)
    npix = S * 392 * 518
    report(f"lift S={S} (f64 out)", npix * (4 + 24), timed(lambda: lift_depth_device(depth, extr, intr, torch.float64)))
    report(f"lift S={S} (f32 out)", npix * (4 + 12), timed(lambda: lift_depth_device(depth, extr, intr, torch.float32)))
    pts64 = lift_depth_device(depth, extr, intr, torch.float64)
    from evoworld_b200.memory import PointMemory
    mem = PointMemory(392, 518, capacity_frames=S, device=dev)
    def fused():
        mem.reset(); mem.append(depth, conf, images, extr, intr)
    report(f"lift+pack fused S={S} (PointMemory.append)", npix * (4 + 12 + 16 + 8), timed(fused))
    report(f"pack S={S} (f64 xyz + f32 rgb)", npix * (24 + 12 + 16), timed(lambda: R.pack_points_device(pts64.reshape(-1, 3), images_nchw=images)))
    pts4 = R.pack_points_device(pts64.reshape(-1, 3), images_nchw=images)
    # select: 4 B/point x (3 histogram + 1 rank + 2 compaction passes) + 16 B in + 16 B out per kept point (~half)
    report(f"conf_select S={S} (50th pct)", npix * (4 * 6 + 16) + (npix // 2) * 16, timed(lambda: R.conf_select_device(conf, pts4, 50.0)))
