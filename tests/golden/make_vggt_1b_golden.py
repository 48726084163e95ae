"""Generate tests/golden/vggt_1b_golden.npz: outputs of the REFERENCE's own VGGT-1B modules (default constructor arguments:
DINOv2 ViT-L/14 with registers, 24 + 24 aggregator blocks, camera head, DPT depth / point heads) on seeded weights and two
seeded 392 x 518 frames, sub-sampled to keep the file small.  Build container only (needs /root/reference):
    python tests/golden/make_vggt_1b_golden.py
Pins oracle/vggt_torch.py at the full architecture and the loop's resolution (tests/test_oracle_vggt.py): the 37 x 37 DINOv2
position table resized to 28 x 37 with antialiasing, 16 heads, the real layer indices 4 / 11 / 17 / 23 of the DPT heads."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path[:0] = ["/root/reference/third_party/vggt"]

from vggt.heads.camera_head import CameraHead  # noqa: E402
from vggt.heads.dpt_head import DPTHead  # noqa: E402
from vggt.models.aggregator import Aggregator  # noqa: E402

from evoworld_b200.vggt import DEFAULT_CONFIG, param_spec, random_state_dict  # noqa: E402
from oracle.vggt_torch import FULL_TEST_SEED, full_test_images, subsample_full  # noqa: E402

t0 = time.time()
agg = Aggregator().eval()
cam = CameraHead(dim_in=2048).eval()
point = DPTHead(dim_in=2048, output_dim=4, activation="inv_log", conf_activation="expp1").eval()
depth = DPTHead(dim_in=2048, output_dim=2, activation="exp", conf_activation="expp1").eval()
mods = {"aggregator.": agg, "camera_head.": cam, "point_head.": point, "depth_head.": depth}
ref_keys = [(pre + k, tuple(v.shape)) for pre, m in mods.items() for k, v in m.state_dict().items()]
assert ref_keys == [(k, tuple(s)) for k, s in param_spec(DEFAULT_CONFIG).items()], "param_spec differs from the reference VGGT-1B state dict"
sd = random_state_dict(DEFAULT_CONFIG, seed=FULL_TEST_SEED)
for pre, m in mods.items():
    m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
images = full_test_images()
with torch.no_grad():
    toks, start = agg(images)
    out = {"pose_enc": cam(toks)[-1]}
    out["depth"], out["depth_conf"] = depth(toks, images=images, patch_start_idx=start)
    out["world_points"], out["world_points_conf"] = point(toks, images=images, patch_start_idx=start)
    out["tokens_last"] = toks[-1]
    out["tokens_4"] = toks[4]
small = {k: v.numpy() for k, v in subsample_full(out).items()}
path = os.path.join(ROOT, "tests", "golden", "vggt_1b_golden.npz")
np.savez_compressed(path, **small)
print({k: v.shape for k, v in small.items()}, os.path.getsize(path), f"{time.time() - t0:.0f} s")
