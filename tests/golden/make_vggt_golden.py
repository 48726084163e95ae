"""Generate tests/golden/vggt_golden.npz from the REFERENCE's own VGGT modules.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_vggt_golden.py
Builds the reference Aggregator / CameraHead / DPTHead (third_party/vggt/vggt) at a small configuration, checks that their
state-dict keys and shapes are exactly evoworld_b200.vggt.param_spec's, loads the seeded parameters of
evoworld_b200.vggt.random_state_dict into them, and stores their fp32 CPU outputs on a seeded 3-frame clip.  The vectors
pin oracle/vggt_torch.py (tests/test_oracle_vggt.py) and are compared directly with the CUDA path (tests/test_gpu_vggt.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path[:0] = ["/root/reference/third_party/vggt"]

from vggt.heads.camera_head import CameraHead  # noqa: E402
from vggt.heads.dpt_head import DPTHead  # noqa: E402
from vggt.models.aggregator import Aggregator  # noqa: E402

from evoworld_b200.vggt import param_spec, random_state_dict  # noqa: E402
from oracle.vggt_torch import SMALL_TEST_CONFIG as CFG, small_test_images  # noqa: E402

torch.manual_seed(0)
d = CFG["embed_dim"]
agg = Aggregator(img_size=CFG["img_size"], patch_size=CFG["patch_size"], embed_dim=d, depth=CFG["depth"], num_heads=CFG["num_heads"],
                 patch_embed="dinov2_vits14_reg").eval()
cam = CameraHead(dim_in=2 * d, num_heads=CFG["camera_heads"], trunk_depth=CFG["camera_trunk_depth"]).eval()
kw = dict(dim_in=2 * d, features=CFG["dpt_features"], out_channels=list(CFG["dpt_out_channels"]), intermediate_layer_idx=list(CFG["dpt_layers"]))
point = DPTHead(output_dim=4, activation="inv_log", conf_activation="expp1", **kw).eval()
depth = DPTHead(output_dim=2, activation="exp", conf_activation="expp1", **kw).eval()
mods = {"aggregator.": agg, "camera_head.": cam, "point_head.": point, "depth_head.": depth}

# the reference's own keys / shapes == our spec, in the same order
ref_keys = [(pre + k, tuple(v.shape)) for pre, m in mods.items() for k, v in m.state_dict().items()]
assert ref_keys == [(k, tuple(s)) for k, s in param_spec(CFG).items()], "param_spec differs from the reference state dict"

sd = random_state_dict(CFG, seed=CFG["seed"])
for pre, m in mods.items():
    m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)

images = small_test_images()
out = {}
with torch.no_grad():
    toks, start = agg(images)
    assert start == 5
    for i, t in enumerate(toks):
        out[f"tokens_{i}"] = t.numpy()
    for i, pe in enumerate(cam(toks, num_iterations=CFG["camera_iterations"])):
        out[f"pose_enc_{i}"] = pe.numpy()
    dm, dc = depth(toks, images=images, patch_start_idx=start)
    out["depth"], out["depth_conf"] = dm.numpy(), dc.numpy()
    pm, pc = point(toks, images=images, patch_start_idx=start, frames_chunk_size=2)     # chunked path of the reference
    out["world_points"], out["world_points_conf"] = pm.numpy(), pc.numpy()
    # intermediate vectors for finer-grained checks
    x = ((images - agg._resnet_mean) / agg._resnet_std).view(-1, 3, *images.shape[-2:])
    out["dino_patch_tokens"] = agg.patch_embed(x)["x_norm_patchtokens"].numpy()
path = os.path.join(ROOT, "tests", "golden", "vggt_golden.npz")
np.savez_compressed(path, **out)
print({k: v.shape for k, v in out.items()}, os.path.getsize(path))
