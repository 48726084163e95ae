"""CPU precision ledger of the VGGT-1B path at the loop's resolution (2 x 392 x 518, the seeded weights / frames of
tests/golden/vggt_1b_golden.npz = outputs of the reference's own modules in fp32):
  1. the product's host orchestration with every kernel replaced by a torch restatement that rounds where the kernels round
     (tests/ops_emulation.py) — predicts the errors the GPU measures;
  2. where the pose error comes from: an exact fp32 camera head on the fp16-path tokens;
  3. the reference's own operating mode — aggregator under bf16 autocast (unified_loop_consistency.py:131-136), fp32 heads.
Needs no GPU and no reference checkout:  python tools/vggt_precision_ledger.py   (about 3 minutes)"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import ops_emulation as E  # noqa: E402

import evoworld_b200.vggt as V  # noqa: E402
from evoworld_b200 import _lib, ops  # noqa: E402
from oracle import vggt_torch as O  # noqa: E402

for n in E.ALL:
    setattr(ops, n, getattr(E, n))
_lib.require_cuda = lambda t, n: None
V._check_device = lambda d: None
cfg = dict(V.DEFAULT_CONFIG)
sd = V.random_state_dict(cfg, seed=O.FULL_TEST_SEED)
vg = np.load(ROOT / "tests" / "golden" / "vggt_1b_golden.npz")
rel = lambda a, b: float(np.linalg.norm(a.double().numpy() - b) / np.linalg.norm(b))
sub = lambda t: torch.cat([t[:, :, :5], t[:, :, 5::16]], 2)
images = O.full_test_images()
with torch.no_grad():
    m = V.VGGT(**cfg)
    m.load_state_dict(sd)
    out = m(images)
    got = O.subsample_full({k: v for k, v in out.items() if k != "images"})
    print("1. fp16-operand path (emulated rounding points) vs the reference in fp32:",
          {k: f"{rel(got[k], vg[k]):.2e}" for k in ("pose_enc", "depth", "depth_conf", "world_points", "world_points_conf")})
    pairs, (B, S, P, H, W) = m._aggregate(images, keep={cfg["depth"] - 1})
    fr, gl = pairs[cfg["depth"] - 1]
    last = torch.cat([fr.view(B, S, P, -1), gl.view(B, S, P, -1)], -1)
    print(f"2. last-layer tokens {rel(sub(last), vg['tokens_last']):.2e}, camera tokens {rel(last[:, :, 0], vg['tokens_last'][:, :, 0]):.2e}; "
          f"pose with an exact fp32 camera head on these tokens {rel(O.camera_head(last, sd, cfg)[-1], vg['pose_enc']):.2e}, "
          f"with the fp16-operand camera head {rel(m._camera(pairs[cfg['depth'] - 1], B, S, P)[-1], vg['pose_enc']):.2e}")
    with torch.autocast("cpu", dtype=torch.bfloat16):
        toks, start = O.aggregator(images, sd, cfg)
    toks = [t.float() for t in toks]
    pose = O.camera_head(toks[-1], sd, cfg)[-1]
    depth, conf = O.dpt_head(toks, 392, 518, start, sd, cfg, "depth_head.", "exp")
    print(f"3. the reference's own mode (bf16-autocast aggregator, fp32 heads) vs fp32: last-layer tokens {rel(sub(toks[-1]), vg['tokens_last']):.2e}, "
          f"pose {rel(pose, vg['pose_enc']):.2e}, depth {rel(depth[:, :, ::4, ::4], vg['depth']):.2e}, confidence {rel(conf[:, :, ::4, ::4], vg['depth_conf']):.2e}")
