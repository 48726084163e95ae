"""Generate tests/golden/pipeline_golden.npz from the REFERENCE's own functions (build container only:
needs /root/reference).  Pins evoworld_b200/image_ops.py::resize_with_antialiasing against
evoworld/trainer/trainer_utils.py:68-179 (the same function as pipeline_evoworld.py:746-850, which cannot be
imported here because diffusers is absent).
    python tests/golden/make_pipeline_golden.py
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path[:0] = [REF]
from evoworld.trainer.trainer_utils import _resize_with_antialiasing  # noqa: E402

out = {}
g = torch.Generator().manual_seed(77)
# same down-scaling factors as 576x1024 -> 224x224 (2.571, 4.571: windows 3 and 7), small enough to commit
x = torch.rand(2, 3, 90, 160, generator=g) * 2 - 1
out["aa_in_90x160"] = x.numpy()
out["aa_out_35x35"] = _resize_with_antialiasing(x, (35, 35)).numpy()
# up-scaling (factor < 1: sigma clamps to 0.001, window 3) and an even window before the odd fix-up (factor 3.0 -> k=4 -> 5)
y = torch.rand(1, 3, 24, 36, generator=g)
out["aa_in_24x36"] = y.numpy()
out["aa_out_40x12"] = _resize_with_antialiasing(y, (40, 12)).numpy()
# the real shape: checksum-sized summary of 576x1024 -> 224x224
z = torch.rand(1, 3, 576, 1024, generator=torch.Generator().manual_seed(5)) * 2 - 1
r = _resize_with_antialiasing(z, (224, 224))
out["aa_full_seed"] = np.array(5)
out["aa_full_rows"] = r[0, :, ::16, :].numpy()
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pipeline_golden.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, os.path.getsize(dst), "bytes")
