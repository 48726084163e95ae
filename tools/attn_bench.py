"""Micro-benchmark of the attention / norm kernels on the UNet's shapes (T=14, 576x1024).
Spatial attention is timed for every kernel variant (evw_set_attention_variant) with a rel-L2 check against
torch SDPA (fp32) on the L1 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from evoworld_b200 import ops, _lib

dev = torch.device("cuda:0")
L = _lib.lib()
torch.manual_seed(0)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


VARIANTS = {0: "v8 (default)", 1: "v8 poly1/8", 2: "v8 stagger", 3: "v7 1thr/row", 4: "v7 poly1/4", 5: "v8 pair barriers"}
if "--new-only" in sys.argv:
    VARIANTS = {0: VARIANTS[0]}
if "--pair" in sys.argv:
    VARIANTS = {0: VARIANTS[0], 5: VARIANTS[5], 10: VARIANTS[0] + " again", 15: VARIANTS[5] + " again"}
shapes = {"L0 spatial": (28, 9216, 5), "L1 spatial": (28, 2304, 10), "L2 spatial": (28, 576, 20), "mid": (28, 144, 20),
          "VGGT global S25": (1, 25 * 1041, 16), "VGGT frame S25": (25, 1041, 16)}
qkvs = {k: torch.randn(f * s, 3 * h * 64, device=dev).half() for k, (f, s, h) in shapes.items()}
f_, s_, h_ = shapes["L1 spatial"]
x = qkvs["L1 spatial"][: 2 * s_].float().view(2, s_, 3, h_, 64).permute(2, 0, 3, 1, 4)
want = F.scaled_dot_product_attention(x[0], x[1], x[2]).permute(0, 2, 1, 3).reshape(2 * s_, h_ * 64)
for var, label in VARIANTS.items():
    L.evw_set_attention_variant(var % 10)
    got = ops.spatial_attention(qkvs["L1 spatial"][: 2 * s_].contiguous(), 2, s_, h_).float()
    err = float((got - want).norm() / want.norm())
    row = []
    for name, (F_, S, H) in shapes.items():
        ms = timeit(lambda: ops.spatial_attention(qkvs[name], F_, S, H))
        row.append(f"{name} {ms:7.3f} ms {4.0 * F_ * H * S * S * 64 / ms / 1e9:7.1f} TF/s")
    print(f"[{var:2d}] {label:24s} relL2={err:.2e} | " + " | ".join(row), flush=True)
L.evw_set_attention_variant(-2)
for name, (B, T, S, H) in {"L0 temporal T14": (2, 14, 9216, 5), "L0 temporal T25": (2, 25, 9216, 5), "L1 temporal T14": (2, 14, 2304, 10)}.items():
    qkv = torch.randn(B * T * S, 3 * H * 64, device=dev).half()
    ms = timeit(lambda: ops.temporal_attention(qkv, B, T, S, H))
    gb = B * T * S * H * 64 * 2 * 4 / 1e9
    print(f"{name:16s}: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s", flush=True)
for C, rows in [(320, 28 * 9216), (640, 28 * 9216), (640, 28 * 2304), (1280, 28 * 2304), (1280, 28 * 576), (2560, 28 * 576)]:
    x = torch.randn(rows, C, device=dev)
    g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
    for name, fn in {"GN spatial(28)": lambda: ops.group_norm(x, g, b, 28, 1e-6, True), "GN temporal(2)": lambda: ops.group_norm(x, g, b, 2, 1e-6, True),
                     "LN": lambda: ops.layer_norm(x, g, b)}.items():
        ms = timeit(fn)
        byts = rows * C * (10 if name.startswith("GN") else 6)
        print(f"{name:16s} rows={rows} C={C}: {ms:8.3f} ms  {byts / ms / 1e6:8.1f} GB/s", flush=True)
