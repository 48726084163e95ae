"""Micro-benchmark of the attention kernels on the UNet's shapes (T=14, 576x1024)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoworld_b200 import ops

dev = torch.device("cuda:0")
for name, (F_, S, H) in {"L0 spatial": (28, 9216, 5), "L1 spatial": (28, 2304, 10), "L2 spatial": (28, 576, 20), "mid": (28, 144, 20)}.items():
    qkv = torch.randn(F_ * S, 3 * H * 64, device=dev).half()
    for _ in range(2):
        ops.spatial_attention(qkv, F_, S, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        ops.spatial_attention(qkv, F_, S, H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:12s} F={F_} S={S} H={H}: {ms:8.3f} ms  {4.0 * F_ * H * S * S * 64 / ms / 1e9:8.1f} TFLOP/s", flush=True)
for name, (B, T, S, H) in {"L0 temporal T14": (2, 14, 9216, 5), "L0 temporal T25": (2, 25, 9216, 5), "L1 temporal T14": (2, 14, 2304, 10)}.items():
    qkv = torch.randn(B * T * S, 3 * H * 64, device=dev).half()
    for _ in range(2):
        ops.temporal_attention(qkv, B, T, S, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        ops.temporal_attention(qkv, B, T, S, H)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = B * T * S * H * 64 * 2 * 4 / 1e9
    print(f"{name:16s}: {ms:8.3f} ms  {gb / ms * 1e3:8.1f} GB/s", flush=True)
x = torch.randn(28 * 9216, 320, device=dev)
g = torch.ones(320, device=dev); b = torch.zeros(320, device=dev)
for name, fn in {"GN L0 (28 inst)": lambda: ops.group_norm(x, g, b, 28, 1e-6, True), "GN L0 temporal (2 inst)": lambda: ops.group_norm(x, g, b, 2, 1e-6, True),
                 "LN L0": lambda: ops.layer_norm(x, g, b)}.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:26s}: {e0.elapsed_time(e1) / 5:8.3f} ms", flush=True)
