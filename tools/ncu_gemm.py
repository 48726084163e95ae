"""Small driver for ncu captures of single kernels: python tools/ncu_gemm.py <case>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoworld_b200 import ops
dev = torch.device("cuda:0")
case = sys.argv[1] if len(sys.argv) > 1 else "geglu"
M = 28 * 9216
if case == "geglu":
    a = torch.randn(M, 320, device=dev).half(); w = torch.randn(2560, 320, device=dev).half() * 0.05
    f = lambda: ops.gemm_f16(a, w, geglu=True)
elif case == "linear":
    a = torch.randn(M, 320, device=dev).half(); w = torch.randn(320, 320, device=dev).half() * 0.05
    f = lambda: ops.gemm_f16(a, w)
elif case == "qkv":
    a = torch.randn(M, 320, device=dev).half(); w = torch.randn(960, 320, device=dev).half() * 0.05
    f = lambda: ops.gemm_f16(a, w)
elif case == "to_out":
    a = torch.randn(M, 320, device=dev).half(); w = torch.randn(320, 320, device=dev).half() * 0.05
    r = torch.randn(M, 320, device=dev); b = torch.randn(320, device=dev)
    f = lambda: ops.gemm_f16(a, w, bias=b, res1=r, out_dtype=torch.float32)
elif case == "ff2":
    a = torch.randn(M, 1280, device=dev).half(); w = torch.randn(320, 1280, device=dev).half() * 0.03
    r = torch.randn(M, 320, device=dev)
    f = lambda: ops.gemm_f16(a, w, res1=r, out_dtype=torch.float32)
elif case == "conv":
    a = torch.randn(2, 14, 72, 128, 320, device=dev).half(); w = torch.randn(320, 9 * 320, device=dev).half() * 0.02
    f = lambda: ops.gemm_f16(a, w, taps=ops.CONV3x3_TAPS)
elif case in ("conv_stats", "conv_rv"):  # level-0 conv1: 3x3 conv + bias + time-embedding row vector, fp32 out (+ GroupNorm sums)
    a = torch.randn(1, 28, 72, 128, 320, device=dev).half(); w = torch.randn(320, 9 * 320, device=dev).half() * 0.02
    b = torch.randn(320, device=dev); rv = torch.randn(28, 320, device=dev)
    st = torch.empty(28, 32, 2, dtype=torch.float64, device=dev) if case == "conv_stats" else None
    f = lambda: ops.gemm_f16(a, w, taps=ops.CONV3x3_TAPS, bias=b, rowvec=rv, rv_div=9216, rv_mod=28, out_dtype=torch.float32,
                             gn_stats=st, gn_rows_per_inst=9216 if st is not None else 0)
elif case == "attn":
    qkv = torch.randn(28 * 9216, 960, device=dev).half()
    f = lambda: ops.spatial_attention(qkv, 28, 9216, 5)
elif case == "gn":
    x = torch.randn(M, 320, device=dev); g = torch.ones(320, device=dev); b = torch.zeros(320, device=dev)
    f = lambda: ops.group_norm(x, g, b, 28, 1e-6, True)
for _ in range(4):
    f()
torch.cuda.synchronize()
