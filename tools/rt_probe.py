import sys, os
sys.path.insert(0, "/root/repo")
import torch
from evoworld_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for (M, K, N) in [(258048, 320, 320), (258048, 1280, 320), (64512, 640, 640), (16128, 1280, 1280), (40000, 320, 320)]:
    a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) / K ** 0.5).half()
    b = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev)
    want = a.float() @ w.float().T + b + r
    for trial in range(3):
        got = ops.gemm_f16(a, w, bias=b, res1=r, out_dtype=torch.float32)
        buf = r.clone(); ops.gemm_f16(a, w, bias=b, res1=buf, out=buf)
        bad = (got - want).abs() > 1e-2
        print(f"M={M} K={K} N={N} trial {trial}: separate {rel(got, want):.2e} in-place {rel(buf, want):.2e} bad={int(bad.sum())}"
              + (f" first bad rows {torch.nonzero(bad)[:4].tolist()}" if bad.any() else ""), flush=True)
# conv 3x3 with residual
import torch.nn.functional as F
B, T, Y, X, C = 1, 28, 72, 128, 320
x = torch.randn(B * T, C, Y, X, device=dev).half(); w = (torch.randn(C, C, 3, 3, device=dev) / (9 * C) ** 0.5).half()
r = torch.randn(B * T * Y * X, C, device=dev)
want = F.conv2d(x.float(), w.float(), padding=1).permute(0, 2, 3, 1).reshape(-1, C) + r
a = x.permute(0, 2, 3, 1).reshape(B, T, Y, X, C).contiguous(); wk = w.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
for trial in range(2):
    got = ops.gemm_f16(a, wk, taps=ops.CONV3x3_TAPS, res1=r, out_dtype=torch.float32)
    print(f"conv3x3 +res trial {trial}: {rel(got, want):.2e}", flush=True)
