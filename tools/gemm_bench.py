"""Micro-benchmark of the tcgen05 GEMM on the UNet's dominant shapes (T=14, 576x1024)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evoworld_b200 import ops, _lib
L = _lib.lib()
MODES = [(1, "cluster"), (0, "plain")]

dev = torch.device("cuda:0")
torch.manual_seed(0)
res = []
BF = 28
shapes = [
    ("L0 linear 320->320", dict(M=BF * 9216, K=320, N=320)),
    ("L0 qkv 320->960", dict(M=BF * 9216, K=320, N=960)),
    ("L0 geglu 320->2560", dict(M=BF * 9216, K=320, N=2560, geglu=True)),
    ("L0 ff2 1280->320", dict(M=BF * 9216, K=1280, N=320)),
    ("L0 to_out 320->320 +res f32", dict(M=BF * 9216, K=320, N=320, res=True)),
    ("L0 ff2 1280->320 +res f32", dict(M=BF * 9216, K=1280, N=320, res=True)),
    ("L1 to_out 640->640 +res f32", dict(M=BF * 2304, K=640, N=640, res=True)),
    ("L1 geglu 640->5120", dict(M=BF * 2304, K=640, N=5120, geglu=True)),
    ("L1 ff2 2560->640", dict(M=BF * 2304, K=2560, N=640)),
    ("L2 geglu 1280->10240", dict(M=BF * 576, K=1280, N=10240, geglu=True)),
    ("L2 ff2 5120->1280", dict(M=BF * 576, K=5120, N=1280)),
]
for name, s in shapes:
    a = torch.randn(s["M"], s["K"], device=dev).half()
    w = (torch.randn(s["N"], s["K"], device=dev) / s["K"] ** 0.5).half()
    kw = dict(geglu=True) if s.get("geglu") else {}
    if s.get("res"):
        kw = dict(res1=torch.randn(s["M"], s["N"], device=dev), out_dtype=torch.float32, bias=torch.randn(s["N"], device=dev))
    for bn in ([0] if "--sweep" not in sys.argv else [128, 160, 256]):
        if s.get("geglu") and bn % 32:
            continue
        if bn and s["N"] % bn:
            continue
        for mode, mname in MODES:
            L.evw_set_gemm_cluster(mode)
            for _ in range(3):
                out = ops.gemm_f16(a, w, block_n=bn, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(10):
                ops.gemm_f16(a, w, block_n=bn, out=out, **kw)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            tf = 2.0 * s["M"] * s["K"] * s["N"] / ms / 1e9
            res.append(dict(name=name, block_n=bn, mode=mname, ms=ms, tflops=tf))
            print(f"{name:28s} bn={bn:3d} {mname:8s} {ms:8.3f} ms {tf:8.1f} TFLOP/s", flush=True)
# conv 3x3 L0 320->320
import torch.nn.functional as F
for name, (Y, X, C, N) in {"L0 conv3x3 320->320": (72, 128, 320, 320), "L1 conv3x3 640->640": (36, 64, 640, 640),
                           "L2 conv3x3 1280->1280": (18, 32, 1280, 1280), "L3 conv3x3 1280->1280": (9, 16, 1280, 1280)}.items():
    a = torch.randn(2, 14, Y, X, C, device=dev).half()
    w = (torch.randn(N, 9 * C, device=dev) / (9 * C) ** 0.5).half()
    for mode, mname in MODES:
        L.evw_set_gemm_cluster(mode)
        for _ in range(3):
            out = ops.gemm_f16(a, w, taps=ops.CONV3x3_TAPS)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            ops.gemm_f16(a, w, taps=ops.CONV3x3_TAPS, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.0 * 28 * Y * X * 9 * C * N / ms / 1e9
        res.append(dict(name=name, mode=mname, ms=ms, tflops=tf))
        print(f"{name:28s}        {mname:8s} {ms:8.3f} ms {tf:8.1f} TFLOP/s", flush=True)
L.evw_set_gemm_cluster(-1)
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "gemm_bench.json"), "w"))
