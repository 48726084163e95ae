"""tcgen05 implicit-GEMM (csrc/tc_gemm.cu) against a plain PyTorch fp32 reference of the same op
on the same fp16-rounded operands.  Tolerance: fp32 accumulation of fp16 products -> relative L2
<= 2e-3 for fp16 outputs (one rounding of the result), <= 2e-5 for fp32 outputs."""
import pytest
import torch
import torch.nn.functional as F

from evoworld_b200 import ops

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(autouse=True, params=[1, 0], ids=["cluster", "plain"])
def _seed(request, cuda_device, built_lib):
    """Every case runs twice: with the 2-CTA cluster / weight-multicast launch (the default) and with the plain launch."""
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    built_lib.evw_set_gemm_cluster(request.param)
    yield
    built_lib.evw_set_gemm_cluster(-1)


@pytest.mark.parametrize("M,K,N", [(128, 64, 64), (1000, 320, 320), (258, 640, 1280), (4096, 1280, 160), (2, 320, 1280),
                                   (777, 64, 16), (300, 128, 960), (512, 2560, 256)])
@pytest.mark.parametrize("out_dtype", [torch.float16, torch.float32])
def test_linear(M, K, N, out_dtype, cuda_device):
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    got = ops.gemm_f16(a, w, bias=b, out_dtype=out_dtype)
    want = a.float() @ w.float().T + b
    assert got.shape == (M, N) and got.dtype == out_dtype
    assert rel_l2(got, want) < (2e-3 if out_dtype == torch.float16 else 2e-5)


@pytest.mark.parametrize("block_n", [16, 32, 64, 128, 160, 256])
def test_linear_block_n(block_n, cuda_device):
    M, K, N = 1500, 320, 640
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).half()
    got = ops.gemm_f16(a, w, out_dtype=torch.float32, block_n=block_n)
    assert rel_l2(got, a.float() @ w.float().T) < 2e-5


def test_linear_full_epilogue(cuda_device):
    M, K, N, T, S = 2 * 3 * 50, 320, 320, 3, 50
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    rv = torch.randn(T, N, device=cuda_device)
    r1 = torch.randn(M, N, device=cuda_device)
    r2 = torch.randn(M, N, device=cuda_device)
    got = ops.gemm_f16(a, w, bias=b, rowvec=rv, rv_div=S, rv_mod=T, res1=r1, s1=0.3, res2=r2, s2=0.7, s0=0.3,
                       out_dtype=torch.float32)
    t_idx = (torch.arange(M, device=cuda_device) // S) % T
    want = 0.3 * (a.float() @ w.float().T + b) + rv[t_idx] + 0.3 * r1 + 0.7 * r2
    assert rel_l2(got, want) < 2e-5
    # fp16 residual, in-place accumulate into the residual buffer
    r1h = r1.half()
    want2 = a.float() @ w.float().T + b + r1h.float()
    buf = r1h.clone()
    ops.gemm_f16(a, w, bias=b, res1=buf, out=buf)
    assert rel_l2(buf, want2) < 2e-3


@pytest.mark.parametrize("M,K,F_", [(640, 320, 1280), (100, 640, 2560)])
def test_geglu(M, K, F_, cuda_device):
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(2 * F_, K, device=cuda_device) / K ** 0.5).half()
    b = torch.randn(2 * F_, device=cuda_device)
    wi, bi = ops.geglu_interleave(w, b)
    got = ops.gemm_f16(a, wi, bias=bi, geglu=True, out_dtype=torch.float16)
    h = a.float() @ w.float().T + b
    want = h[:, :F_] * F.gelu(h[:, F_:])
    assert got.shape == (M, F_)
    assert rel_l2(got, want) < 2e-3


@pytest.mark.parametrize("B,T,Y,X,C,N", [(1, 1, 8, 16, 64, 64), (2, 3, 9, 16, 128, 64), (1, 2, 18, 32, 64, 128),
                                         (1, 2, 36, 64, 64, 64), (1, 1, 72, 128, 64, 32), (1, 2, 5, 7, 64, 64),
                                         (1, 1, 20, 72, 64, 64)])
def test_conv3x3(B, T, Y, X, C, N, cuda_device):
    x = torch.randn(B * T, C, Y, X, device=cuda_device).half()
    w = (torch.randn(N, C, 3, 3, device=cuda_device) / (9 * C) ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    want = F.conv2d(x.float(), w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    a = x.permute(0, 2, 3, 1).reshape(B, T, Y, X, C).contiguous()
    wk = w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()  # [N, ky, kx, C]
    got = ops.gemm_f16(a, wk, taps=ops.CONV3x3_TAPS, bias=b, out_dtype=torch.float32)
    assert rel_l2(got, want) < 2e-5


def test_conv3x3_with_shortcut_and_temb(cuda_device):
    B, T, Y, X, C, C1, N = 2, 2, 9, 16, 128, 192, 64
    x = torch.randn(B * T, C, Y, X, device=cuda_device).half()
    xs = torch.randn(B * T, C1, Y, X, device=cuda_device).half()
    w = (torch.randn(N, C, 3, 3, device=cuda_device) / (9 * C) ** 0.5).half()
    ws = (torch.randn(N, C1, 1, 1, device=cuda_device) / C1 ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    temb = torch.randn(B * T, N, device=cuda_device)
    want = F.conv2d(x.float(), w.float(), b, padding=1) + F.conv2d(xs.float(), ws.float()) + temb[:, :, None, None]
    want = want.permute(0, 2, 3, 1).reshape(-1, N)
    a0 = x.permute(0, 2, 3, 1).reshape(B, T, Y, X, C).contiguous()
    a1 = xs.permute(0, 2, 3, 1).reshape(B, T, Y, X, C1).contiguous()
    wk = torch.cat([w.permute(0, 2, 3, 1).reshape(N, 9 * C), ws.reshape(N, C1)], dim=1).contiguous()
    got = ops.gemm_f16(a0, wk, taps=ops.CONV3x3_TAPS + [(0, 0, 0, 1)], a1=a1, bias=b, rowvec=temb, rv_div=Y * X,
                       rv_mod=B * T, out_dtype=torch.float32)
    assert rel_l2(got, want) < 2e-5


def test_temporal_conv(cuda_device):
    B, T, S, C, N = 2, 5, 200, 128, 64
    x = torch.randn(B, C, T, S, 1, device=cuda_device).half()
    w = (torch.randn(N, C, 3, 1, 1, device=cuda_device) / (3 * C) ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    want = F.conv3d(x.float(), w.float(), b, padding=(1, 0, 0))  # [B,N,T,S,1]
    want = want[..., 0].permute(0, 2, 3, 1).reshape(-1, N)
    a = x[..., 0].permute(0, 2, 3, 1).reshape(B, T, 1, S, C).contiguous()
    wk = w[..., 0, 0].permute(0, 2, 1).reshape(N, 3 * C).contiguous()  # [N, kt, C]
    got = ops.gemm_f16(a, wk, taps=ops.TEMPORAL_TAPS, bias=b, out_dtype=torch.float32)
    assert rel_l2(got, want) < 2e-5


def test_large_persistent(cuda_device):
    """More tiles than SMs with a long K loop: exercises ring wrap-around and both accumulator stages."""
    M, K, N = 148 * 128 * 3 + 77, 1280, 320
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).half()
    got = ops.gemm_f16(a, w, out_dtype=torch.float16)
    want = (a @ w.T).float()
    assert rel_l2(got, want) < 3e-3


def _hi_lo(x):
    hi = x.half()
    return hi, (x - hi.float()).half()


@pytest.mark.parametrize("M,K,N", [(1000, 320, 320), (4096, 640, 640), (258, 1280, 1280)])
def test_split_precision_linear(M, K, N, cuda_device):
    """proj_in / proj_out in split precision: [a_hi | a_lo | a_hi] x [W_hi | W_hi | W_lo] keeps ~22 bits of the
    product (error 100x below the plain fp16-operand GEMM) — tools/precision_sim.py's "split" mode on the device."""
    a = torch.randn(M, K, device=cuda_device)
    w = torch.randn(N, K, device=cuda_device) / K ** 0.5
    b = torch.randn(N, device=cuda_device)
    want = a.double() @ w.double().T + b.double()
    a_hi, a_lo = _hi_lo(a)
    w_hi, w_lo = _hi_lo(w)
    wk = torch.cat([w_hi, w_hi, w_lo], dim=1).contiguous()
    got = ops.gemm_f16(a_hi, wk, taps=[(0, 0, 0, 0), (0, 0, 0, 1), (0, 0, 0, 0)], a1=a_lo, bias=b, out_dtype=torch.float32)
    plain = ops.gemm_f16(a_hi, w_hi, bias=b, out_dtype=torch.float32)
    e_split, e_plain = rel_l2(got, want), rel_l2(plain, want)
    print(f"split-precision linear {M}x{N}x{K}: rel L2 {e_split:.2e} (plain fp16 operands {e_plain:.2e})")
    assert e_split < 5e-6 and e_plain > 20 * e_split


def test_epilogue_tail_output(cuda_device):
    """out_lo: the fp16 tail written next to an fp16 output restores the fp32 value (head + tail)."""
    M, K, N = 700, 320, 320
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).half()
    r1 = torch.randn(M, N, device=cuda_device)
    want = ops.gemm_f16(a, w, res1=r1, s1=0.5, out_dtype=torch.float32)
    lo = torch.full((M, N), 7.0, dtype=torch.float16, device=cuda_device)
    hi = ops.gemm_f16(a, w, res1=r1, s1=0.5, out_dtype=torch.float16, out_lo=lo)
    assert torch.equal(hi, want.half())
    assert torch.equal(lo, (want - want.half().float()).half())
    assert rel_l2(hi.float() + lo.float(), want) < 2e-7


def test_conv3x3_split_taps(cuda_device):
    """conv_out in split precision: 9 taps on the head + 9 on the tail of the input, weight tail in extra output rows."""
    B, T, Y, X, C, Co = 1, 2, 16, 32, 64, 4
    x = torch.randn(B * T, C, Y, X, device=cuda_device)
    w = torch.randn(Co, C, 3, 3, device=cuda_device) / (9 * C) ** 0.5
    want = F.conv2d(x.double(), w.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, Co)
    xl = x.permute(0, 2, 3, 1).reshape(B, T, Y, X, C).contiguous()
    x_hi, x_lo = _hi_lo(xl)
    w_hi, w_lo = _hi_lo(w.permute(0, 2, 3, 1).reshape(Co, 9 * C))
    wk = torch.cat([torch.cat([w_hi, w_hi], 1), torch.cat([w_lo, torch.zeros_like(w_lo)], 1)], 0)
    wk = F.pad(wk, (0, 0, 0, 16 - 2 * Co)).contiguous()
    taps = ops.CONV3x3_TAPS + [(dx, dy, dt, 1) for dx, dy, dt, _ in ops.CONV3x3_TAPS]
    y = ops.gemm_f16(x_hi, wk, taps=taps, a1=x_lo, out_dtype=torch.float32)
    got = y[:, :Co] + y[:, Co:2 * Co]
    assert rel_l2(got, want) < 5e-6


@pytest.mark.parametrize("M,K,N", [(258048, 320, 320), (64512, 640, 640), (40000, 1280, 1280)])
def test_residual_ring_at_scale(M, K, N, cuda_device):
    """fp32 residual staged through the TMA-filled shared-memory ring, many tiles per CTA, separate and in place: the first
    version released a box before its loads had landed and corrupted a few thousand 16-byte chunks per 80 M outputs — only
    at sizes where the ring wraps many times (tools/rt_probe.py)."""
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    r = torch.randn(M, N, device=cuda_device)
    want = a.float() @ w.float().T + b + r
    for _ in range(2):
        got = ops.gemm_f16(a, w, bias=b, res1=r, out_dtype=torch.float32)
        assert int(((got - want).abs() > 1e-2).sum()) == 0 and rel_l2(got, want) < 2e-5
        buf = r.clone()
        ops.gemm_f16(a, w, bias=b, res1=buf, out=buf)
        assert int(((buf - want).abs() > 1e-2).sum()) == 0 and rel_l2(buf, want) < 2e-5
