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
    """Every case runs twice: forced CTA pairs (tcgen05.mma.cta_group::2, each CTA holding half of the weight tile) and forced
    single CTAs.  The library default picks per GEMM (pairs where K_total >= 1024, evw_set_gemm_cluster(-1)); both modes are
    bit-identical, so covering the two forced modes covers the default."""
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


def _check_gn_sums(out, stats, insts, N, tol=2e-6):
    """stats [insts, 32, 2] against float64 sums of the stored output."""
    o = out.double().reshape(insts, -1, 32, N // 32)
    want = torch.stack([o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))], dim=-1)  # [insts, 32, 2]
    scale = want[..., 1].sqrt().unsqueeze(-1) * (o.shape[1] * o.shape[3]) ** 0.5 + 1e-30  # |sum| <= sqrt(n * sum of squares)
    err_sum = ((stats[..., 0] - want[..., 0]).abs() / scale[..., 0]).max()
    err_sq = ((stats[..., 1] - want[..., 1]).abs() / (want[..., 1] + 1e-30)).max()
    assert float(err_sum) < tol and float(err_sq) < tol, (float(err_sum), float(err_sq))


@pytest.mark.parametrize("T,Y,X,C,N,per_frame", [(3, 24, 40, 64, 320, True), (4, 9, 16, 64, 640, False), (2, 36, 64, 64, 1280, True),
                                                 (2, 72, 128, 64, 256, False), (28, 18, 32, 128, 320, False)])
@pytest.mark.parametrize("out_dtype", [torch.float16, torch.float32])
def test_conv3x3_group_norm_statistics(T, Y, X, C, N, per_frame, out_dtype, cuda_device):
    """The epilogue's GroupNorm sums (evw_gemm_f16_gn) of a 3x3 convolution + bias + per-frame row vector: partial tiles
    (rows outside the frame), groups straddling column steps and N tiles, instance = frame or = all frames."""
    x = torch.randn(1, T, Y, X, C, device=cuda_device).half()
    wk = (torch.randn(N, 9 * C, device=cuda_device) / (9 * C) ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    temb = torch.randn(T, N, device=cuda_device)
    plain = ops.gemm_f16(x, wk, taps=ops.CONV3x3_TAPS, bias=b, rowvec=temb, rv_div=Y * X, rv_mod=T, out_dtype=out_dtype)
    insts = T if per_frame else 1
    stats = torch.full((insts, 32, 2), 7.0, dtype=torch.float64, device=cuda_device)  # the call clears it
    got = ops.gemm_f16(x, wk, taps=ops.CONV3x3_TAPS, bias=b, rowvec=temb, rv_div=Y * X, rv_mod=T, out_dtype=out_dtype,
                       gn_stats=stats, gn_rows_per_inst=(Y * X if per_frame else T * Y * X))
    assert torch.equal(got, plain)
    # the sums are taken before the fp16 rounding of the store: compare with the fp32 result
    ref = got if out_dtype == torch.float32 else ops.gemm_f16(x, wk, taps=ops.CONV3x3_TAPS, bias=b, rowvec=temb, rv_div=Y * X,
                                                               rv_mod=T, out_dtype=torch.float32)
    _check_gn_sums(ref, stats, insts, N)


@pytest.mark.parametrize("M,K,N,rpi", [(9216 * 2, 320, 320, 9216), (2304 * 3, 640, 640, 2304), (1280, 1280, 1280, 640),
                                       (129024, 320, 320, 9216)])
def test_linear_residual_group_norm_statistics(M, K, N, rpi, cuda_device):
    """Token-row linear with the fp32 residual staged by TMA (proj_out / the temporal conv2 shape): instance = rpi rows."""
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    r = torch.randn(M, N, device=cuda_device)
    stats = torch.empty((M // rpi, 32, 2), dtype=torch.float64, device=cuda_device)
    got = ops.gemm_f16(a, w, bias=b, res1=r, out_dtype=torch.float32, gn_stats=stats, gn_rows_per_inst=rpi)
    assert torch.equal(got, ops.gemm_f16(a, w, bias=b, res1=r, out_dtype=torch.float32))
    _check_gn_sums(got, stats, M // rpi, N)
    # no residual, no row vector
    got = ops.gemm_f16(a, w, bias=b, out_dtype=torch.float32, gn_stats=stats, gn_rows_per_inst=rpi)
    _check_gn_sums(got, stats, M // rpi, N)


def test_group_norm_statistics_refused(cuda_device):
    """Epilogues / geometries the statistics are not built for fail loudly instead of returning wrong sums."""
    a = torch.randn(1000, 320, device=cuda_device).half()
    w = (torch.randn(320, 320, device=cuda_device) / 320 ** 0.5).half()
    stats = torch.empty((10, 32, 2), dtype=torch.float64, device=cuda_device)
    with pytest.raises(RuntimeError):  # 100 rows per instance: a 128-row tile would straddle instances
        ops.gemm_f16(a, w, out_dtype=torch.float32, gn_stats=stats, gn_rows_per_inst=100)
    a = torch.randn(1024, 320, device=cuda_device).half()
    r16 = torch.randn(1024, 320, device=cuda_device).half()
    stats = torch.empty((1, 32, 2), dtype=torch.float64, device=cuda_device)
    with pytest.raises(RuntimeError):  # fp16 residual: not staged by TMA
        ops.gemm_f16(a, w, res1=r16, out_dtype=torch.float32, gn_stats=stats, gn_rows_per_inst=1024)
    w96 = (torch.randn(96, 320, device=cuda_device) / 320 ** 0.5).half()
    with pytest.raises(RuntimeError):  # 3 channels per group: more groups per 16-column step than the epilogue has slots for
        ops.gemm_f16(a, w96, out_dtype=torch.float32, gn_stats=stats, gn_rows_per_inst=1024)


def test_pair_and_single_cta_modes_are_bit_identical(cuda_device, built_lib):
    """cta_group::2 pairs (M = 256 over two SMs) and single CTAs accumulate every output element over the same K order:
    identical bits for a convolution with GroupNorm sums, a residual linear and a GEGLU projection."""
    x = torch.randn(1, 6, 24, 40, 128, device=cuda_device).half()
    wk = (torch.randn(320, 9 * 128, device=cuda_device) / (9 * 128) ** 0.5).half()
    a = torch.randn(3000, 640, device=cuda_device).half()
    w = (torch.randn(640, 640, device=cuda_device) / 640 ** 0.5).half()
    wg = (torch.randn(2560, 640, device=cuda_device) / 640 ** 0.5).half()
    r = torch.randn(3000, 640, device=cuda_device)
    outs = []
    for mode in (1, 0):
        built_lib.evw_set_gemm_cluster(mode)
        st = torch.empty(6, 32, 2, dtype=torch.float64, device=cuda_device)
        c = ops.gemm_f16(x, wk, taps=ops.CONV3x3_TAPS, out_dtype=torch.float32, gn_stats=st, gn_rows_per_inst=24 * 40)
        outs.append((c, st.clone(), ops.gemm_f16(a, w, res1=r, out_dtype=torch.float32), ops.gemm_f16(a, wg, geglu=True)))
    for u, v in zip(*outs):
        if u.dtype == torch.float64:  # the per-tile sums are identical; their fp64 atomics land in any order
            assert torch.allclose(u, v, rtol=1e-12, atol=0)
        else:
            assert torch.equal(u, v)


def test_tma_store_and_direct_store_are_bit_identical(cuda_device, built_lib):
    """The TMA-store epilogue (per-warp swizzled slabs, stores clipped by the output's tensor map) against per-thread global
    stores: partial tiles in x, y and N, 16-wide image rows (a warp's 32 rows = two image rows), fp16 / fp32 outputs, residual,
    GroupNorm sums, GEGLU, in-place residual update."""
    torch.manual_seed(1)
    cases = []
    x = torch.randn(1, 3, 24, 40, 64, device=cuda_device).half()      # bx = 32, partial tiles in x and y
    w = (torch.randn(328, 9 * 64, device=cuda_device) / 24).half()    # N = 328: 8-column tail
    cases.append(lambda: ops.gemm_f16(x, w, taps=ops.CONV3x3_TAPS, out_dtype=torch.float32))
    cases.append(lambda: ops.gemm_f16(x, w[:320], taps=ops.CONV3x3_TAPS, out_dtype=torch.float16))
    x16 = torch.randn(2, 2, 9, 16, 128, device=cuda_device).half()    # X = 16 = bx: tile spans the image width
    w16 = (torch.randn(64, 9 * 128, device=cuda_device) / 34).half()
    cases.append(lambda: ops.gemm_f16(x16, w16, taps=ops.CONV3x3_TAPS, out_dtype=torch.float32))
    a = torch.randn(1000, 320, device=cuda_device).half()
    wl = (torch.randn(320, 320, device=cuda_device) / 18).half()
    r = torch.randn(1000, 320, device=cuda_device)
    b = torch.randn(320, device=cuda_device)
    cases.append(lambda: ops.gemm_f16(a, wl, bias=b, res1=r, out_dtype=torch.float32))
    cases.append(lambda: ops.gemm_f16(a, wl, bias=b, out_dtype=torch.float16))
    wg = (torch.randn(2560, 320, device=cuda_device) / 18).half()
    cases.append(lambda: ops.gemm_f16(a, wg, geglu=True))

    def in_place():
        buf = r.clone()
        ops.gemm_f16(a, wl, bias=b, res1=buf, out=buf)
        return buf

    cases.append(in_place)

    def with_sums():
        st = torch.empty(3, 32, 2, dtype=torch.float64, device=cuda_device)
        out = ops.gemm_f16(x, w[:320], taps=ops.CONV3x3_TAPS, out_dtype=torch.float32, gn_stats=st, gn_rows_per_inst=24 * 40)
        return torch.cat([out.flatten(), st.flatten().float()])

    cases.append(with_sums)
    try:
        outs = []
        for mode in (1, 0):
            built_lib.evw_set_gemm_store_tma(mode)
            outs.append([f() for f in cases])
    finally:
        built_lib.evw_set_gemm_store_tma(-1)
    for i, (u, v) in enumerate(zip(*outs)):
        assert torch.equal(u, v), f"case {i}"


@pytest.mark.parametrize("Fr,h,w,C,N", [(2, 9, 16, 64, 64), (3, 18, 32, 128, 320), (1, 36, 64, 64, 128), (2, 5, 40, 64, 80)])
def test_fused_upsample_conv(Fr, h, w, C, N, cuda_device, built_lib):
    """diffusers Upsample2D (nearest x2, then Conv2d 3x3 padding 1) as four 2x2 phase convolutions of the low-resolution input
    stored through strided TMA maps (evw_upconv2x_f16) against the literal torch form on the same fp16-rounded operands; the
    phase weights are fp32 sums of fp16-rounded taps here so that only the accumulation order differs."""
    x = torch.randn(Fr, C, h, w, device=cuda_device).half()
    wt = (torch.randn(N, C, 3, 3, device=cuda_device) / (9 * C) ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    want = F.conv2d(F.interpolate(x.float(), scale_factor=2.0, mode="nearest"), wt.float(), b, padding=1).permute(0, 2, 3, 1)
    w4 = ops.upconv_weights(wt)
    got = ops.upconv2x(x.permute(0, 2, 3, 1).contiguous(), w4, b)
    assert got.shape == (Fr, 2 * h, 2 * w, N)
    # the summed phase weights are rounded to fp16 once more: 2.8e-4 relative per weight, averaged over K
    assert rel_l2(got, want) < 3e-4


@pytest.mark.parametrize("M,K,N", [(1000, 320, 1280), (4164, 1024, 4096), (3, 2048, 1024), (777, 64, 16), (300, 384, 1536)])
@pytest.mark.parametrize("out_dtype", [torch.float16, torch.float32])
def test_linear_gelu_epilogue(M, K, N, out_dtype, cuda_device, built_lib):
    """fc1 of the ViT-style MLPs (VGGT / CLIP): exact GELU of acc + bias in the epilogue (evw_gemm_f16 geglu = 2), single CTAs
    and CTA pairs (K >= 1024), direct and TMA stores bit-identical; operand combinations it is not built for are refused."""
    a = torch.randn(M, K, device=cuda_device).half()
    w = (torch.randn(N, K, device=cuda_device) / K ** 0.5).half()
    b = torch.randn(N, device=cuda_device)
    want = F.gelu(a.float() @ w.float().T + b)
    try:
        outs = []
        for mode in (1, 0):
            built_lib.evw_set_gemm_store_tma(mode)
            outs.append(ops.gemm_f16(a, w, bias=b, act="gelu", out_dtype=out_dtype))
    finally:
        built_lib.evw_set_gemm_store_tma(-1)
    assert torch.equal(outs[0], outs[1])
    assert outs[0].dtype == out_dtype and rel_l2(outs[0], want) < (2e-3 if out_dtype == torch.float16 else 2e-5)
    with pytest.raises(RuntimeError):
        ops.gemm_f16(a, w, bias=b, act="gelu", res1=torch.zeros(M, N, device=cuda_device), out_dtype=torch.float32)
    with pytest.raises(ValueError):
        ops.gemm_f16(a, w, bias=b, act="relu")
