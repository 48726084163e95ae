"""Attention / normalisation kernels of the UNet path against plain PyTorch fp32 references on the
same fp16-rounded inputs.  Tolerances (relative L2): attention 3e-3 (fp16 P and fp16 output),
norms 1.5e-3 (fp16 output)."""
import pytest
import torch
import torch.nn.functional as F

from evoworld_b200 import ops

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(autouse=True)
def _seed(cuda_device, built_lib):
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False


def ref_attention(q, k, v):  # [..., L, 64] fp32
    s = (q @ k.transpose(-1, -2)) * 0.125
    return torch.softmax(s, dim=-1) @ v


@pytest.mark.parametrize("frames,S,heads", [(1, 128, 1), (2, 256, 2), (3, 144, 5), (2, 576, 3), (1, 2304, 2), (2, 100, 1),
                                            (1, 9216, 1), (1, 1, 1), (1, 129, 2)])
@pytest.mark.parametrize("scale", [1.0, 4.0])
def test_spatial_attention(frames, S, heads, scale, cuda_device):
    C = heads * 64
    qkv = (torch.randn(frames * S, 3 * C, device=cuda_device) * scale).half()
    got = ops.spatial_attention(qkv, frames, S, heads)
    x = qkv.float().view(frames, S, 3, heads, 64).permute(2, 0, 3, 1, 4)  # [3, F, H, S, 64]
    want = ref_attention(x[0], x[1], x[2]).permute(0, 2, 1, 3).reshape(frames * S, C)
    assert got.shape == want.shape
    assert torch.isfinite(got.float()).all()
    assert rel_l2(got, want) < 3e-3


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("frames,S,heads", [(2, 576, 3), (1, 129, 2), (1, 2304, 2), (3, 144, 5), (1, 300, 1), (1, 1, 1)])
def test_spatial_attention_variants(variant, frames, S, heads, cuda_device):
    """Every variant behind evw_set_attention_variant (v8 default / partial FMA-pipe exponentials / staggered groups,
    v7 one thread per row) against the same fp32 reference."""
    from evoworld_b200 import _lib

    C = heads * 64
    qkv = (torch.randn(frames * S, 3 * C, device=cuda_device) * 2.0).half()
    x = qkv.float().view(frames, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    want = ref_attention(x[0], x[1], x[2]).permute(0, 2, 1, 3).reshape(frames * S, C)
    L = _lib.lib()
    L.evw_set_attention_variant(variant)
    try:
        got = ops.spatial_attention(qkv, frames, S, heads)
        torch.cuda.synchronize()
    finally:
        L.evw_set_attention_variant(-2)
    assert torch.isfinite(got.float()).all()
    assert rel_l2(got, want) < 3e-3


@pytest.mark.parametrize("B,T,S,heads", [(1, 14, 64, 1), (2, 14, 200, 5), (2, 25, 144, 2), (1, 1, 10, 1), (1, 3, 33, 2),
                                         (2, 16, 37, 1), (1, 32, 20, 1), (2, 17, 5, 3)])
def test_temporal_attention(B, T, S, heads, cuda_device):
    C = heads * 64
    qkv = torch.randn(B * T * S, 3 * C, device=cuda_device).half()
    got = ops.temporal_attention(qkv, B, T, S, heads)
    x = qkv.float().view(B, T, S, 3, heads, 64).permute(3, 0, 2, 4, 1, 5)  # [3, B, S, H, T, 64]
    want = ref_attention(x[0], x[1], x[2]).permute(0, 3, 1, 2, 4).reshape(B * T * S, C)  # [B,T,S,H,64]
    assert rel_l2(got, want) < 3e-3


@pytest.mark.parametrize("insts,rows,C0,C1,silu,src16", [(4, 144, 320, 0, True, False), (2, 1000, 640, 0, False, False),
                                                         (3, 77, 1280, 640, True, False), (2, 300, 640, 320, True, False),
                                                         (4, 144, 320, 0, True, True), (1, 5000, 1280, 1280, True, False)])
def test_group_norm(insts, rows, C0, C1, silu, src16, cuda_device):
    C = C0 + C1
    x0 = torch.randn(insts * rows, C0, device=cuda_device) * 2 + 0.5
    if src16:
        x0 = x0.half()
    x1 = torch.randn(insts * rows, C1, device=cuda_device) - 1.0 if C1 else None
    g, b = torch.randn(C, device=cuda_device), torch.randn(C, device=cuda_device)
    got, raw = ops.group_norm(x0, g, b, insts, 1e-6, silu, src1=x1, want_raw=True)
    x = torch.cat([x0.float()] + ([x1] if C1 else []), dim=1)
    xr = x.view(insts, rows, C).permute(0, 2, 1)  # [N, C, L]
    want = F.group_norm(xr, 32, g, b, eps=1e-6)
    if silu:
        want = F.silu(want)
    want = want.permute(0, 2, 1).reshape(insts * rows, C)
    assert rel_l2(got, want) < 1.5e-3
    assert rel_l2(raw, x) < 1e-3


@pytest.mark.parametrize("rows,C", [(1000, 320), (333, 640), (64, 1280), (5, 64), (17, 2560)])
def test_layer_norm(rows, C, cuda_device):
    x = torch.randn(rows, C, device=cuda_device) * 3 + 1
    g, b = torch.randn(C, device=cuda_device), torch.randn(C, device=cuda_device)
    got = ops.layer_norm(x, g, b)
    assert rel_l2(got, F.layer_norm(x, (C,), g, b, 1e-5)) < 1.5e-3
    T, S = 5, max(rows // 10, 1)
    rv = torch.randn(T, C, device=cuda_device)
    idx = (torch.arange(rows, device=cuda_device) // S) % T
    got = ops.layer_norm(x, g, b, rowvec=rv, rv_div=S, rv_mod=T)
    assert rel_l2(got, F.layer_norm(x + rv[idx], (C,), g, b, 1e-5)) < 1.5e-3


def test_group_norm_tail_output(cuda_device):
    """out_lo: fp16 tail of the normalised output (split-precision operand of proj_in / conv_out)."""
    insts, rows, C = 3, 200, 320
    x = torch.randn(insts * rows, C, device=cuda_device) * 2 + 0.5
    g, b = torch.randn(C, device=cuda_device), torch.randn(C, device=cuda_device)
    hi, lo = ops.group_norm(x, g, b, insts, 1e-6, True, want_lo=True)
    assert torch.equal(hi, ops.group_norm(x, g, b, insts, 1e-6, True))
    want = F.silu(F.group_norm(x.view(insts, rows, C).permute(0, 2, 1), 32, g, b, eps=1e-6)).permute(0, 2, 1).reshape(-1, C)
    e_hi, e_sum = rel_l2(hi, want), rel_l2(hi.float() + lo.float(), want)
    assert e_sum < 2e-5 and e_sum < 0.1 * e_hi  # limited by silu_fast / fp32 statistics, not by fp16 storage


def test_state_dict_views_are_copied(cuda_device, built_lib):
    """Parameters handed over as views at odd offsets of one flat buffer (16-byte misaligned) must not reach the kernels'
    vector loads: load_state_dict takes private, freshly allocated copies (round-2 smoke failure)."""
    from evoworld_b200.unet import UNetSpatioTemporalConditionModel

    cfg = dict(in_channels=18, block_out_channels=(64, 64, 64, 64), num_attention_heads=(1, 1, 1, 1), cross_attention_dim=64)
    a = UNetSpatioTemporalConditionModel(**cfg).init_random(seed=3, device=cuda_device)
    sd = a.state_dict()
    flat = torch.empty(sum(v.numel() + 1 for v in sd.values()) + 1, device=cuda_device)
    views, off = {}, 1
    for k, v in sd.items():
        views[k] = flat[off:off + v.numel()].view(v.shape).copy_(v)
        off += v.numel() + 1
    b = UNetSpatioTemporalConditionModel(**cfg).to(cuda_device)
    b.load_state_dict(views)
    x = torch.randn(2, 2, 18, 8, 16, device=cuda_device)
    ehs, ids = torch.randn(2, 1, 64, device=cuda_device), torch.tensor([[6.0, 127.0, 0.02]] * 2, device=cuda_device)
    assert torch.equal(a(x, 0.5, ehs, ids).sample, b(x, 0.5, ehs, ids).sample)
