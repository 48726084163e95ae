"""CLIP ViT image encoder on the GPU (evoworld_b200/clip.py: tcgen05 GEMMs + LayerNorm + small-S attention) against
transformers' own CLIPVisionModelWithProjection — the class the reference pipeline calls (pipeline_evoworld.py:22,289) — in
fp32 on the same GPU with TF32 off.  Tolerance: fp16 GEMM / attention operands with fp32 accumulation and an fp32 residual
stream through 32 pre-LN layers: measured 3.4e-4 - 4.0e-4, asserted <= 1e-3 relative L2 on the image embedding."""
import pytest
import torch
import torch.nn.functional as F

from evoworld_b200 import clip as K
from evoworld_b200 import ops

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")

TOL_CLIP = 1e-3


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def make_pair(cfg, dev, seed=0):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    torch.manual_seed(seed)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    hf = CLIPVisionModelWithProjection(CLIPVisionConfig(**cfg)).eval()
    with torch.no_grad():  # transformers initialises LayerNorms to 1 / 0 and biases to 0: make every parameter matter
        for n, p in hf.named_parameters():
            if "norm" in n:
                p.copy_(torch.randn_like(p) * 0.2 + (1.0 if n.endswith("weight") else 0.0))
            elif n.endswith("bias"):
                p.copy_(torch.randn_like(p) * 0.05)
    hf = hf.to(dev)
    ours = K.CLIPVisionModelWithProjection(**cfg).to(dev)
    ours.load_state_dict(hf.state_dict())
    return hf, ours


@pytest.mark.parametrize("B,S,H,D", [(2, 257, 16, 80), (1, 17, 4, 80), (3, 100, 2, 64), (1, 1024, 1, 16), (2, 33, 3, 96)])
def test_small_attention(B, S, H, D, cuda_device, built_lib):
    torch.manual_seed(0)
    qkv = torch.randn(B * S, 3 * H * D, device=cuda_device).half()
    got = ops.small_attention(qkv, B, S, H, D, D ** -0.5)
    q, k, v = [t.reshape(B, S, H, D).transpose(1, 2).float() for t in qkv.chunk(3, dim=1)]
    want = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * S, H * D)
    assert rel_l2(got, want) < 6e-4  # the fp16 rounding of the output


@pytest.mark.parametrize("mode", ["gelu", "quick_gelu"])
def test_activation(mode, cuda_device, built_lib):
    x = torch.randn(1000, 640, device=cuda_device) * 3
    want = F.gelu(x) if mode == "gelu" else x * torch.sigmoid(1.702 * x)
    assert rel_l2(ops.activation_f16(x, mode), want) < 4e-4


def test_layer_norm_f32(cuda_device, built_lib):
    x = torch.randn(257, 1280, device=cuda_device) * 2 + 0.3
    g, b = torch.randn(1280, device=cuda_device), torch.randn(1280, device=cuda_device)
    assert rel_l2(ops.layer_norm_f32(x, g, b, 1e-5), F.layer_norm(x, (1280,), g, b, 1e-5)) < 1e-6


@pytest.mark.parametrize("cfg", [
    dict(hidden_size=320, intermediate_size=640, num_hidden_layers=2, num_attention_heads=4, image_size=56, patch_size=14,
         projection_dim=64, hidden_act="gelu"),
    dict(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, image_size=64, patch_size=16,
         projection_dim=100, hidden_act="quick_gelu"),
], ids=["gelu-hd80", "quickgelu-hd64"])
def test_small_configs(cfg, cuda_device, built_lib):
    hf, ours = make_pair(cfg, cuda_device, seed=1)
    x = torch.randn(3, 3, cfg["image_size"], cfg["image_size"], device=cuda_device)
    with torch.no_grad():
        want = hf(x)
    got = ours(x)
    e1, e2 = rel_l2(got.image_embeds, want.image_embeds), rel_l2(got.last_hidden_state, want.last_hidden_state)
    print(f"clip small: image_embeds rel L2 {e1:.3e}, last_hidden_state {e2:.3e}")
    assert got.image_embeds.shape == (3, cfg["projection_dim"])
    assert e1 < TOL_CLIP and e2 < TOL_CLIP


def test_vit_h_14(cuda_device, built_lib):
    """The image encoder of Stable Video Diffusion: ViT-H/14, 632 M parameters, 257 tokens, 16 heads of 80."""
    hf, ours = make_pair(dict(K.DEFAULT_CONFIG), cuda_device, seed=2)
    x = torch.randn(2, 3, 224, 224, device=cuda_device)
    with torch.no_grad():
        want = hf(x).image_embeds
    got = ours(x).image_embeds
    err = rel_l2(got, want)
    print(f"clip ViT-H/14: image_embeds rel L2 {err:.3e}")
    assert got.shape == (2, 1024) and err < TOL_CLIP
    assert torch.equal(ours(x).image_embeds, got)


def test_pipeline_encode_image_with_the_native_encoder(cuda_device, built_lib):
    """`_encode_image` (pipeline_evoworld.py:255-305): [0,1] image -> antialiased resize to 224 -> CLIP normalisation ->
    image encoder -> [uncond zeros; embedding]; native encoder vs transformers' on the same pre-processed pixels."""
    from evoworld_b200.pipeline import StableVideoDiffusionPipeline

    cfg = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2, image_size=224, patch_size=14,
               projection_dim=64, hidden_act="gelu")
    hf, ours = make_pair(cfg, cuda_device, seed=3)
    torch.manual_seed(4)
    img = torch.rand(1, 3, 128, 256, device=cuda_device)
    a = StableVideoDiffusionPipeline(image_encoder=ours)._encode_image(img, cuda_device, 1, True)
    b = StableVideoDiffusionPipeline(image_encoder=hf)._encode_image(img, cuda_device, 1, True)
    assert a.shape == b.shape == (2, 1, 64) and torch.equal(a[0], torch.zeros_like(a[0]))
    assert rel_l2(a[1], b[1]) < TOL_CLIP
