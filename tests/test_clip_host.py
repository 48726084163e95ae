"""CPU checks of evoworld_b200/clip.py against the library the reference calls (transformers CLIPVisionModelWithProjection,
evoworld/pipeline/pipeline_evoworld.py:22,289): parameter names / shapes, config handling, no CPU fallback."""
import json

import pytest
import torch

from evoworld_b200 import clip as K

transformers = pytest.importorskip("transformers")


def _hf(cfg, device="cpu"):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection

    with torch.device(device):
        return CLIPVisionModelWithProjection(CLIPVisionConfig(**cfg)).eval()


SMALL = dict(hidden_size=320, intermediate_size=640, num_hidden_layers=2, num_attention_heads=4, image_size=56, patch_size=14,
             projection_dim=64, hidden_act="gelu")


@pytest.mark.parametrize("cfg", [SMALL, dict(K.DEFAULT_CONFIG)], ids=["small", "vit-h-14"])
def test_param_spec_equals_transformers_state_dict(cfg):
    hf = _hf(cfg, device="meta")
    want = {k: tuple(v.shape) for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
    ours = K.CLIPVisionModelWithProjection(**cfg)
    assert dict(ours._spec) == want
    if cfg is not SMALL:
        assert ours.num_parameters() == 632_076_800  # ViT-H/14 vision tower + projection


def test_checkpoint_round_trip_and_errors(tmp_path):
    hf = _hf(SMALL)
    hf.save_pretrained(str(tmp_path / "image_encoder"), safe_serialization=True)
    m = K.CLIPVisionModelWithProjection.from_pretrained(str(tmp_path), subfolder="image_encoder")
    assert m.config.hidden_size == 320 and m.config.projection_dim == 64 and m.config.hidden_act == "gelu"
    sd = hf.state_dict()
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    with pytest.raises(RuntimeError):  # no CPU fallback
        m(torch.zeros(1, 3, 56, 56))
    with pytest.raises(NotImplementedError):
        K.CLIPVisionModelWithProjection(hidden_size=100)
    m.save_pretrained(str(tmp_path / "resaved"))
    cfg = json.load(open(tmp_path / "resaved" / "config.json"))
    assert cfg["num_hidden_layers"] == 2
    m2 = K.CLIPVisionModelWithProjection.from_pretrained(str(tmp_path / "resaved"))
    assert all(torch.equal(v, sd[k]) for k, v in m2.state_dict().items())
    bad = dict(sd)
    bad.pop("visual_projection.weight")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_host_orchestration_against_transformers(act, monkeypatch):
    """The host logic of clip.py (weight packing, fused q/k/v, class / position rows, GELU in the GEMM epilogue or as a cast
    pass) with the kernels swapped for the torch restatements of tests/ops_emulation.py, against transformers on the CPU."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import ops_emulation as E

    from evoworld_b200 import _lib, ops

    for n in E.ALL:
        monkeypatch.setattr(ops, n, getattr(E, n))
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: None)
    cfg = dict(SMALL, hidden_act=act)
    torch.manual_seed(0)
    hf = _hf(cfg)
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if "norm" in n:
                p.copy_(torch.randn_like(p) * 0.2 + (1.0 if n.endswith("weight") else 0.0))
            elif n.endswith("bias"):
                p.copy_(torch.randn_like(p) * 0.05)
    m = K.CLIPVisionModelWithProjection(**cfg)
    m.load_state_dict(hf.state_dict())
    x = torch.randn(2, 3, 56, 56)

    class _Dev:   # the pack step only asks the device for its type
        type = "cuda"

    real = m._device
    m._device = _Dev()
    try:
        T = m._pack()
    finally:
        m._device = real
    assert T is m._packed
    with torch.no_grad():
        want = hf(pixel_values=x).image_embeds
    got = m(x).image_embeds
    err = float((got.double() - want.double()).norm() / want.double().norm())
    assert got.shape == want.shape and err < 2e-3, err
