"""Host-side UNet logic (CPU): parameter specification == the oracle's state-dict layout, analytic
parameter count, packing shapes.  No CUDA call."""
import pytest
import torch

from evoworld_b200 import unet as U


def test_param_spec_matches_oracle_and_published_count():
    from oracle.unet_torch import UNetSpatioTemporalConditionModel as Oracle

    cfg = dict(U.DEFAULT_CONFIG, in_channels=18)
    spec = U.param_spec(cfg)
    with torch.device("meta"):
        o = Oracle()
    sd = o.state_dict()
    assert set(spec) == set(sd)
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    m = U.UNetSpatioTemporalConditionModel(in_channels=18)
    # published SVD UNet: 1 524 623 082 parameters; EvoWorld adds 10 conv_in channels x 9 x 320
    assert m.num_parameters() == 1_524_623_082 + 28_800
    assert m.config.in_channels == 18 and m.config.addition_time_embed_dim == 256 and m.config.num_frames == 25
    assert m.add_embedding.linear_1.in_features == 768


def test_block_layout_counts():
    lay = U.block_layout(dict(U.DEFAULT_CONFIG, in_channels=18))
    assert len(lay["res"]) == 22 and len(lay["att"]) == 16 and len(lay["samplers"]) == 6
    ups = [r for r in lay["res"] if r[0].startswith("up_blocks")]
    assert [r[1] for r in ups] == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]


def test_config_validation_and_no_cpu_path():
    with pytest.raises(ValueError):
        U.UNetSpatioTemporalConditionModel(block_out_channels=(320, 640, 1280))
    m = U.UNetSpatioTemporalConditionModel(in_channels=18, block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4),
                                           cross_attention_dim=64)
    m.init_random(0)
    assert len(m.state_dict()) == len(U.param_spec(m._cfg))
    with pytest.raises(RuntimeError):
        m.forward(torch.zeros(1, 2, 18, 8, 16), 1.0, torch.zeros(1, 1, 64), torch.zeros(1, 3))


def test_state_dict_roundtrip(tmp_path):
    kw = dict(in_channels=18, block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), cross_attention_dim=64)
    m = U.UNetSpatioTemporalConditionModel(**kw).init_random(1)
    m.save_pretrained(str(tmp_path / "unet"))
    m2 = U.UNetSpatioTemporalConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    assert m2.config.block_out_channels == [64, 128, 256, 256] or tuple(m2.config.block_out_channels) == (64, 128, 256, 256)
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k])
    with pytest.raises(FileNotFoundError):
        U.UNetSpatioTemporalConditionModel.from_pretrained(str(tmp_path / "nope"), subfolder="unet")
