"""CPU checks of the VAE: the oracle restatement (oracle/vae_torch.py) against the published structure, the product's
parameter specification against the oracle's state dict, and the host logic of evoworld_b200/vae.py."""
import pytest
import torch

from evoworld_b200 import vae as V
from oracle import vae_torch as O


def test_oracle_structure_matches_published_vae():
    m = O.AutoencoderKLTemporalDecoder()
    count = lambda mod: sum(p.numel() for p in mod.parameters())
    # the encoder is the AutoencoderKL encoder of SD / SVD: 34 163 592 parameters; quant_conv 8 -> 8
    assert count(m.encoder) == 34_163_592 and count(m.quant_conv) == 72
    assert count(m.decoder) == 63_579_183 and count(m) == 97_742_847
    keys = set(m.state_dict())
    for k in ("encoder.down_blocks.0.downsamplers.0.conv.weight", "encoder.mid_block.attentions.0.group_norm.weight",
              "encoder.mid_block.attentions.0.to_out.0.bias", "decoder.mid_block.resnets.1.temporal_res_block.conv2.weight",
              "decoder.up_blocks.2.resnets.0.spatial_res_block.conv_shortcut.weight", "decoder.up_blocks.3.resnets.2.time_mixer.mix_factor",
              "decoder.up_blocks.0.upsamplers.0.conv.bias", "decoder.time_conv_out.weight", "quant_conv.bias"):
        assert k in keys, k
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in keys and "post_quant_conv.weight" not in keys


@pytest.mark.parametrize("boc", [(128, 256, 512, 512), (64, 128, 128, 128)])
def test_param_spec_equals_oracle_state_dict(boc):
    ours = V.AutoencoderKLTemporalDecoder(block_out_channels=boc)
    sd = O.AutoencoderKLTemporalDecoder(block_out_channels=boc).state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == dict(ours._spec)
    assert ours.num_parameters() == sum(v.numel() for v in sd.values())


def test_oracle_shape_walk_and_temporal_mixing():
    torch.manual_seed(0)
    m = O.AutoencoderKLTemporalDecoder(block_out_channels=(32, 32, 64, 64)).eval()
    x = torch.randn(4, 3, 32, 64)
    with torch.no_grad():
        mean, logvar = m.encode_moments(x)
        assert mean.shape == (4, 4, 4, 8) and logvar.shape == (4, 4, 4, 8)
        y2 = m.decode(mean, num_frames=2)
        y1 = m.decode(mean, num_frames=1)
    assert y2.shape == (4, 3, 32, 64)
    assert not torch.allclose(y1, y2)  # the temporal blocks and time_conv_out mix the frames of a video
    # frames of different videos do not mix
    with torch.no_grad():
        ya = m.decode(mean[:2], num_frames=2)
    assert torch.allclose(ya, y2[:2], atol=1e-5)


def test_distribution_and_host_errors():
    p = torch.randn(2, 8, 4, 4)
    p[:, 4:] = 100.0
    d = V.DiagonalGaussianDistribution(p)
    assert torch.equal(d.mode(), p[:, :4]) and float(d.logvar.max()) == 20.0
    g = torch.Generator().manual_seed(3)
    s1 = d.sample(generator=g)
    g = torch.Generator().manual_seed(3)
    assert torch.equal(s1, d.sample(generator=g)) and s1.shape == (2, 4, 4, 4)
    m = V.AutoencoderKLTemporalDecoder(block_out_channels=(64, 128, 128, 128))
    assert m.config.scaling_factor == 0.18215 and m.config.force_upcast is True
    with pytest.raises(TypeError):
        V.AutoencoderKLTemporalDecoder(bogus=1)
    with pytest.raises(RuntimeError):  # product path needs the GPU library: no CPU fallback
        m.init_random(0).encode(torch.zeros(1, 3, 64, 64))
    import inspect
    assert "num_frames" in inspect.signature(m.forward).parameters  # decode_latents looks for it (pipeline_evoworld.py:364-365)
    sd = O.AutoencoderKLTemporalDecoder(block_out_channels=(64, 128, 128, 128)).state_dict()
    r = m.load_state_dict(sd)
    assert not r.missing_keys and not r.unexpected_keys
    sd.pop("quant_conv.bias")
    with pytest.raises(RuntimeError):
        m.load_state_dict(sd)
