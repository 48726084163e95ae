"""CPU tests of the pipeline-side host logic: CLIP pre-processing pinned against the reference's own function
(golden: tests/golden/make_pipeline_golden.py imports evoworld/trainer/trainer_utils.py:68-179), the scheduler against
the oracle's restatement, and `from_pretrained`'s component loading (fail at construction, clearly)."""
import json
import math
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from evoworld_b200 import image_ops
from evoworld_b200.scheduler import EulerDiscreteScheduler
from oracle import pipeline_torch as OP
from oracle import unet_torch as O

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def pgold():
    return np.load(ROOT / "tests" / "golden" / "pipeline_golden.npz")


def test_resize_with_antialiasing_matches_reference(pgold):
    for src, dst, size in (("aa_in_90x160", "aa_out_35x35", (35, 35)), ("aa_in_24x36", "aa_out_40x12", (40, 12))):
        x = torch.from_numpy(pgold[src])
        got = image_ops.resize_with_antialiasing(x, size).numpy()
        assert np.array_equal(got, pgold[dst]), f"{src}: max diff {np.abs(got - pgold[dst]).max()}"
        assert np.array_equal(OP.resize_with_antialiasing(x, size).numpy(), pgold[dst])  # the oracle's restatement too
    z = torch.rand(1, 3, 576, 1024, generator=torch.Generator().manual_seed(int(pgold["aa_full_seed"]))) * 2 - 1
    got = image_ops.resize_with_antialiasing(z, (224, 224))[0, :, ::16, :].numpy()
    assert np.array_equal(got, pgold["aa_full_rows"])


def test_blur_window_rules():
    assert image_ops.blur_window(576, 576 / 224) == (3, (576 / 224 - 1) / 2)
    assert image_ops.blur_window(1024, 1024 / 224)[0] == 7
    assert image_ops.blur_window(10, 0.5) == (3, 0.001)         # up-scaling: sigma clamps, window 3
    assert image_ops.blur_window(36, 3.0)[0] == 5               # int(4 sigma) = 4 is even -> 5
    with pytest.raises(ValueError):
        image_ops.resize_with_antialiasing(torch.zeros(3, 8, 8), (4, 4))


def test_clip_preprocess_matches_reference_recipe():
    """x*2-1 -> resize -> (x+1)/2 -> (x-mean)/std (pipeline_evoworld.py:270-286); with transformers' own
    CLIPImageProcessor in the loop the result is the same as the built-in constants."""
    g = torch.Generator().manual_seed(3)
    img = torch.rand(1, 3, 96, 192, generator=g)
    got = image_ops.clip_preprocess(img)
    want = OP.resize_with_antialiasing(img * 2 - 1, (224, 224))
    want = ((want + 1) / 2 - torch.tensor(OP.CLIP_MEAN).view(1, 3, 1, 1)) / torch.tensor(OP.CLIP_STD).view(1, 3, 1, 1)
    assert got.shape == (1, 3, 224, 224)
    assert torch.equal(got, want)
    transformers = pytest.importorskip("transformers")
    fe = transformers.CLIPImageProcessor()
    via_fe = image_ops.clip_preprocess(img, fe)
    assert torch.allclose(via_fe, got, atol=1e-6), float((via_fe - got).abs().max())


def test_scheduler_matches_oracle():
    for n in (25, 50, 4, 1):
        s = EulerDiscreteScheduler()
        s.set_timesteps(n)
        sig = O.karras_sigmas(n)
        assert torch.equal(s.sigmas, sig), n
        assert torch.allclose(s.timesteps, O.sigma_to_timestep(sig[:-1]), rtol=0, atol=1e-6)
        assert s.init_noise_sigma == pytest.approx(float((sig.max() ** 2 + 1) ** 0.5))
    assert EulerDiscreteScheduler().init_noise_sigma == pytest.approx(700.0007142, rel=1e-7)
    # scale_model_input / step == the oracle's loop body with guidance 1 (v_cond only)
    s = EulerDiscreteScheduler()
    s.set_timesteps(25)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 4, 8, 16, generator=g) * 700
    v = torch.randn(1, 3, 4, 8, 16, generator=g)
    for i in (0, 7, 24):
        s._step_index = None
        t = s.timesteps[i]
        sigma, nxt = float(s.sigmas[i]), float(s.sigmas[i + 1])
        assert torch.equal(s.scale_model_input(x, t), x / ((sigma ** 2 + 1) ** 0.5))
        got = s.step(v, t, x).prev_sample
        x0 = v * (-sigma / (sigma ** 2 + 1) ** 0.5) + x / (sigma ** 2 + 1)
        assert torch.allclose(got, x + (x - x0) / sigma * (nxt - sigma), rtol=1e-6, atol=1e-4)
    with pytest.raises(NotImplementedError):
        EulerDiscreteScheduler(prediction_type="epsilon")


def _write_checkpoint(tmp_path, subdirs=()):
    from evoworld_b200.unet import UNetSpatioTemporalConditionModel

    cfg = dict(in_channels=18, block_out_channels=(64, 64, 64, 64), num_attention_heads=(1, 1, 1, 1), cross_attention_dim=64)
    UNetSpatioTemporalConditionModel(**cfg).init_random(seed=0, device="cpu").save_pretrained(str(tmp_path / "unet"))
    os.makedirs(tmp_path / "scheduler")
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps({"sigma_max": 700.0, "use_karras_sigmas": True}))
    for d in subdirs:
        os.makedirs(tmp_path / d)
        (tmp_path / d / "config.json").write_text("{}")
    return cfg


def test_from_pretrained_components(tmp_path):
    from evoworld_b200.pipeline import StableVideoDiffusionPipeline
    from evoworld_b200.unet import UNetSpatioTemporalConditionModel

    from evoworld_b200.vae import AutoencoderKLTemporalDecoder

    _write_checkpoint(tmp_path)
    unet = UNetSpatioTemporalConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    pipe = StableVideoDiffusionPipeline.from_pretrained(str(tmp_path), unet=unet, strict_components=False)
    assert pipe.vae is None and pipe.image_encoder is None and pipe.feature_extractor is None
    assert pipe.unet is unet and pipe.scheduler.init_noise_sigma == pytest.approx(700.0007142, rel=1e-7)
    # a `vae/` sub-folder in diffusers format is loaded by the native VAE — the reference's own call pattern
    # (forward_evoworld.py:103) then needs nothing injected for encode / decode
    saved = AutoencoderKLTemporalDecoder(block_out_channels=(64, 64, 64, 64)).init_random(seed=1)
    saved.save_pretrained(str(tmp_path / "vae"))
    pipe = StableVideoDiffusionPipeline.from_pretrained(str(tmp_path), unet=unet, local_files_only=True, low_cpu_mem_usage=True)
    assert isinstance(pipe.vae, AutoencoderKLTemporalDecoder) and pipe.vae.config.scaling_factor == 0.18215
    assert all(torch.equal(v, saved.state_dict()[k]) for k, v in pipe.vae.state_dict().items())
    # ... and an `image_encoder/` folder in transformers format by the native CLIP encoder: nothing from diffusers or
    # transformers is needed to construct the pipeline
    from evoworld_b200.clip import CLIPVisionModelWithProjection

    clip_saved = CLIPVisionModelWithProjection(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2,
                                               image_size=28, patch_size=14, projection_dim=64).init_random(seed=2)
    clip_saved.save_pretrained(str(tmp_path / "image_encoder"))
    pipe = StableVideoDiffusionPipeline.from_pretrained(str(tmp_path), unet=unet, local_files_only=True, low_cpu_mem_usage=True)
    assert isinstance(pipe.image_encoder, CLIPVisionModelWithProjection) and pipe.image_encoder.config.projection_dim == 64
    assert isinstance(pipe.vae, AutoencoderKLTemporalDecoder) and pipe.feature_extractor is None
    os.remove(tmp_path / "vae" / "diffusion_pytorch_model.safetensors")
    with pytest.raises(FileNotFoundError, match="no weights"):
        StableVideoDiffusionPipeline.from_pretrained(str(tmp_path), unet=unet)
    # injected components are kept as they are
    vae, clip = OP.StubVAE(), OP.StubCLIP()
    pipe = StableVideoDiffusionPipeline.from_pretrained(str(tmp_path), unet=unet, vae=vae, image_encoder=clip)
    assert pipe.vae is vae and pipe.image_encoder is clip
    # without encoders the call itself says what to do
    pipe = StableVideoDiffusionPipeline(unet=unet)
    with pytest.raises(RuntimeError, match="no CLIP image encoder"):
        pipe._encode_image(torch.zeros(1, 3, 16, 32), "cpu")
    with pytest.raises(RuntimeError, match="no VAE"):
        pipe._encode_vae_image(torch.zeros(1, 3, 16, 32), "cpu")


def test_encode_image_uses_the_reference_preprocessing():
    """A recording encoder sees [B,3,224,224] CLIP-normalised pixel values, not the raw panorama."""
    from types import SimpleNamespace

    from evoworld_b200.pipeline import StableVideoDiffusionPipeline

    seen = {}

    class Recorder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))

        def forward(self, pv):
            seen["pv"] = pv
            return SimpleNamespace(image_embeds=pv.mean(dim=(2, 3)))

    pipe = StableVideoDiffusionPipeline(unet=None, image_encoder=Recorder())
    img = torch.rand(1, 3, 64, 128, generator=torch.Generator().manual_seed(9))
    emb = pipe._encode_image(img, "cpu", 1, True)
    assert seen["pv"].shape == (1, 3, 224, 224)
    assert torch.equal(seen["pv"], image_ops.clip_preprocess(img))
    assert emb.shape == (2, 1, 3) and torch.count_nonzero(emb[0]) == 0
    want = OP.encode_image(Recorder(), img, "cpu")
    assert torch.equal(emb, want)


def test_process_batch_signature_matches_reference():
    """Same function names and parameter names as evoworld/inference/forward_evoworld.py (parsed, not imported:
    the module needs diffusers).  Skipped where the reference checkout is absent (the GPU box)."""
    import ast
    import inspect

    ref = Path("/root/reference/evoworld/inference/forward_evoworld.py")
    if not ref.exists():
        pytest.skip("reference checkout not present")
    from evoworld_b200 import inference

    fns = {n.name: [a.arg for a in n.args.args] for n in ast.parse(ref.read_text()).body if isinstance(n, ast.FunctionDef)}
    for name in ("prepare_batch_data", "process_batch", "save_frames"):
        ours = [p for p in inspect.signature(getattr(inference, name)).parameters]
        assert ours[:len(fns[name])] == fns[name], (name, ours, fns[name])
