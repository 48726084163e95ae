"""StableVideoDiffusionPipeline.__call__ and forward_evoworld.process_batch on the GPU against the oracle's restatement
of evoworld/pipeline/pipeline_evoworld.py:456-741 (oracle/pipeline_torch.py), with the same stub VAE / CLIP objects on
both sides and identical seeds: conditioning assembly (a3) must be bit-identical (same torch ops on the same device),
the denoise loop agrees within the UNet tolerance."""
import argparse
from types import SimpleNamespace
import math

import numpy as np
import pytest
import torch

from evoworld_b200.pipeline import StableVideoDiffusionPipeline
from evoworld_b200.plucker import equirectangular_to_ray
from evoworld_b200.unet import UNetSpatioTemporalConditionModel
from oracle import pipeline_torch as OP
from oracle import unet_torch as O

pytestmark = pytest.mark.gpu

CFG = dict(in_channels=18, block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), cross_attention_dim=64)
H, W, T = 128, 256, 3


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.fixture(scope="module")
def rig(cuda_device, built_lib):
    dev = cuda_device
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    oracle = O.UNetSpatioTemporalConditionModel(**CFG).eval().to(dev)
    ours = UNetSpatioTemporalConditionModel(**CFG, num_frames=T).to(dev)
    ours.load_state_dict(oracle.state_dict())
    vae, clip = OP.StubVAE().to(dev).eval(), OP.StubCLIP(64).to(dev).eval()
    pipe = StableVideoDiffusionPipeline(vae=vae, image_encoder=clip, unet=ours).to(dev)
    g = torch.Generator().manual_seed(4)
    image = (torch.rand(1, 3, H, W, generator=g) * 2 - 1).to(dev)
    memory = (torch.rand(1, T, 3, H, W, generator=g) * 2 - 1).to(dev)
    plucker = torch.randn(1, T, 6, H // 8, W // 8, generator=g).to(dev)
    return dict(dev=dev, oracle=oracle, ours=ours, vae=vae, clip=clip, pipe=pipe, image=image, memory=memory, plucker=plucker)


def _oracle_state(rig, generator, mask_mem, steps):
    with torch.no_grad():
        return OP.prepare(rig["oracle"], rig["vae"], rig["clip"], rig["image"], rig["memory"], rig["plucker"], height=H, width=W,
                          num_frames=T, num_inference_steps=steps, generator=generator, mask_mem=mask_mem, device=rig["dev"])


@pytest.mark.parametrize("mask_mem", [False, True])
@pytest.mark.parametrize("cpu_generator", [False, True])
def test_conditioning_assembly_is_bit_identical(rig, mask_mem, cpu_generator):
    """a3 (:570-656): CLIP embedding, VAE latents (+noise augmentation), mask_mem zeroing, Plücker duplicated (not zeroed)
    for the unconditional half, first-frame latent repeated T times, channel order [first | memory | Plücker]."""
    pipe, dev = rig["pipe"], rig["dev"]

    def gen():
        if cpu_generator:
            return torch.Generator().manual_seed(-1 & 0xFFFFFFFF)  # Navigator: torch.manual_seed(-1) global CPU generator
        torch.manual_seed(42)  # process_batch: no generator, CUDA Philox under the global seed (unified:515)
        return None

    st = _oracle_state(rig, gen(), mask_mem, 4)
    emb, cond, ids = pipe.prepare_conditioning(rig["image"], rig["memory"], rig["plucker"], H, W, T, 7 - 1, 127, 0.02, gen(), mask_mem)
    assert torch.equal(emb, st["image_embeddings"])
    assert torch.equal(cond, st["conditional_latents"])
    assert torch.equal(ids, st["added_time_ids"])
    assert cond.shape == (2, T, 14, H // 8, W // 8)
    assert torch.count_nonzero(cond[0, :, :4]) == 0 and torch.count_nonzero(emb[0]) == 0   # unconditional half
    assert torch.equal(cond[0, :, 8:], cond[1, :, 8:]) and torch.equal(cond[1, :, 8:], rig["plucker"][0])  # Plücker kept
    assert torch.equal(cond[1, 0, :4], cond[1, T - 1, :4])                                 # first-frame latent repeated
    assert (torch.count_nonzero(cond[:, :, 4:8]) == 0) == mask_mem


@pytest.mark.parametrize("cpu_generator", [False, True])
def test_pipeline_call_latents(rig, cpu_generator):
    """__call__(output_type="latent") vs the oracle loop: same initial noise, first step within the fused-step tolerance,
    final latents within the accumulated UNet tolerance; callback_on_step_end sees every step and can replace latents."""
    pipe, dev = rig["pipe"], rig["dev"]
    steps = 4

    def gen():
        if cpu_generator:
            return torch.Generator().manual_seed(7)
        torch.manual_seed(42)
        return None

    st = _oracle_state(rig, gen(), False, steps)
    trace = []
    with torch.no_grad():
        want = OP.denoise_loop(rig["oracle"], st, callback=lambda i, t, x: trace.append(x.clone()))
    seen = []

    def cb(p, i, t, kw):
        assert p is pipe and set(kw) == {"latents"}
        seen.append((i, float(t), kw["latents"].clone()))
        return {}

    out = pipe(rig["image"], height=H, width=W, num_frames=T, num_inference_steps=steps, generator=gen(),
               plucker_embedding=rig["plucker"], memorized_pixel_values=rig["memory"], mask_mem=False, output_type="latent",
               callback_on_step_end=cb)
    got = out.frames
    assert got.shape == (1, T, 4, H // 8, W // 8) and torch.isfinite(got).all()
    assert [s[0] for s in seen] == list(range(steps)) and pipe.num_timesteps == steps
    assert np.allclose([s[1] for s in seen], [0.25 * math.log(float(s)) for s in st["sigmas"][:-1]], atol=1e-6)
    errs = [rel_l2(seen[i][2], trace[i]) for i in range(steps)]
    print(f"pipeline latents (cpu_generator={cpu_generator}): per-step rel L2 {['%.2e' % e for e in errs]}, final {rel_l2(got, want):.3e}")
    assert errs[0] < 1e-5
    assert rel_l2(got, want) < 5e-3
    # return_dict=False and a callback that rewrites the latents (both sides halve them after every step)
    with torch.no_grad():
        want2 = OP.denoise_loop(rig["oracle"], _oracle_state(rig, gen(), True, 2), callback=lambda i, t, x: x * 0.5)
    got2 = pipe(rig["image"], height=H, width=W, num_frames=T, num_inference_steps=2, generator=gen(),
                plucker_embedding=rig["plucker"], memorized_pixel_values=rig["memory"], mask_mem=True, output_type="latent",
                callback_on_step_end=lambda p, i, t, kw: {"latents": kw["latents"] * 0.5}, return_dict=False)
    assert rel_l2(got2, want2) < 5e-3


def test_pipeline_decode_outputs(rig):
    """output_type pt / np / pil through decode_latents (:358-385, chunked by decode_chunk_size) with the stub VAE."""
    pipe = rig["pipe"]
    torch.manual_seed(1)
    lat = pipe(rig["image"], height=H, width=W, num_frames=T, num_inference_steps=2, plucker_embedding=rig["plucker"],
               memorized_pixel_values=rig["memory"], output_type="latent").frames
    with torch.no_grad():
        want = OP.decode(rig["vae"], lat, T, 2)
    torch.manual_seed(1)
    pt = pipe(rig["image"], height=H, width=W, num_frames=T, num_inference_steps=2, plucker_embedding=rig["plucker"],
              memorized_pixel_values=rig["memory"], output_type="pt", decode_chunk_size=2).frames
    assert pt.shape == (1, T, 3, H, W) and torch.allclose(pt, want, atol=1e-6)
    torch.manual_seed(1)
    pil = pipe(rig["image"], height=H, width=W, num_frames=T, num_inference_steps=2, plucker_embedding=rig["plucker"],
               memorized_pixel_values=rig["memory"], decode_chunk_size=8).frames[0]
    assert len(pil) == T and pil[0].size == (W, H)
    assert np.array_equal(np.asarray(pil[1]), (want[0, 1].permute(1, 2, 0).cpu().numpy() * 255).round().astype("uint8"))


def test_pipeline_argument_errors(rig):
    pipe = rig["pipe"]
    kw = dict(height=H, width=W, num_frames=T, num_inference_steps=1, plucker_embedding=rig["plucker"],
              memorized_pixel_values=rig["memory"], output_type="latent")
    with pytest.raises(ValueError, match="divisible by 8"):
        pipe(rig["image"], **{**kw, "height": H + 4})
    with pytest.raises(ValueError, match="has to be of type"):
        pipe([1, 2, 3], **kw)
    with pytest.raises(ValueError, match="memory frames and Plücker"):
        pipe(rig["image"], **{**kw, "num_frames": T + 1})
    with pytest.raises(ValueError, match="plucker_embedding"):
        pipe(rig["image"], **{**kw, "plucker_embedding": None})


def test_process_batch_through_dropin(rig, tmp_path, monkeypatch):
    """forward_evoworld.process_batch (:183-211) resolved through the dropin overlay: dataset-style batch -> relative
    c2w -> Plücker -> pipeline -> PNG files; the prepared tensors equal the oracle's restatement of :119-156."""
    import importlib
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    monkeypatch.syspath_prepend(str(root / "dropin"))
    for m in [k for k in sys.modules if k.split(".")[0] == "evoworld"]:
        monkeypatch.delitem(sys.modules, m)
    fe = importlib.import_module("evoworld.inference.forward_evoworld")
    assert "dropin" in fe.__file__
    dev = rig["dev"]
    from evoworld_b200 import synthetic

    poses = torch.from_numpy(synthetic.curve_trajectory()[100:100 + T].astype(np.float32))[None]
    poses[..., :3] *= 0.1
    g = torch.Generator().manual_seed(11)
    batch = {"pixel_values": torch.rand(1, T, 3, H, W, generator=g) * 2 - 1, "cam_traj": poses,
             "memorized_pixel_values": torch.rand(1, T, 3, H, W, generator=g) * 2 - 1}
    args = argparse.Namespace(num_frames=T, height=H, width=W, mask_mem=False)
    rays = torch.tensor(equirectangular_to_ray(target_H=H // 8, target_W=W // 8)).float().to(dev)
    first, traj, plucker, mem, images = fe.prepare_batch_data(batch, args, rays, torch.float32)
    o_first, o_traj, o_plucker, o_mem, _ = OP.prepare_batch_data(batch, T, H, W, rays.cpu(), dev)
    assert torch.equal(first, o_first) and torch.equal(mem, o_mem)
    assert torch.allclose(traj, o_traj, atol=1e-6) and torch.allclose(plucker, o_plucker, atol=2e-6)
    # the reference fixes 25 steps inside process_batch's pipeline call; shorten through the default for the test
    orig = StableVideoDiffusionPipeline.__call__

    def short(self, *a, **k):
        k.setdefault("num_inference_steps", 2)
        return orig(self, *a, **k)

    monkeypatch.setattr(StableVideoDiffusionPipeline, "__call__", short)
    torch.manual_seed(42)
    frames = fe.process_batch(batch, args, rig["pipe"], rays, torch.float32, str(tmp_path), "case_000")
    assert len(frames) == T
    for sub in ("predictions", "predictions_gt"):
        files = sorted(p.name for p in (tmp_path / "case_000" / sub).iterdir())
        assert files == [f"{i + 1:03}.png" for i in range(T)]
    # same seed, the oracle's composition of the same call
    torch.manual_seed(42)
    with torch.no_grad():
        st = OP.prepare(rig["oracle"], rig["vae"], rig["clip"], o_first, o_mem, o_plucker, height=H, width=W, num_frames=T,
                        num_inference_steps=2, generator=None, mask_mem=False, device=dev)
        want = OP.decode(rig["vae"], OP.denoise_loop(rig["oracle"], st), T, 8)
    got = np.stack([np.asarray(f) for f in frames]).astype(np.int32)
    ref = (want[0].permute(0, 2, 3, 1).cpu().numpy() * 255).round().astype(np.int32)
    # uint8 frames decoded from latents that agree to ~1e-3: never more than one code value apart
    assert np.abs(got - ref).max() <= 1 and (got != ref).mean() < 0.15


class _OracleVAE(torch.nn.Module):
    """The diffusers call surface (encode().latent_dist.mode(), decode(z, num_frames).sample) over oracle/vae_torch.py."""

    def __init__(self, core):
        super().__init__()
        self.core = core
        self.config = SimpleNamespace(scaling_factor=0.18215, force_upcast=True, block_out_channels=(64, 128, 128, 128))

    @property
    def dtype(self):
        return torch.float32

    def encode(self, x):
        mean, _ = self.core.encode_moments(x)
        return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: mean))

    def decode(self, z, num_frames=1):
        return SimpleNamespace(sample=self.core.decode(z, num_frames))


def test_pipeline_with_the_native_vae(rig):
    """The whole `__call__` (pipeline_evoworld.py:456-744) with nothing but the CLIP encoder injected: VAE encode of the
    first frame + memory frames, the denoise loop, chunked temporal decode (decode_chunk_size 2 -> videos of 2 and 1
    frames) — against the oracle restatement of the same call with the oracle VAE."""
    from evoworld_b200.vae import AutoencoderKLTemporalDecoder
    from oracle import vae_torch as OV

    dev = rig["dev"]
    torch.manual_seed(21)
    boc = (64, 128, 128, 128)
    with torch.device(dev):
        ovae = OV.AutoencoderKLTemporalDecoder(block_out_channels=boc).eval()
    vae = AutoencoderKLTemporalDecoder(block_out_channels=boc).to(dev)
    vae.load_state_dict(ovae.state_dict())
    pipe = StableVideoDiffusionPipeline(vae=vae, image_encoder=rig["clip"], unet=rig["ours"]).to(dev)
    steps = 2
    got = pipe(rig["image"], height=H, width=W, num_frames=T, num_inference_steps=steps, plucker_embedding=rig["plucker"],
               memorized_pixel_values=rig["memory"], generator=torch.Generator(device=dev).manual_seed(5), output_type="pt",
               decode_chunk_size=2).frames
    o = _OracleVAE(ovae)
    with torch.no_grad():
        st = OP.prepare(rig["oracle"], o, rig["clip"], rig["image"], rig["memory"], rig["plucker"], height=H, width=W, num_frames=T,
                        num_inference_steps=steps, generator=torch.Generator(device=dev).manual_seed(5), mask_mem=False, device=dev)
        want = OP.decode(o, OP.denoise_loop(rig["oracle"], st), T, 2)
    err = rel_l2(got, want)
    print(f"pipeline with the native VAE: frames rel L2 {err:.3e}")
    assert got.shape == (1, T, 3, H, W) and torch.isfinite(got).all()
    assert err < 5e-3  # VAE encode (<= 3e-3) -> 2 denoise steps -> VAE decode (<= 3e-3), frames in [0, 1]


def test_whole_pipeline_at_the_baseline_size_with_native_components(cuda_device, built_lib):
    """`__call__` at 576 x 1024 x 14 frames with nothing injected: full-width UNet (1.525 B), VAE (97.7 M) and CLIP ViT-H/14
    (632 M), random-init, 3 denoise steps, decode_chunk_size 8 — CLIP embedding, VAE encode of the first frame + 14 memory
    frames, the fused steps and the temporal decode all run on this repository's kernels.  (Numerical parity of each part is
    covered at this size by test_gpu_unet / test_gpu_vae / test_gpu_clip; here: it runs, shapes, finiteness, determinism.)"""
    import time

    from evoworld_b200.clip import CLIPVisionModelWithProjection
    from evoworld_b200.vae import AutoencoderKLTemporalDecoder

    dev = cuda_device
    Tn, Hn, Wn = 14, 576, 1024
    unet = UNetSpatioTemporalConditionModel(in_channels=18, num_frames=Tn).init_random(seed=0, device=dev)
    unet._ensure_handle()
    unet.free_master_parameters()
    vae = AutoencoderKLTemporalDecoder().init_random(seed=1, device=dev)
    vae.free_master_parameters()
    clip = CLIPVisionModelWithProjection().init_random(seed=2, device=dev)
    clip.free_master_parameters()
    pipe = StableVideoDiffusionPipeline(vae=vae, image_encoder=clip, unet=unet).to(dev)
    g = torch.Generator().manual_seed(3)
    image = (torch.rand(1, 3, Hn, Wn, generator=g) * 2 - 1).to(dev)
    memory = (torch.rand(1, Tn, 3, Hn, Wn, generator=g) * 2 - 1).to(dev)
    plucker = torch.randn(1, Tn, 6, Hn // 8, Wn // 8, generator=g).to(dev)

    def run():
        return pipe(image, height=Hn, width=Wn, num_frames=Tn, num_inference_steps=3, plucker_embedding=plucker,
                    memorized_pixel_values=memory, generator=torch.Generator(device=dev).manual_seed(7), output_type="pt",
                    decode_chunk_size=8).frames

    a = run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b = run()
    torch.cuda.synchronize()
    print(f"whole pipeline call at 576x1024x14f, 3 steps, native VAE / CLIP / UNet: {time.perf_counter() - t0:.2f} s")
    assert a.shape == (1, Tn, 3, Hn, Wn) and torch.isfinite(a).all()
    assert float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    assert torch.equal(a, b)
