"""Host logic of evoworld_b200/vggt.py without a GPU: the kernels are replaced by the torch restatements of
tests/ops_emulation.py (same signatures and rounding points) and the result is compared with the golden vectors of the
REFERENCE modules — checks weight packing (LayerScale folding, ConvTranspose / stride-2 / padded layouts), tap order, the
broadcast indices of the positional embeddings, token assembly and the output dictionary."""
from pathlib import Path

import numpy as np
import pytest
import torch

import evoworld_b200.vggt as V
from evoworld_b200 import _lib, ops
from oracle import vggt_torch as O
import sys

sys.path.insert(0, str(Path(__file__).resolve().parent))
import ops_emulation as E  # noqa: E402

CFG = O.SMALL_TEST_CONFIG


@pytest.fixture()
def emulated(monkeypatch):
    for n in E.ALL:
        monkeypatch.setattr(ops, n, getattr(E, n))
    monkeypatch.setattr(_lib, "require_cuda", lambda t, name: None)
    monkeypatch.setattr(V, "_check_device", lambda dev: None)
    m = V.VGGT(**{k: v for k, v in CFG.items() if k != "seed"})
    m.load_state_dict(V.random_state_dict(CFG, seed=CFG["seed"]))
    return m


def rel_l2(a, b):
    a = a.double().numpy() if isinstance(a, torch.Tensor) else a.astype(np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_host_orchestration_against_the_reference(emulated):
    vg = np.load(Path(__file__).resolve().parent / "golden" / "vggt_golden.npz")
    images = O.small_test_images()
    toks, start = emulated.aggregator(images)
    assert start == 5 and len(toks) == CFG["depth"]
    errs = {f"tokens_{i}": rel_l2(t, vg[f"tokens_{i}"]) for i, t in enumerate(toks)}
    emulated.free_master_parameters()          # the packed set alone must serve a forward
    out = emulated(images[0], frames_chunk_size=2)
    for k in ("depth", "depth_conf", "world_points", "world_points_conf"):
        assert out[k].shape == vg[k].shape
        errs[k] = rel_l2(out[k], vg[k])
    errs["pose_enc"] = rel_l2(out["pose_enc"], vg["pose_enc_3"])
    print(errs)
    assert max(errs.values()) < 5e-3, errs
    assert out["images"].shape == (1, 3, 3, 70, 98)


def test_batched_chunked_orchestration_against_the_oracle(emulated):
    """Batch of two sequences, DPT heads chunked 2 + 1 frames: the row order (batch, frame, patch) through the aggregator, the
    camera head and the chunked heads against the oracle restatement on the CPU."""
    g = torch.Generator().manual_seed(5)
    images = torch.rand((2, 3, 3, 56, 84), generator=g)
    mcfg = {k: v for k, v in CFG.items() if k != "seed"}
    sd = V.random_state_dict(CFG, seed=CFG["seed"])
    with torch.no_grad():
        want = O.vggt_forward(images, sd, mcfg)
    out = emulated(images, frames_chunk_size=2)
    for k in ("pose_enc", "depth", "depth_conf", "world_points", "world_points_conf"):
        assert out[k].shape == want[k].shape, k
        assert rel_l2(out[k], want[k].double().numpy()) < 5e-3, k
    whole = emulated(images, frames_chunk_size=None)
    assert torch.equal(whole["depth"], out["depth"]) and torch.equal(whole["world_points_conf"], out["world_points_conf"])
