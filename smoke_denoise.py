"""smoke() leg of the denoise path (called by __graft_entry__.smoke only; test infrastructure, not product code):
one tiny UNet forward + fused denoise step on cuda:0 against the fp32 PyTorch oracle.  The oracle is initialised on
the host and copied over, so the first kernels the process launches are the library's own (tc_gemm_kernel,
spatial_attn8_kernel, ...), not a few hundred torch RNG-init launches."""
from __future__ import annotations

import torch


def run(dev) -> None:
    from evoworld_b200.unet import UNetSpatioTemporalConditionModel
    from oracle import unet_torch as O

    cfg = dict(in_channels=18, block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), cross_attention_dim=64)
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    oracle = O.UNetSpatioTemporalConditionModel(**cfg).eval()  # host init: no device RNG launches
    flat = torch.nn.utils.parameters_to_vector(oracle.parameters()).to(dev)  # one H2D copy
    oracle = oracle.to_empty(device=dev)
    torch.nn.utils.vector_to_parameters(flat, oracle.parameters())
    ours = UNetSpatioTemporalConditionModel(**cfg).to(dev)
    ours.load_state_dict(oracle.state_dict())
    T, h, w = 3, 16, 32
    lat = (torch.randn(1, T, 4, h, w) * 700.0007).to(dev)
    cond = torch.randn(2, T, 14, h, w).to(dev)
    ehs = torch.randn(2, 1, 64).to(dev)
    ids = torch.tensor([[6.0, 127.0, 0.02]] * 2).to(dev)
    guid = torch.linspace(1.0, 3.0, T, device=dev).view(1, T, 1, 1, 1)
    with torch.no_grad():
        want_v = oracle(torch.cat([torch.cat([lat] * 2) / (700.0 ** 2 + 1) ** 0.5, cond], dim=2), 0.25 * torch.log(torch.tensor(700.0)), ehs, ids)
        want = O.denoise_step(oracle, lat, cond, 700.0, 545.7, ehs, ids, guid)
    got_v = ours(torch.cat([torch.cat([lat] * 2) / (700.0 ** 2 + 1) ** 0.5, cond], dim=2), 0.25 * float(torch.log(torch.tensor(700.0))), ehs, ids).sample
    got = ours.denoise_step(lat.clone(), cond, 700.0, 545.7, ehs, ids, 1.0, 3.0)
    e_v = float((got_v - want_v).norm() / want_v.norm())
    e_x = float((got - want).norm() / want.norm())
    assert e_v < 1e-3 and e_x < 1e-5, f"denoise smoke: UNet rel-L2 {e_v:.2e}, step rel-L2 {e_x:.2e}"
    print(f"smoke: denoise step OK (UNet rel-L2 {e_v:.2e}, latents rel-L2 {e_x:.2e} vs fp32 oracle)")
