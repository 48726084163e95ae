"""Denoise-step arm of bench.py (imported by it): UNet forward on the CFG batch + CFG combine +
Euler step (pipeline_evoworld.py:689-725) at 576x1024 (72x128 latents), T frames, random-init UNet."""
from __future__ import annotations

import os
import time

import numpy as np

UNET_CFG = dict(in_channels=18, block_out_channels=(320, 640, 1280, 1280), num_attention_heads=(5, 10, 20, 20),
                cross_attention_dim=1024)
LAT_H, LAT_W = 72, 128


def make_inputs(T, h, w, dev, seed):
    """SURVEY §8d synthetic inputs: latents ~ 700 N(0,1), cond = [first-frame latent | memory latent | Plücker]."""
    import torch

    from evoworld_b200 import synthetic
    from evoworld_b200.plucker import equirectangular_to_ray, ray_c2w_to_plucker

    g = torch.Generator(device="cpu").manual_seed(seed)
    lat = torch.randn(1, T, 4, h, w, generator=g) * 700.0007
    cond = torch.randn(2, T, 14, h, w, generator=g)
    cond[0, :, :8] = 0
    ehs = torch.randn(2, 1, 1024, generator=g)
    ehs[0] = 0
    ids = torch.tensor([[6.0, 127.0, 0.02]] * 2)
    if dev is not None:
        ray = torch.from_numpy(equirectangular_to_ray(h, w)).to(dev)
        poses = synthetic.curve_trajectory()[101:101 + T].copy()
        poses[:, :3] *= 0.1
        c2w = torch.from_numpy(synthetic.euler_c2w(poses)[:, :3, :4]).float().to(dev)
        pl = ray_c2w_to_plucker(ray, c2w).cpu()
        cond[:, :, 8:14] = pl[None]
    return lat, cond, ehs, ids


def run_ours(args, dev, rank, world, barrier, allreduce_max, peaks):
    import torch

    from evoworld_b200.scheduler import EulerDiscreteScheduler
    from evoworld_b200.unet import UNetSpatioTemporalConditionModel, algorithmic_flops, DEFAULT_CONFIG

    T = args.frames
    h, w = getattr(args, "pano_height", 576) // 8, getattr(args, "pano_width", 1024) // 8
    unet = UNetSpatioTemporalConditionModel(**UNET_CFG).init_random(seed=0, device=dev)
    unet._ensure_handle()
    unet.free_master_parameters()
    lat_h, cond_h, ehs_h, ids_h = make_inputs(T, h, w, dev, seed=rank)
    lat_p, cond_p = lat_h.pin_memory(), cond_h.pin_memory()
    out_p = torch.empty_like(lat_h).pin_memory()
    lat, cond, ehs, ids = lat_h.to(dev), cond_h.to(dev), ehs_h.to(dev), ids_h.to(dev)
    sched = EulerDiscreteScheduler()
    sched.set_timesteps(25)
    sig = [float(s) for s in sched.sigmas]

    def step(i, x):
        unet.denoise_step(x, cond, sig[i % 25], sig[i % 25 + 1], ehs, ids, 1.0, 3.0)

    x = lat.clone()
    for i in range(args.warmup):
        step(i, x)
    torch.cuda.synchronize(dev)
    launches, plan_flops = unet.plan_info()
    x.copy_(lat)
    barrier(world)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from evoworld_b200.distributed import gather_latents

    replays0 = unet.graph_replays()
    e0.record()
    for i in range(args.steps):
        step(i, x)
    graph_replays_timed = unet.graph_replays() - replays0
    gathered = gather_latents(x)  # clip boundary: the only collective of the path (no-op at world == 1)
    e1.record()
    torch.cuda.synchronize(dev)
    barrier(world)
    ms = e0.elapsed_time(e1)
    finite = bool(torch.isfinite(x).all())

    # end to end: the step's inputs come from pinned host memory, the result goes back to the host
    xd = torch.empty_like(lat)
    cd = torch.empty_like(cond)
    for i in range(min(args.warmup, 2)):  # untimed: these device buffers are new to the plan -> their CUDA graph is captured here
        xd.copy_(lat_p, non_blocking=True)
        cd.copy_(cond_p, non_blocking=True)
        unet.denoise_step(xd, cd, sig[i % 25], sig[i % 25 + 1], ehs, ids, 1.0, 3.0)
        out_p.copy_(xd, non_blocking=True)
    torch.cuda.synchronize(dev)
    barrier(world)
    t0 = time.perf_counter()
    for i in range(args.steps):
        xd.copy_(lat_p, non_blocking=True)
        cd.copy_(cond_p, non_blocking=True)
        unet.denoise_step(xd, cd, sig[i % 25], sig[i % 25 + 1], ehs, ids, 1.0, 3.0)
        out_p.copy_(xd, non_blocking=True)
        torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0

    ms = allreduce_max(ms, dev, world)
    e2e_s = allreduce_max(e2e_s, dev, world)
    fl = algorithmic_flops(dict(DEFAULT_CONFIG, **UNET_CFG), 2, T, h, w)
    kernels = per_kernel_roofline(unet, lambda: step(0, xd), fl, peaks, dev) if rank == 0 else None
    steps_per_s = world * args.steps / (ms * 1e-3)
    achieved = fl["total"] * (args.steps / (ms * 1e-3))  # TFLOP/s per GPU
    res = {
        "metric": "denoise-steps/sec", "unit": "steps/s", "value": steps_per_s, "ms_per_step": ms / args.steps, "dtype": "fp16",
        "e2e": {"value": world * args.steps / e2e_s, "unit": "steps/s",
                "h2d_bytes_per_step": lat_p.numel() * 4 + cond_p.numel() * 4, "d2h_bytes_per_step": out_p.numel() * 4},
        "roofline": {"bound": "tensor", "kernel": "whole denoise step (tc_gemm_kernel + spatial_attn8_kernel dominate; per-kernel figures under kernels)",
                     "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
                     "traffic": None, "algorithmic_tflop_per_step": fl, "executed_tflop_per_step": plan_flops / 1e12,
                     # the same fraction on the FLOPs the plan executes (folded cross-attention projections, fused up-sampling
                     # convolutions): what the tensor pipe really does per second
                     "frac_executed": plan_flops / 1e12 * (args.steps / (ms * 1e-3)) / peaks["tf_sustained"],
                     "kernels": kernels,
                     "peak_source": peaks["source"] + " (sustained bf16/fp16 dense)"},
        "gpu_launches": launches * args.steps,
        "config": {"workload": f"{'config 2' if (h, w) == (LAT_H, LAT_W) else 'config 5 panorama size'}: single {8 * h}x{8 * w}x{T}f clip, CFG batch 2, {h}x{w} latents, random-init 1.525B-param UNet, "
                               f"Karras sigmas (25-step schedule)", "frames": T, "finite_output": finite,
                   "collective": f"all-gather of latents {tuple(gathered.shape)} at the clip boundary",
                   "l2": "working set (activations + 3 GB of fp16 weights) >> 126 MB L2; K steps in one CUDA-event pair",
                   "graph_replays": graph_replays_timed,
                   "launch": "each step = 1 set-args kernel + 1 cudaGraphLaunch of the captured plan (gpu_launches counts the kernels inside)"},
    }
    if rank == 0 and world == 1 and not getattr(args, "no_eager_baseline", False):
        try:
            eager = run_gpu_eager(T, h, w, dev)
            best = min(v["ms_per_step"] for k, v in eager.items() if isinstance(v, dict))
            eager["ours_vs_best_eager"] = best / (ms / args.steps)
            eager["ours_vs_fp32_eager"] = eager["fp32"]["ms_per_step"] / (ms / args.steps)
            res["gpu_eager_baseline"] = eager
        except Exception as exc:  # a baseline leg must not take the bench line down
            res["gpu_eager_baseline"] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the CPU baseline is an N = 1 figure
        res["cpu_baseline"] = run_cpu(T, steps=1)
    return res


def per_kernel_roofline(unet, run_step, fl, peaks, dev):
    """Live per-kernel figures: one extra step replayed with a CUDA-event pair around every planned op
    (EVW_UNET_PROFILE, csrc/unet_host.cu) — serialised, so the sum is a little above the pipelined step time."""
    import re
    import tempfile

    import torch

    path = os.path.join(tempfile.gettempdir(), f"evw_unet_profile_{os.getpid()}.txt")
    if os.path.exists(path):
        os.remove(path)
    os.environ["EVW_UNET_PROFILE"] = path
    try:
        run_step()
        torch.cuda.synchronize(dev)
    finally:
        del os.environ["EVW_UNET_PROFILE"]
    gemm_ms = gemm_fl = attn_ms = tattn_ms = norm_ms = other_ms = 0.0
    try:
        for line in open(path):
            if line.startswith("END"):
                break
            _, t, label = line.split(" ", 2)
            t = float(t)
            m = re.search(r"(\d+)x(\d+)x(\d+)\s*$", label)
            name = label.split(" ")[0]
            if m:
                M, N, K = map(int, m.groups())
                gemm_ms += t
                gemm_fl += 2.0 * M * N * K
            elif "sdpa" in name:
                if "temporal_transformer" in name:
                    tattn_ms += t
                else:
                    attn_ms += t
            elif "norm" in name:
                norm_ms += t
            else:
                other_ms += t
        os.remove(path)
    except OSError:
        return None
    if gemm_ms <= 0 or attn_ms <= 0:
        return None
    pk = peaks["tf_sustained"]
    return {"how": "one step replayed with CUDA events around every op (serialised)",
            "tc_gemm_kernel": {"ms": gemm_ms, "tflops": gemm_fl / gemm_ms / 1e9, "frac": gemm_fl / gemm_ms / 1e9 / pk},
            "spatial_attn_kernel": {"ms": attn_ms, "tflops": fl["sdpa_spatial"] * 1e3 / attn_ms,
                                    "frac": fl["sdpa_spatial"] * 1e3 / attn_ms / pk},
            "temporal_attn_ms": tattn_ms, "group_layer_norm_ms": norm_ms, "other_ms": other_ms}


def _cpu_oracle():
    """fp32 PyTorch oracle UNet (full 1.525 B width) on the host with small random weights."""
    import torch

    from oracle import unet_torch as O

    with torch.device("meta"):
        m = O.UNetSpatioTemporalConditionModel()
    m = m.to_empty(device="cpu")
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.fill_(1.0)
            elif "norm" in n or n.endswith("bias"):
                p.zero_()
            elif n.endswith("mix_factor"):
                p.fill_(0.5)
            else:
                p.uniform_(-0.02, 0.02, generator=g)
    return m.eval()


def _use_sdpa(model, chunk_frames=0):
    """Route the oracle's self-attention through F.scaled_dot_product_attention — what diffusers' AttnProcessor2_0
    (the reference's attention) calls — instead of the oracle's materialised softmax(QK^T)V; same
    arithmetic, O(S) memory.  chunk_frames > 0 additionally evaluates it a few frames at a time (host memory)."""
    import torch
    import torch.nn.functional as F

    from oracle import unet_torch as O

    for name, mod in model.named_modules():
        if isinstance(mod, O.Attention):

            def fwd(x, context=None, mod=mod):
                ctx = x if context is None else context
                b, n, _ = x.shape

                def one(xs, cs):
                    bb = xs.shape[0]
                    q = mod.to_q(xs).view(bb, n, mod.heads, -1).transpose(1, 2)
                    k = mod.to_k(cs).view(bb, cs.shape[1], mod.heads, -1).transpose(1, 2)
                    v = mod.to_v(cs).view(bb, cs.shape[1], mod.heads, -1).transpose(1, 2)
                    o = F.scaled_dot_product_attention(q, k, v)
                    return mod.to_out[0](o.transpose(1, 2).reshape(bb, n, -1))

                if chunk_frames and b > chunk_frames and n > 1024:
                    return torch.cat([one(x[i:i + chunk_frames], ctx[i:i + chunk_frames]) for i in range(0, b, chunk_frames)])
                return one(x, ctx)

            mod.forward = fwd
    return model


def _time_cpu_steps(T, h, w, steps, warmup):
    import torch

    from oracle import unet_torch as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = _use_sdpa(_cpu_oracle(), chunk_frames=4)
    lat, cond, ehs, ids = make_inputs(T, h, w, None, 0)
    sig = O.karras_sigmas(25)
    guid = torch.linspace(1.0, 3.0, T).view(1, T, 1, 1, 1)
    with torch.no_grad():
        for i in range(warmup):
            O.denoise_step(m, lat, cond, float(sig[0]), float(sig[1]), ehs, ids, guid)
        t0 = time.perf_counter()
        x = lat
        for i in range(steps):
            x = O.denoise_step(m, x, cond, float(sig[i % 25]), float(sig[i % 25 + 1]), ehs, ids, guid)
        dt = time.perf_counter() - t0
    return dt, threads


def run_cpu(T, steps=1, warmup=0):
    """cpu_baseline of the default run: the reference's CPU path for one denoise step, restated by the fp32 PyTorch
    oracle (diffusers is not installable here), on a BOUNDED SAMPLE — the full-width UNet at 24x32 latents (1/12 of the
    pixels of config 2, a few seconds).  `value` is the sample's own measured rate scaled to 72x128 by the algorithmic
    FLOP ratio and is flagged as such (extrapolated / same_config false); the un-extrapolated measurement of the
    named config is what `bench.py --impl reference` times."""
    from evoworld_b200.unet import algorithmic_flops, DEFAULT_CONFIG

    h, w = 24, 32  # divisible by 8: three stride-2 stages must round-trip
    dt, threads = _time_cpu_steps(T, h, w, steps, warmup)
    cfg = dict(DEFAULT_CONFIG, **UNET_CFG)
    f_small = algorithmic_flops(cfg, 2, T, h, w)["total"]
    f_full = algorithmic_flops(cfg, 2, T, LAT_H, LAT_W)["total"]
    small_sps = steps / dt
    return {"value": small_sps * f_small / f_full, "unit": "steps/s", "cores": threads, "kind": "port",
            "extrapolated": True, "same_config": False,
            "sample": f"{steps} step(s) of the fp32 PyTorch oracle UNet (full 1.525B width) at {h}x{w} latents, T={T}: "
                      f"{dt:.1f} s, {f_small:.2f} TFLOP/step; value = measured rate scaled to 72x128 by the FLOP ratio "
                      f"{f_full / f_small:.1f} (context only; --impl reference times the real 72x128 step)",
            "seconds": dt, "measured_sample_steps_per_s": small_sps}


def run_reference(args):
    """Reference arm: the reference's own CPU path for the named config — one REAL denoise step at 72x128 latents,
    T frames, CFG batch 2, full-width fp32 UNet (oracle port: diffusers is not installable), all host threads.
    One step is ~90 TFLOP of fp32 CPU work (about a minute on 16 cores), so the arm times min(K, 2) steps without
    warm-up and reports exactly what it timed (`steps` = steps timed, `requested_steps` = K)."""
    h, w = getattr(args, "pano_height", 576) // 8, getattr(args, "pano_width", 1024) // 8
    steps = max(1, min(args.steps, int(os.environ.get("EVW_REFERENCE_MAX_STEPS", "2"))))
    dt, threads = _time_cpu_steps(args.frames, h, w, steps, 0)
    value = steps / dt
    cpu = {"value": value, "unit": "steps/s", "cores": threads, "kind": "port", "extrapolated": False, "same_config": True,
           "sample": f"{steps} real step(s) of the fp32 PyTorch oracle UNet at {h}x{w} latents, T={args.frames}, CFG batch 2 "
                     f"(the named config, no scaling): {dt:.1f} s", "seconds": dt}
    return {"metric": "denoise-steps/sec", "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": steps,
            "requested_steps": args.steps, "warmup": 0, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"config 2: single {8 * h}x{8 * w}x{args.frames}f clip, CFG batch 2, {h}x{w} latents, random-init 1.525B-param UNet, "
                                   f"Karras sigmas (25-step schedule)", "frames": args.frames,
                       "arm": "reference CPU path: fp32 PyTorch oracle UNet + CFG + Euler step on the host cores, the real "
                              "config (no extrapolation); min(K,2) steps timed, no warm-up"},
            "impl": "reference", "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_gpu_eager(T, h, w, dev, steps=3, warmup=3):
    """SURVEY §8(d) "reference PyTorch-CUDA" arm of config 2: the same PyTorch restatement of the reference's denoise step
    (oracle UNet with SDPA attention, as diffusers runs it) on THIS B200 through torch eager — cuDNN / cuBLAS / SDPA
    library kernels — in the three precisions a user of the reference can pick:
      fp32          weights + activations fp32, PyTorch defaults (cuDNN conv TF32 allowed, matmul fp32): how the
                    reference ships (unified_loop_consistency.py:188 keeps torch_dtype=float16 commented out);
      tf32          the same with torch.backends.cuda.matmul.allow_tf32 = True (train_evoworld.py:276 --allow_tf32);
      fp16_autocast torch.autocast(float16): tensor-core GEMMs/convs + FlashAttention SDPA — the fastest stock path.
    CUDA-event timed after `warmup` steps.  These are the library kernels the hand-written path replaces."""
    import torch

    from oracle import unet_torch as O

    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.manual_seed(0)
    with torch.device(dev):
        m = O.UNetSpatioTemporalConditionModel()
    m = _use_sdpa(m.eval())
    lat, cond, ehs, ids = [t.to(dev) for t in make_inputs(T, h, w, dev, seed=0)]
    sig = O.karras_sigmas(25)
    guid = torch.linspace(1.0, 3.0, T, device=dev).view(1, T, 1, 1, 1)
    out = {}

    def timed(name, ctx_factory):
        with torch.no_grad():
            x = lat.clone()
            for i in range(warmup):
                with ctx_factory():
                    x = O.denoise_step(m, x, cond, float(sig[i]), float(sig[i + 1]), ehs, ids, guid).float()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x = lat.clone()
            e0.record()
            for i in range(steps):
                with ctx_factory():
                    x = O.denoise_step(m, x, cond, float(sig[i]), float(sig[i + 1]), ehs, ids, guid).float()
            e1.record()
            torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"ms_per_step": ms, "steps_per_s": 1e3 / ms, "finite": bool(torch.isfinite(x).all())}

    import contextlib

    try:
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = False
        timed("fp32", contextlib.nullcontext)
        torch.backends.cuda.matmul.allow_tf32 = True
        timed("tf32", contextlib.nullcontext)
        timed("fp16_autocast", lambda: torch.autocast("cuda", dtype=torch.float16))
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del m
    torch.cuda.empty_cache()
    out["how"] = (f"oracle UNet (PyTorch restatement of the reference's diffusers blocks, SDPA attention) + CFG + Euler step, torch "
                  f"{torch.__version__} eager on this GPU; {warmup} warm-up + {steps} timed steps per mode, CUDA events")
    return out
