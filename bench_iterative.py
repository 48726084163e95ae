"""Iterative arm of bench.py (BASELINE configs 3 / 5): the 3-clip loop of unified_loop_consistency.process_episode
(unified_loop_consistency.py:398-492) with everything device-resident.

Per scene (one per rank), for segment 0, 1, 2:
  1. generate      `steps` fused denoise steps (evw_denoise_step) of a T-frame clip, CFG batch 2, then the      [built: hot path 1]
                   NCCL all-gather of the clip's latents over the ranks (distributed.gather_latents)
  2. decode        VAE temporal decode of the clip's latents -> T panoramas, decode_chunk_size 8                 [built: evw_vae_decode,
                   random-init 97.7 M-parameter AutoencoderKLTemporalDecoder] -> uint8 [T,3,H,W]
  and, when another segment follows (:442-485):
  3. pano -> pers  Equi2Pers(384, 512, fov 90) with the look-at yaw of every frame                           [built: evw_equi2pers_u8]
                   reference: ALL frames so far (25, then 49); incremental: only the segment's new frames
  4. VGGT-1B       Pillow-exact bicubic 384x512 -> 392x518 of EVERY frame so far (25, then 49: global attention couples  [built: evoworld_b200.vggt,
                   all frames, so nothing is incremental here), then the 1.19 B-parameter network (random init)          random init]
                   -> depth / confidence / pose.  With random weights these predictions describe no scene, so the
                   geometry handed to stages 5-8 stays the seeded synthetic prediction set; the network's
                   time is inside the timed region and its outputs are checked finite
  5. lift + pack   new frames appended to the device-resident PointMemory (evw_lift_pack_points)             [built]
  6. filter        joint 50th-percentile confidence filter + compaction (evw_conf_select)                    [built]
  7. align         similarity alignment of the GT trajectory (host numpy float64, 24 poses)                  [built: host]
  8. splat         24 target views: cube splat + cube->equirect resolve (evw_splat_cube_equirect)            [built: hot path 2]
  9. memory frames 24 panoramas 1000x2000 -> 576x1024, bit-exact PIL bilinear Resize (evw_resize_pil_u8), [-1,1]  [built]
 10. VAE encode    memory latents of the next clip: first frame + 24 reprojections                           [built: evw_vae_encode]
`mode="reference"` re-warps and re-lifts every frame generated so far each segment, as the reference does;
`mode="incremental"` (default) only touches the new frames — the scene the splat sees is bit-identical
(tests/test_gpu_reproj.py::test_point_memory_incremental_equals_one_shot).
"""
from __future__ import annotations

import math
import time

import numpy as np


def run_ours(args, dev, rank, world, barrier, allreduce_max, peaks):
    import torch
    import torch.nn.functional as F

    import bench_denoise as bd
    from evoworld_b200 import reprojection as R
    from evoworld_b200 import image_ops, segments, synthetic
    from evoworld_b200.distributed import gather_latents
    from evoworld_b200.equi2pers import Equi2Pers
    from evoworld_b200.memory import PointMemory
    from evoworld_b200.scheduler import EulerDiscreteScheduler
    from evoworld_b200.unet import UNetSpatioTemporalConditionModel
    from evoworld_b200.vae import AutoencoderKLTemporalDecoder
    from evoworld_b200.vggt import VGGT, DEFAULT_CONFIG as VGGT_CFG, random_state_dict as vggt_random, run_vggt_inference

    T = args.iter_frames
    steps = args.iter_steps
    H, W = args.pano_height, args.pano_width
    h, w = H // 8, W // 8
    n_seg = 3
    mode = args.iter_mode
    unet = UNetSpatioTemporalConditionModel(**bd.UNET_CFG).init_random(seed=0, device=dev)
    unet._ensure_handle()
    unet.free_master_parameters()
    vae = AutoencoderKLTemporalDecoder().init_random(seed=1, device=dev)
    vae.free_master_parameters()
    sf = vae.config.scaling_factor
    lat0, cond0, ehs, ids = [t.to(dev) for t in bd.make_inputs(T, h, w, dev, seed=rank)]
    sched = EulerDiscreteScheduler()
    sched.set_timesteps(steps)
    sig = [float(s) for s in sched.sigmas]
    vggt_net = None
    if args.iter_vggt:
        vcfg = dict(VGGT_CFG, point_head=False)   # the loop reads depth / depth_conf / pose_enc only (unified_loop_consistency.py:352-366)
        vggt_net = VGGT(**vcfg).to(dev)
        vggt_net.load_state_dict(vggt_random(vcfg, seed=2, device=dev))
        vggt_net.free_master_parameters()
    vggt_ok = []
    # geometry of the scene (random VGGT weights describe none): seeded synthetic predictions, resident before the clock starts
    frames_u8 = torch.empty((T, 3, H, W), dtype=torch.uint8, device=dev)                         # decoded panoramas of a clip
    S_all = (n_seg - 1) * (T - 1) + 1                                                              # 49 frames after two segments
    p = synthetic.reprojection_predictions(S=S_all, H=392, W=518, seed=rank)
    vggt = {k: torch.from_numpy(p[k]).to(dev) for k in ("depth", "depth_conf", "images", "extrinsic", "intrinsic")}
    poses = synthetic.curve_trajectory().astype(np.float64)                                         # [126, 6] RDF, degrees
    e2p = Equi2Pers(height=384, width=512, fov_x=90, mode="bilinear", device=dev)
    mem = PointMemory(392, 518, capacity_frames=S_all, device=dev)
    sb = R.SceneBuilder(dev)
    G = args.views_per_pass
    zbuf = torch.empty(R.splat_workspace_bytes(G, 512), dtype=torch.uint8, device=dev)
    panos = torch.empty((24, 1000, 2000, 3), dtype=torch.uint8, device=dev)
    all_frames = torch.empty((n_seg * (T - 1) + 1, 3, H, W), dtype=torch.uint8, device=dev)     # every frame of the episode
    ev = lambda: torch.cuda.Event(enable_timing=True)

    # persistent device buffers: the captured graph of the denoise plan is keyed on the latents / conditioning addresses
    x = lat0.clone()
    cond = cond0.clone()

    def episode(timers=None, nsteps=None):
        nsteps = steps if nsteps is None else nsteps
        mem.reset()
        cond.copy_(cond0)
        n_frames = 0
        stage = {}

        def mark(name, e0, e1):
            if timers is not None:
                timers.setdefault(name, []).append((e0, e1))

        for seg in range(n_seg):
            e0 = ev(); e0.record()
            x.copy_(lat0)
            for i in range(nsteps):
                unet.denoise_step(x, cond, sig[i], sig[i + 1], ehs, ids, 1.0, 3.0)
            gathered = gather_latents(x)  # clip boundary: the only collective of the path (no-op at world == 1)
            e1 = ev(); e1.record(); mark("denoise", e0, e1)
            # decode_latents (pipeline_evoworld.py:358-385) with decode_chunk_size 8, then [-1,1] -> uint8 panoramas
            z = x[0] / sf
            for i in range(0, T, 8):
                fr = vae.decode(z[i:i + 8], num_frames=z[i:i + 8].shape[0]).sample
                frames_u8[i:i + 8] = (fr * 127.5 + 127.5).clamp_(0, 255).to(torch.uint8)
            e1b = ev(); e1b.record(); mark("vae decode", e1, e1b)
            # the first frame of a later segment repeats the previous last one
            new = frames_u8 if seg == 0 else frames_u8[1:]
            all_frames[n_frames:n_frames + new.shape[0]].copy_(new)
            first_new, n_frames = n_frames, n_frames + new.shape[0]
            if seg == n_seg - 1:
                break
            e2 = ev(); e2.record()
            _, _, look_at = segments.calculate_segment_indices(seg)
            lo = 0 if mode == "reference" else first_new
            rots = [{"pitch": 0.0, "roll": 0.0, "yaw": segments.calculate_target_yaw(poses, i + 1, look_at)} for i in range(lo, n_frames)]
            pers = e2p(all_frames[lo:n_frames], rots)
            e3 = ev(); e3.record(); mark("equi2pers", e2, e3)
            if vggt_net is not None:
                if lo != 0:   # VGGT sees every frame generated so far, re-warped towards this segment's look-at frame (:444-462)
                    rots_all = [{"pitch": 0.0, "roll": 0.0, "yaw": segments.calculate_target_yaw(poses, i + 1, look_at)} for i in range(n_frames)]
                    pers_all = e2p(all_frames[0:n_frames], rots_all)
                else:
                    pers_all = pers
                vp = run_vggt_inference(vggt_net, pers_all.permute(0, 2, 3, 1).contiguous())
                vggt_ok.append(torch.isfinite(vp["depth"]).all() & torch.isfinite(vp["pose_enc"]).all() & torch.isfinite(vp["depth_conf"]).all())
                e3b = ev(); e3b.record(); mark("vggt", e3, e3b)
                e3 = e3b
            if mode == "reference":
                mem.reset()
                sl = slice(0, n_frames)
            else:
                sl = slice(first_new, n_frames)
            mem.append(vggt["depth"][sl], vggt["depth_conf"][sl], vggt["images"][sl], vggt["extrinsic"][sl], vggt["intrinsic"][sl])
            scene = mem.scene(50.0)
            e4 = ev(); e4.record(); mark("lift+filter", e3, e4)
            tgt = sb.align_extrinsics(p["camera_pose"], p["extrinsic"][:n_frames], 24, f"bench_{seg}", False)
            w2c = torch.from_numpy(R.front_w2c_matrices(tgt)).to(dev, non_blocking=True)
            R.splat_to_panoramas_device(scene, w2c, 2000, 1000, 512, G, out=panos, zbuf=zbuf)
            e5 = ev(); e5.record(); mark("splat", e4, e5)
            # memory frames of the next clip: slot 0 = the episode's first frame, slots 1..24 = the reprojections
            # transforms.Resize((H, W)) + ToTensor on the PIL copy of every panorama (CameraTrajDataset.py:597-600): Pillow-exact
            m = image_ops.resize_pil_u8(panos, H, W).permute(0, 3, 1, 2).float() / 127.5 - 1.0
            first = all_frames[0:1].float() / 127.5 - 1.0
            mem_frames = torch.cat([first, m], dim=0)[:T]                                           # [T,3,H,W]
            e6 = ev(); e6.record(); mark("memory frames", e5, e6)
            cond[1, :, 4:8] = vae.encode(mem_frames).latent_dist.mode()                              # :610-617
            e7 = ev(); e7.record(); mark("vae encode", e6, e7)
            stage["points"] = stage.get("points", []) + [scene]
        return x, stage

    # warm-up: one full episode (plans, graphs, allocator)
    # (--iter-warmup-steps: fewer denoise steps per clip in the warm-up episode; every kernel, plan and graph is still exercised)
    for _ in range(max(1, min(args.warmup, 1))):
        episode(nsteps=min(steps, args.iter_warmup_steps) if args.iter_warmup_steps > 0 else None)
    torch.cuda.synchronize(dev)
    barrier(world)
    n_ep = max(1, args.iter_episodes)
    timers = {}
    e_start, e_end = ev(), ev()
    t0 = time.perf_counter()
    e_start.record()
    for _ in range(n_ep):
        x, stage = episode(timers)
    e_end.record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    barrier(world)
    ms = e_start.elapsed_time(e_end)
    ms = allreduce_max(ms, dev, world)
    clips = n_ep * n_seg
    per_stage = {k: sum(a.elapsed_time(b) for a, b in v) / n_ep for k, v in timers.items()}
    pts = [int(s.num_points()) for s in stage["points"]]
    # our kernels per episode: the denoise plan per step + the VAE plans per chunk (reprojection kernels: a few dozen, not counted)
    unet_launches, _ = unet.plan_info()
    dec_launches = vae.plan_info(1)[0]
    enc_launches = vae.plan_info(0)[0]
    launches = n_seg * steps * unet_launches + n_seg * -(-T // 8) * dec_launches + (n_seg - 1) * -(-T // 8) * enc_launches
    vggt_forward = None
    if vggt_net is not None:   # the network alone on one segment's frames (depth head only), 3 timed forwards after a warm one
        vi = torch.rand((1, T, 3, 392, 518), device=dev)
        vggt_net(vi)
        ve = [ev() for _ in range(4)]
        ve[0].record()
        for i in range(3):
            vggt_net(vi)
            ve[i + 1].record()
        torch.cuda.synchronize(dev)
        vms = ve[0].elapsed_time(ve[3]) / 3
        P_tok = 28 * 37 + 5
        # executed FLOPs of the attention (4 S^2 d per head and sequence): 24 DINOv2 + 24 frame blocks over T sequences of P_tok
        # tokens, 24 global blocks over one sequence of T * P_tok tokens; GEMM FLOPs counted by tools/vggt_bench.py (60.0 TFLOP at T = 25)
        att = 4.0 * 64 * 16 * (48 * T * P_tok ** 2 + 24 * (T * P_tok) ** 2) / 1e12
        vggt_forward = {"frames": T, "size": [392, 518], "ms": vms, "frames_per_s": T / vms * 1e3, "attention_tflop": att,
                        "note": "VGGT-1B (random init, depth + camera heads), device-resident input; see profiles/r02am_vggt_bench_S25.json "
                                "for the FLOP count and the PyTorch-eager legs"}
    return {
        "metric": "iterative clips/sec (3-clip episode, evolving point memory)", "unit": "clips/s",
        "value": world * clips / (ms * 1e-3), "ms_per_episode": ms / n_ep, "episodes": n_ep, "wall_s": wall,
        "ms_per_stage_per_episode": per_stage, "memory_points_per_segment": pts, "launches_per_episode": int(launches), "vggt_forward": vggt_forward, "finite_output": bool(torch.isfinite(x).all()),
        "config": {"workload": f"config 3: 3-clip iterative episode, {H}x{W}x{T}f clips, {steps} denoise steps per clip, CFG batch 2, "
                               f"evolving point memory {pts} points (S = {T}, {S_all} frames), 24 target views per segment",
                   "mode": mode, "scenes_per_rank": 1,
                   "vae": "native AutoencoderKLTemporalDecoder (random init): temporal decode of every clip in chunks of 8, "
                          "encode of the 25 memory frames of the next clip",
                   "vggt": ("native VGGT-1B (1.19 B parameters, random init, point head off) on every frame generated so far, inside the "
                            "timed region; outputs finite: " + str(bool(all(bool(t) for t in vggt_ok))) if vggt_net is not None else "off (--no-iter-vggt)"),
                   "stand_ins": "random VGGT weights describe no scene: the geometry consumed by lift / filter / splat is the seeded "
                                "synthetic prediction set resident on the device (see bench_iterative.py)"},
    }
