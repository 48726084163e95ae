"""Single-clip caller of the pipeline (the row next to the hot path).

Mirrors evoworld/inference/forward_evoworld.py:119-211 — `prepare_batch_data` (pose rows -> relative 3x4 c2w ->
Plücker embedding per clip), `process_batch` (one pipeline call with the reference's fixed arguments, then PNGs) and
`save_frames` — with the same names, argument order and assertions, on top of evoworld_b200's own pose / Plücker /
pipeline implementations.  Everything stays on the device between the dataset batch and the pipeline call.
"""
from __future__ import annotations

import os

import torch

from .geometry import xyz_euler_to_three_by_four_matrix_batch
from .plucker import ray_c2w_to_plucker


def prepare_batch_data(batch, args, rays, weight_dtype, device="cuda"):
    """batch: dict with pixel_values [B,F,3,H,W] in [-1,1], cam_traj [B,F,6] (x,y,z,rx,ry,rz; degrees),
    memorized_pixel_values [B,F,3,H,W].  Returns (first_frame, camera_traj [B,T,3,4], plucker_embedding
    [B,T,6,H/8,W/8], memorized_pixel_values, images) as forward_evoworld.py:119-156."""
    images = batch["pixel_values"]
    first_frame = images[:, 0].to(device)
    camera_traj_raw = batch["cam_traj"].to(device)
    n = camera_traj_raw.shape[0]
    camera_traj = torch.zeros(n, args.num_frames, 3, 4, dtype=weight_dtype, device=device)
    plucker_embedding = torch.zeros(n, args.num_frames, 6, args.height // 8, args.width // 8, dtype=weight_dtype, device=device)
    for i in range(n):
        camera_traj[i] = xyz_euler_to_three_by_four_matrix_batch(camera_traj_raw[i], relative=True)
        plucker_embedding[i] = ray_c2w_to_plucker(rays, camera_traj[i])
    memorized_pixel_values = batch["memorized_pixel_values"].to(device)
    return first_frame, camera_traj, plucker_embedding, memorized_pixel_values, images


def save_frames(video_frames, gt_frames, frames_path: str, frames_gt_path: str, num_frames: int):
    """PNG dump of predictions and ground truth (forward_evoworld.py:159-180): `{i+1:03}.png`."""
    from PIL import Image

    os.makedirs(frames_path, exist_ok=True)
    os.makedirs(frames_gt_path, exist_ok=True)
    assert len(video_frames) == num_frames, f"video frames {len(video_frames)} should equal num_frames {num_frames}!"
    for i in range(num_frames):
        gt = gt_frames[i] * 0.5 + 0.5
        gt = Image.fromarray(gt.mul(255).byte().detach().cpu().numpy().transpose(1, 2, 0))
        video_frames[i].save(os.path.join(frames_path, f"{i + 1:03}.png"))
        gt.save(os.path.join(frames_gt_path, f"{i + 1:03}.png"))


def process_batch(batch, args, pipeline, rays, weight_dtype, output_path: str, episode: str):
    """One clip: prepare -> pipeline(...).frames[0] -> PNGs (forward_evoworld.py:183-211; the pipeline arguments are
    the reference's: decode_chunk_size 8, motion_bucket_id 127, fps 7, noise_aug_strength 0.02, no generator)."""
    first_frame, camera_traj, plucker_embedding, memorized_pixel_values, images = prepare_batch_data(batch, args, rays, weight_dtype)
    with torch.inference_mode():
        video_frames = pipeline(first_frame, height=args.height, width=args.width, num_frames=args.num_frames,
                                decode_chunk_size=8, motion_bucket_id=127, fps=7, noise_aug_strength=0.02,
                                plucker_embedding=plucker_embedding, memorized_pixel_values=memorized_pixel_values,
                                mask_mem=args.mask_mem).frames[0]
    frames_path = os.path.join(output_path, episode, "predictions")
    frames_gt_path = os.path.join(output_path, episode, "predictions_gt")
    save_frames(video_frames, images[0], frames_path, frames_gt_path, args.num_frames)
    return video_frames
