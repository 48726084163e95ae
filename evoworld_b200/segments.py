"""Segment / look-at bookkeeping of the iterative loop (host side).

  calculate_segment_indices   evoworld/reprojection/pano_to_pers_utils.py:5-14
  calculate_target_yaw        unified_loop_consistency.py:308-323 (== pano_to_pers_per_segment.py:77-86)
  read_camera_file_and_convert_to_rdf / write_camera_file   pano_to_pers_utils.py:17-34
Pure Python float64 so that the values written back to the camera file are bit-identical.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

UNITY_TO_OPENCV = [1, -1, 1, -1, 1, -1]


def calculate_segment_indices(segment_id: int) -> Tuple[int, int, int]:
    """(start_idx, end_idx, look_at_idx): 25-frame segments overlapping by one frame."""
    look_at_idx = (segment_id + 1) * 24 + 24
    start_idx = segment_id * 24 + 1
    if segment_id == 0:
        start_idx = start_idx - 1
    return start_idx, start_idx + 25, look_at_idx


def calculate_target_yaw(camera_params: np.ndarray, current_idx: int, look_at_idx: int) -> float:
    """yaw_diff (radians) = rad(yaw_i) - atan2(x_L - x_i, z_L - z_i); current_idx is 1-based."""
    if current_idx > len(camera_params):
        return 0.0
    cur = camera_params[current_idx - 1]
    look = camera_params[min(look_at_idx, len(camera_params) - 1)]
    target = math.atan2(look[0] - cur[0], look[2] - cur[2])
    return math.radians(cur[4]) - target


def read_camera_file_and_convert_to_rdf(camera_file: str) -> np.ndarray:
    with open(camera_file, "r") as f:
        lines = f.readlines()
    params = np.array([list(map(float, line.strip().split(",")))[1:] for line in lines[1:]])
    return params * UNITY_TO_OPENCV


def write_camera_file(camera_params: np.ndarray, output_camera_file: str):
    with open(output_camera_file, "w") as f:
        for i in range(len(camera_params)):
            p = camera_params[i]
            f.write(f"{i+1} {p[0]} {p[1]} {p[2]} {p[3]} {p[4]} {p[5]}\n")
