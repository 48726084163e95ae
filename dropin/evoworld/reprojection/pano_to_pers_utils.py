"""Shadows evoworld/reprojection/pano_to_pers_utils.py (host-side index / camera-file helpers)."""
from evoworld_b200.segments import (  # noqa: F401
    UNITY_TO_OPENCV, calculate_segment_indices, calculate_target_yaw, read_camera_file_and_convert_to_rdf, write_camera_file,
)
