"""Shadows evoworld/reprojection/reproject_vggt_open3d_utils.py: GPU point splat instead of Open3D."""
from evoworld_b200.reprojection import (  # noqa: F401
    CUBEMAP, CUBEMAP_TRANSFORMS, CubemapRenderer, PointCloudProcessor, PointScene, SceneBuilder,
    align_first_and_last_points, predictions_to_target_view, rotation_from_vectors,
)
