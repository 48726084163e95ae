"""The reference's import surface resolves to evoworld_b200 through dropin/ (CPU, import only)."""
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def test_import_surface(monkeypatch):
    monkeypatch.syspath_prepend(str(ROOT / "dropin"))
    for m in [k for k in sys.modules if k.split(".")[0] in ("evoworld", "utils", "equilib", "third_party")]:
        monkeypatch.delitem(sys.modules, m)
    surface = {
        "evoworld.pipeline.pipeline_evoworld": ["StableVideoDiffusionPipeline"],
        "evoworld.trainer.unet_plucker": ["UNetSpatioTemporalConditionModel"],
        "evoworld.reprojection.reproject_vggt_open3d_utils": ["PointCloudProcessor", "SceneBuilder", "CubemapRenderer",
                                                              "predictions_to_target_view"],
        "evoworld.reprojection.pano_to_pers_utils": ["calculate_segment_indices"],
        "utils.plucker_embedding": ["equirectangular_to_ray", "ray_c2w_to_plucker"],
        "utils.geometry": ["xyz_euler_to_four_by_four_matrix_batch"],
        "equilib": ["Equi2Pers"],
        "third_party.vggt.vggt.utils.geometry": ["unproject_depth_map_to_point_map"],
        "third_party.vggt.vggt.utils.pose_enc": ["pose_encoding_to_extri_intri"],
        "third_party.vggt.vggt.models.vggt": ["VGGT"],
        "third_party.vggt.vggt.utils.load_fn": ["load_and_preprocess_images"],
    }
    for mod, names in surface.items():
        m = importlib.import_module(mod)
        assert "dropin" in (m.__file__ or ""), mod
        for n in names:
            assert hasattr(m, n), (mod, n)
    # constructing the processors must not download models or open GL contexts
    ru = importlib.import_module("evoworld.reprojection.reproject_vggt_open3d_utils")
    ru.PointCloudProcessor(), ru.SceneBuilder(), ru.CubemapRenderer()


def test_segment_helpers(golden, tmp_path):
    from evoworld_b200 import segments as S

    np.testing.assert_array_equal(np.array([S.calculate_segment_indices(s) for s in range(4)]), golden["segment_indices"])
    poses = golden["poses_rdf"].astype(np.float64)
    y = S.calculate_target_yaw(poses, 1, 48)
    import math

    want = math.radians(poses[0][4]) - math.atan2(poses[48][0] - poses[0][0], poses[48][2] - poses[0][2])
    assert y == want
    assert S.calculate_target_yaw(poses, len(poses) + 1, 48) == 0.0
    f = tmp_path / "cam.txt"
    S.write_camera_file(poses[:3], str(f))
    assert len(f.read_text().strip().splitlines()) == 3
