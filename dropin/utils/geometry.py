"""Shadows utils/geometry.py."""
from evoworld_b200.geometry import xyz_euler_to_four_by_four_matrix_batch, xyz_euler_to_three_by_four_matrix_batch  # noqa: F401
