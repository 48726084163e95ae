"""Shadows utils/plucker_embedding.py: ray grid on the host, Plücker construction on the GPU."""
from evoworld_b200.plucker import equirectangular_to_ray, ray_c2w_to_plucker  # noqa: F401
