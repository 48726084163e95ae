"""evoworld_b200 — B200-native (sm_100a) hot paths of EvoWorld behind the reference's call surface.

Two paths (BASELINE.json north_star):
  1. the SVD spatio-temporal UNet denoise step  (evoworld_b200.unet / .pipeline)
  2. the 3D-memory reprojection path            (evoworld_b200.plucker / .equi2pers / .lift / .reprojection)

All compute goes through the C-ABI shared library `libevoworld_b200.so` (include/evoworld_b200.h);
there is no CPU fallback: importing a compute function without the built library raises.
"""
__version__ = "0.1.0"
