"""Stands in for pyequilib's `equilib` package: EvoWorld only uses Equi2Pers (uint8, bilinear)."""
from evoworld_b200.equi2pers import Equi2Pers  # noqa: F401
