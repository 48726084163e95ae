"""Shadows evoworld/trainer/unet_plucker.py (and unet.py): the UNet forward runs on evoworld_b200."""
from evoworld_b200.unet import UNetSpatioTemporalConditionModel, UNetSpatioTemporalConditionOutput  # noqa: F401
