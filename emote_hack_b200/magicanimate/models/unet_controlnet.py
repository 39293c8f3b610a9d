"""Mirror of reference magicanimate/models/unet_controlnet.py (UNet3DConditionModel :54-525)."""
from ...unet3d import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401
