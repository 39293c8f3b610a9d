"""Mirror of reference magicanimate/models/unet.py (same network without the ControlNet residual inputs)."""
from ...unet3d import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401
