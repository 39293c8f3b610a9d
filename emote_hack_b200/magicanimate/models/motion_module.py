"""Mirror of reference magicanimate/models/motion_module.py."""
from ...unet3d import (PositionalEncoding, TemporalTransformer3DModel, TemporalTransformerBlock, VanillaTemporalModule,  # noqa: F401
                       VersatileAttention, get_motion_module, zero_module)
