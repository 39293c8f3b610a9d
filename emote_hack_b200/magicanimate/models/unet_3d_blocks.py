"""Mirror of reference magicanimate/models/unet_3d_blocks.py."""
from ...unet3d import (CrossAttnDownBlock3D, CrossAttnUpBlock3D, DownBlock3D, UNetMidBlock3DCrossAttn, UpBlock3D,  # noqa: F401
                       get_down_block, get_up_block)
