"""Mirror of reference magicanimate/models/resnet.py."""
from ...unet3d import Downsample3D, InflatedConv3d, ResnetBlock3D, Upsample3D  # noqa: F401
