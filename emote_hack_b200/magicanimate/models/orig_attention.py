"""Mirror of the hot-path classes of reference magicanimate/models/orig_attention.py."""
from ...unet3d import GEGLU, CrossAttention, FeedForward  # noqa: F401
