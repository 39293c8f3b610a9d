"""Mirror of reference magicanimate/models/attention.py."""
from ...unet3d import BasicTransformerBlock, CrossAttention, FeedForward, Transformer3DModel, Transformer3DModelOutput  # noqa: F401
