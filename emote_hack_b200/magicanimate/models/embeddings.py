"""Mirror of the hot-path classes of reference magicanimate/models/embeddings.py."""
from ...unet3d import TimestepEmbedding, Timesteps  # noqa: F401
