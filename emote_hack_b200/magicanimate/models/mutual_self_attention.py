"""Mirror of reference magicanimate/models/mutual_self_attention.py (ReferenceAttentionControl :128-641)."""
from ...unet3d import ReferenceAttentionControl, torch_dfs  # noqa: F401
