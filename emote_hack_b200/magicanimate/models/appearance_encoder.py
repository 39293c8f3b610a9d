"""Mirror of reference magicanimate/models/appearance_encoder.py (AppearanceEncoderModel :126-1066)."""
from ...appearance_encoder import AppearanceEncoderModel, UNet2DConditionOutput  # noqa: F401
