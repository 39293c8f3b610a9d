"""Drop-in namespace mirroring the reference's `magicanimate` package for the hot path: alias
`sys.modules['magicanimate.models.<name>']` to `emote_hack_b200.magicanimate.models.<name>` and
`from magicanimate.models.unet_controlnet import UNet3DConditionModel` resolves to the B200 implementation."""
from . import models, pipelines, utils  # noqa: F401
