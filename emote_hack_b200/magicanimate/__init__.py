"""Drop-in namespace mirroring the reference's `magicanimate` package for the hot path: put
`emote_hack_b200` on sys.path ahead of the reference (or alias `sys.modules['magicanimate']` to this package) and
`from magicanimate.models.unet_controlnet import UNet3DConditionModel` resolves to the B200 implementation."""
