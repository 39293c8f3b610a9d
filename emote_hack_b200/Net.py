"""Reference import path `from Net import Wav2VecFeatureExtractor, SpeedEncoder` (EMOAnimationPipeline.py:62): the two
Net.py classes on the hot path's input side (Net.py:198-258, 607-797), implemented in emote_hack_b200/audio.py.
Alias with `sys.modules["Net"] = emote_hack_b200.Net`."""
from .audio import SpeedEncoder, Wav2VecFeatureExtractor  # noqa: F401
