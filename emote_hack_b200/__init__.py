"""emote_hack_b200 — B200-native (sm_100a) implementation of the Emote-hack denoising hot path.

Sub-modules:
  _lib, ops            ctypes binding of libemote_b200.so (include/emote_b200.h) and the tensor-level operator layer
  unet3d, vae, sampler host-side mirror of the reference's module interface for the path (built in later files)
"""
__version__ = "0.1.0"
