from . import (attention, embeddings, motion_module, mutual_self_attention, orig_attention, resnet, unet,  # noqa: F401
               unet_3d_blocks, unet_controlnet)
