"""ReferenceNet writer: host-side mirror of magicanimate/models/appearance_encoder.py (AppearanceEncoderModel :126-1066).

The reference class is diffusers' 2-D `UNet2DConditionModel` (SD-1.5 layout) with the tail cut off: no
`conv_norm_out` / `conv_out`, and the last transformer block (`up_blocks.3.attentions.2`) reduced to its `norm1`
(:613-621) — nothing after that LayerNorm feeds a reference bank, and the only caller discards the returned sample
(`EMOAnimationPipeline.py:711-716`).  A 2-D SD UNet is exactly the 3-D one with a single frame and no motion modules
(that identity is what `UNet3DConditionModel.from_pretrained_2d`, unet_controlnet.py:485-525, relies on), so this class
reuses the UNet3D block containers and CUDA kernels with F = 1 and keeps the 2-D checkpoint's state_dict keys.

What it produces is the side effect the reference wants: in `write` mode (`ReferenceAttentionControl(encoder,
mode="write")`, mutual_self_attention.py:226-232) every mid / up `BasicTransformerBlock` appends LayerNorm1(x)
`[B, HW, C]` to its `.bank`; `ReferenceAttentionControl.update(writer)` hands those to the UNet3D reader blocks.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch

from ._lib import EmoteKernelError
from . import ops
from .unet3d import UNet3DConditionModel

_BLOCK_2D_TO_3D = {"CrossAttnDownBlock2D": "CrossAttnDownBlock3D", "DownBlock2D": "DownBlock3D",
                   "CrossAttnUpBlock2D": "CrossAttnUpBlock3D", "UpBlock2D": "UpBlock3D",
                   "UNetMidBlock2DCrossAttn": "UNetMidBlock3DCrossAttn"}


# constructor keywords of the reference (appearance_encoder.py:243-270) that select code paths outside SD-1.5: accepted
# at their default value only
_COLD_DEFAULTS = dict(transformer_layers_per_block=1, encoder_hid_dim=None, encoder_hid_dim_type=None,
                      num_attention_heads=None, addition_embed_type=None, addition_time_embed_dim=None,
                      resnet_skip_time_act=False, resnet_out_scale_factor=1.0, time_embedding_type="positional",
                      time_embedding_dim=None, time_embedding_act_fn=None, timestep_post_act=None, time_cond_proj_dim=None,
                      conv_in_kernel=3, conv_out_kernel=3, projection_class_embeddings_input_dim=None,
                      attention_type="default", class_embeddings_concat=False, mid_block_only_cross_attention=None,
                      cross_attention_norm=None, addition_embed_type_num_heads=64)


@dataclass
class UNet2DConditionOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


class AppearanceEncoderModel(UNet3DConditionModel):
    """appearance_encoder.py:126-1066.  Constructor keywords follow the reference (2-D block type names); options that
    select other diffusers sub-architectures are cold paths and raise."""

    def __init__(self, sample_size: Optional[int] = None, in_channels: int = 4, out_channels: int = 4,
                 center_input_sample: bool = False, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 down_block_types: Tuple[str] = ("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                 mid_block_type: Optional[str] = "UNetMidBlock2DCrossAttn",
                 up_block_types: Tuple[str] = ("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
                 only_cross_attention: Union[bool, Tuple[bool]] = False,
                 block_out_channels: Tuple[int] = (320, 640, 1280, 1280), layers_per_block: int = 2,
                 downsample_padding: int = 1, mid_block_scale_factor: float = 1, act_fn: str = "silu",
                 norm_num_groups: Optional[int] = 32, norm_eps: float = 1e-5, cross_attention_dim: int = 1280,
                 attention_head_dim: Union[int, Tuple[int]] = 8, dual_cross_attention: bool = False,
                 use_linear_projection: bool = False, class_embed_type: Optional[str] = None,
                 num_class_embeds: Optional[int] = None, upcast_attention: bool = False,
                 resnet_time_scale_shift: str = "default", **cold):
        unknown = sorted(set(cold) - set(_COLD_DEFAULTS))
        if unknown:
            raise TypeError(f"AppearanceEncoderModel: unexpected keyword arguments {unknown}")
        unsupported = sorted(k for k, v in cold.items() if v != _COLD_DEFAULTS[k])
        if unsupported:
            raise NotImplementedError(f"AppearanceEncoderModel: options {unsupported} are cold paths of the reference "
                                      "(other diffusers sub-architectures) and not implemented")
        try:
            down3 = tuple(_BLOCK_2D_TO_3D[t] for t in down_block_types)
            up3 = tuple(_BLOCK_2D_TO_3D[t] for t in up_block_types)
            mid3 = _BLOCK_2D_TO_3D[mid_block_type]
        except KeyError as e:
            raise NotImplementedError(f"AppearanceEncoderModel: block type {e} is not implemented") from e
        if up3[-1] != "CrossAttnUpBlock3D":
            raise NotImplementedError("AppearanceEncoderModel: the trimmed tail expects a final CrossAttnUpBlock2D")
        super().__init__(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                         center_input_sample=center_input_sample, flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift,
                         down_block_types=down3, mid_block_type=mid3, up_block_types=up3,
                         only_cross_attention=only_cross_attention, block_out_channels=block_out_channels,
                         layers_per_block=layers_per_block, downsample_padding=downsample_padding,
                         mid_block_scale_factor=mid_block_scale_factor, act_fn=act_fn, norm_num_groups=norm_num_groups,
                         norm_eps=norm_eps, cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
                         dual_cross_attention=dual_cross_attention, use_linear_projection=use_linear_projection,
                         class_embed_type=class_embed_type, num_class_embeds=num_class_embeds,
                         upcast_attention=upcast_attention, resnet_time_scale_shift=resnet_time_scale_shift,
                         use_motion_module=False, unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)
        self.config.update(down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
                           mid_block_type=mid_block_type)
        # appearance_encoder.py: no conv_norm_out / conv_out; :613-621 strips the last transformer block to its norm1
        del self.conv_norm_out, self.conv_act, self.conv_out
        tail = self.up_blocks[-1].attentions[-1]
        blk = tail.transformer_blocks[0]
        blk.attn1, blk.attn2, blk.norm2, blk.norm3, blk.ff = None, None, None, None, None
        tail.proj_out = None

    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                class_labels=None, timestep_cond=None, attention_mask=None, cross_attention_kwargs=None,
                added_cond_kwargs=None, down_block_additional_residuals=None, mid_block_additional_residual=None,
                encoder_attention_mask=None, return_dict: bool = True):
        """appearance_encoder.py:777-1066.  sample [B, 4, H, W].  Returns the hidden states after the last resnet
        ([B, C0, H, W]); the reference returns that tensor pushed through its identity-stubbed tail, a by-product its
        only caller discards — the banks written on the way are the result."""
        if not sample.is_cuda:
            raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
        for name, v in (("class_labels", class_labels), ("timestep_cond", timestep_cond), ("attention_mask", attention_mask),
                        ("cross_attention_kwargs", cross_attention_kwargs), ("added_cond_kwargs", added_cond_kwargs),
                        ("down_block_additional_residuals", down_block_additional_residuals),
                        ("mid_block_additional_residual", mid_block_additional_residual),
                        ("encoder_attention_mask", encoder_attention_mask)):
            if v is not None:
                raise NotImplementedError(f"AppearanceEncoderModel.forward: `{name}` is a cold path and not implemented")
        if sample.dim() != 4:
            raise ValueError(f"AppearanceEncoderModel expects [B, C, H, W], got {tuple(sample.shape)}")
        if any(s % (2 ** self.num_upsamplers) != 0 for s in sample.shape[-2:]):
            raise NotImplementedError("sample height/width must be multiples of 2**num_upsamplers")
        in_dtype = sample.dtype
        x = sample.float()[:, :, None]                                   # one frame
        if self.config.center_input_sample:
            x = 2 * x - 1.0
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.float32, device=x.device)
        timesteps = timesteps.reshape(-1).to(device=x.device, dtype=torch.float32).expand(x.shape[0]).contiguous()
        emb = self.time_embedding(self.time_proj(timesteps))
        emb._emote_silu_bf16 = ops.silu_bf16(emb)
        x = self.conv_in(x.contiguous())
        down_res = (x,)
        for blk in self.down_blocks:
            x, res = blk(hidden_states=x, temb=emb, encoder_hidden_states=encoder_hidden_states)
            down_res += res
        x = self.mid_block(x, emb, encoder_hidden_states=encoder_hidden_states)
        for blk in self.up_blocks:
            k = len(blk.resnets)
            res, down_res = down_res[-k:], down_res[:-k]
            x = blk(hidden_states=x, temb=emb, res_hidden_states_tuple=res, encoder_hidden_states=encoder_hidden_states)
        out = x[:, :, 0].contiguous(memory_format=torch.contiguous_format)
        if in_dtype != torch.float32:
            out = out.to(in_dtype)
        return UNet2DConditionOutput(sample=out) if return_dict else (out,)
