"""Host-side mirror of the reference's model interface for the denoising hot path.

Same class names, constructor kwargs, parameter / buffer names (state_dict keys) and forward() signatures as
`magicanimate/models/{unet_controlnet,unet_3d_blocks,resnet,attention,motion_module,mutual_self_attention}.py`
(reference file:line cited per class), so SD-1.5 / AnimateDiff / MagicAnimate checkpoints load unchanged and the
modules drop into `EMOAnimationPipeline` — but every forward() here launches the sm_100a kernels of
libemote_b200.so through `ops` instead of ATen/cuDNN/cuBLAS.  nn.Linear / nn.Conv2d / nn.GroupNorm / nn.LayerNorm
children are parameter containers only (for key-name compatibility); their own forward() is never called on the
hot path.  There is no CPU fallback: calling a forward on CPU tensors raises.

Data layout: activations are fp32 [b, c, f, h, w] tensors stored channels-last (torch.channels_last_3d), i.e. a
row-major "tokens" matrix [(b f h w), c]; the fp32 residual stream is kept between kernels and rounded to bf16
only as tensor-core operands.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
from torch import nn

from . import ops
from ._lib import EmoteKernelError

F32, OP16 = torch.float32, ops.OP16

# The context K/V projections are cached per context tensor (constant over the denoising steps).  CUDA-graph capture
# switches the cache off so the projections are part of the captured step and replays stay correct when the static
# context buffer is refilled (pipeline.GraphedUNet).
CTX_KV_CACHE_ENABLED = True


# =============================================================================================== helpers
class AttrDict(dict):
    """config container with attribute access (`unet.config.in_channels`), like diffusers' FrozenDict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _tokens(x: torch.Tensor) -> Tuple[torch.Tensor, Tuple[int, int, int, int, int]]:
    """[b, c, f, h, w] (any layout/dtype) -> fp32 tokens view/copy [(b f h w), c] + dims."""
    if x.dim() != 5:
        raise ValueError(f"expected a 5-D [b, c, f, h, w] tensor, got shape {tuple(x.shape)}")
    if not x.is_cuda:
        raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
    b, c, f, h, w = x.shape
    rec = getattr(x, "_emote_tok", None)
    if rec is not None and rec[0]._version == rec[1] and x.dtype == F32 and rec[0].shape == (b * f * h * w, c):
        return rec[0], (b, c, f, h, w)   # the token tensor a previous module produced (keeps its fused GN statistics)
    if x.dtype != F32:
        x = x.float()
    xt = x.permute(0, 2, 3, 4, 1)
    if xt.is_contiguous():
        return xt.reshape(b * f * h * w, c), (b, c, f, h, w)
    return ops.ncfhw_to_tokens(x.contiguous()), (b, c, f, h, w)


def _untokens(tok: torch.Tensor, b: int, c: int, f: int, h: int, w: int) -> torch.Tensor:
    """tokens [(b f h w), c] -> [b, c, f, h, w] view with channels-last strides."""
    v = tok.view(b, f, h, w, c).permute(0, 4, 1, 2, 3)
    v._emote_tok = (tok, tok._version)
    return v


def _sig(*tensors) -> tuple:
    """identity + in-place version of the source parameters of a packed copy: changes on load_state_dict / copy_ / mul_
    (version bump) and on .to() / .cuda() (new storage)"""
    return tuple((t.data_ptr(), t._version) for t in tensors if t is not None)


class _PackedModule(nn.Module):
    """Caches kernel-layout (16-bit, packed) copies of the fp32 parameters; rebuilt whenever a source parameter (weight or
    bias) was rewritten in place or re-allocated since the copy was made (load_state_dict, .to(), param.copy_(), a LoRA
    merge ...)."""

    def __init__(self):
        super().__init__()
        self._pk: Optional[dict] = None
        self._pk_sig: Optional[tuple] = None

    def _pack(self) -> dict:
        raise NotImplementedError

    @property
    def pk(self) -> dict:
        sig = _sig(*self.parameters())
        if self._pk is None or self._pk_sig != sig:
            with torch.no_grad():
                self._pk = self._pack()
            self._pk_sig = sig
        return self._pk


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(F32).contiguous()


# =============================================================================================== resnet.py
class InflatedConv3d(nn.Conv2d):
    """resnet.py:30-38 — a 2-D convolution applied to every frame of [b, c, f, h, w]."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._pk = None

    def packed(self):
        sig = _sig(self.weight, self.bias)
        if self._pk is None or self._pk[2] != sig:
            ks = self.kernel_size[0]
            with torch.no_grad():
                if ks == 1:
                    w = ops.pack_linear(self.weight)
                elif self.in_channels <= 7:
                    w = ops.pack_conv3x3_small(self.weight)
                else:
                    w = ops.pack_conv3x3(self.weight)
                b = _f32c(self.bias) if self.bias is not None else None
            self._pk = (w, b, sig)
        return self._pk[0], self._pk[1]

    def forward_tokens(self, a_bf16: torch.Tensor, n_img: int, h: int, w: int, **epi) -> torch.Tensor:
        """bf16 NHWC operand -> conv output tokens (epilogue options forwarded to the GEMM)."""
        wp, bias = self.packed()
        ks, st = self.kernel_size[0], self.stride[0]
        if ks == 1:
            return ops.gemm(a_bf16, wp, bias=bias, **epi)
        if ks == 3 and st == 1 and self.padding[0] == 1:
            return ops.conv3x3(a_bf16, wp, n_img, h, w, self.in_channels, bias=bias, **epi)
        raise NotImplementedError("InflatedConv3d: only 1x1 and 3x3/pad-1 convolutions are on the hot path")

    def forward(self, x):
        if self.kernel_size[0] == 3 and self.in_channels <= 7:  # conv_in on raw latents
            if not x.is_cuda:
                raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
            b, c, f, h, w = x.shape
            wp, bias = self.packed()
            out = ops.gemm(ops.latent_im2col(x.float().contiguous()), wp, bias=bias,
                           stats_rows=ops.stats_rows_for(h * w, f * h * w))
            return _untokens(out, b, self.out_channels, f, h, w)
        tok, (b, c, f, h, w) = _tokens(x)
        wp, bias = self.packed()
        if self.kernel_size[0] == 3 and self.stride[0] == 2:
            cols = ops.im2col_s2(tok, b * f, h, w, c)
            ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            return _untokens(ops.gemm(cols, wp, bias=bias, stats_rows=ops.stats_rows_for(ho * wo, f * ho * wo)),
                             b, self.out_channels, f, ho, wo)
        out = self.forward_tokens(ops.cast_bf16(tok), b * f, h, w)
        return _untokens(out, b, self.out_channels, f, h, w)


class Upsample3D(nn.Module):
    """resnet.py:41-84 — nearest x2 on (h, w) then 3x3 conv; the upsample is a gather feeding the conv operand."""

    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv"):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.use_conv_transpose, self.name = use_conv, use_conv_transpose, name
        if use_conv_transpose:
            raise NotImplementedError
        if use_conv:
            self.conv = InflatedConv3d(self.channels, self.out_channels, 3, padding=1)

    def forward(self, hidden_states, output_size=None):
        assert hidden_states.shape[1] == self.channels
        tok, (b, c, f, h, w) = _tokens(hidden_states)
        if output_size is None:
            ho, wo = 2 * h, 2 * w
            up = ops.upsample2x(tok, b * f, h, w, c)
        else:   # forced size (resnet.py:75-76): [f, ho, wo] like F.interpolate(size=...) on the 5-D tensor, or (ho, wo)
            size = tuple(int(s) for s in output_size)
            if len(size) == 3:
                if size[0] != f:
                    raise NotImplementedError("Upsample3D: output_size must keep the number of frames")
                size = size[1:]
            ho, wo = size
            up = ops.upsample_nearest(tok, b * f, h, w, c, ho, wo)
        out = self.conv.forward_tokens(up, b * f, ho, wo, stats_rows=ops.stats_rows_for(ho * wo, f * ho * wo))
        return _untokens(out, b, self.out_channels, f, ho, wo)


class Downsample3D(nn.Module):
    """resnet.py:87-110 — stride-2 3x3 conv (gather + GEMM)."""

    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv"):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.padding, self.name = use_conv, padding, name
        if not use_conv:
            raise NotImplementedError
        self.conv = InflatedConv3d(self.channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, hidden_states):
        assert hidden_states.shape[1] == self.channels
        if self.padding == 0:
            raise NotImplementedError
        return self.conv(hidden_states)


class ResnetBlock3D(nn.Module):
    """resnet.py:113-207 — GN(5-D) -> SiLU -> conv3x3 -> +temb -> GN -> SiLU -> conv3x3 -> (+1x1 shortcut) -> /scale.

    Kernel plan: gn_stats + gn_apply(SiLU) -> bf16 | conv1 implicit GEMM, epilogue +bias +temb[b] -> fp32 |
    gn_stats + gn_apply(SiLU) | [shortcut GEMM] | conv2 implicit GEMM, epilogue +bias +residual, *1/scale.
    `input_tensor` may be a tuple (hidden, skip): the channel concatenation of unet_3d_blocks.py:629,731 is then
    never materialised in fp32 (both norm kernels and the shortcut operand read the two sources directly).
    """

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512,
                 groups=32, groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish",
                 time_embedding_norm="default", output_scale_factor=1.0, use_in_shortcut=None):
        super().__init__()
        self.pre_norm = True
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.time_embedding_norm = time_embedding_norm
        self.output_scale_factor = output_scale_factor
        if time_embedding_norm != "default":
            raise NotImplementedError("ResnetBlock3D: only time_embedding_norm='default' is on the hot path")
        if non_linearity not in ("swish", "silu"):
            raise NotImplementedError("ResnetBlock3D: only SiLU/swish")
        groups_out = groups if groups_out is None else groups_out
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = InflatedConv3d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups_out, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = InflatedConv3d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.use_in_shortcut = self.in_channels != self.out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = InflatedConv3d(in_channels, out_channels, 1) if self.use_in_shortcut else None
        self._temb_pk = None

    def _temb_packed(self):
        p = self.time_emb_proj
        sig = _sig(p.weight, p.bias)
        if self._temb_pk is None or self._temb_pk[2] != sig:
            self._temb_pk = (ops.pack_linear(p.weight), _f32c(p.bias), sig)
        return self._temb_pk[0], self._temb_pk[1]

    def forward(self, input_tensor, temb):
        srcs = input_tensor if isinstance(input_tensor, (tuple, list)) else (input_tensor,)
        toks, dims = [], None
        for s in srcs:
            t, d = _tokens(s)
            toks.append(t)
            dims = d if dims is None else dims
        b, _, f, h, w = dims
        n_img, rows_pb = b * f, f * h * w
        cin = sum(t.shape[1] for t in toks)
        assert cin == self.in_channels, f"ResnetBlock3D: got {cin} input channels, expected {self.in_channels}"
        g1, g2 = self.norm1, self.norm2
        need_raw = self.conv_shortcut is not None
        a1, raw = ops.group_norm(toks, g1.num_groups, rows_pb, b, g1.weight, g1.bias, g1.eps, True, want_raw=need_raw)
        row_bias = None
        if temb is not None and self.time_emb_proj is not None:
            act = getattr(temb, "_emote_silu_bf16", None)
            if act is None:
                act = ops.silu_bf16(temb.float().contiguous())
            wt, bt = self._temb_packed()
            row_bias = ops.gemm(act, wt, bias=bt)  # [b, cout] fp32
        h1 = self.conv1.forward_tokens(a1, n_img, h, w, row_bias=row_bias, rows_per_group=rows_pb,
                                       stats_rows=rows_pb if rows_pb % 128 == 0 else 0)  # feeds norm2 (5-D)
        a2, _ = ops.group_norm([h1], g2.num_groups, rows_pb, b, g2.weight, g2.bias, g2.eps, True)
        if need_raw:
            res = self.conv_shortcut.forward_tokens(raw, n_img, h, w)
        else:
            res = toks[0]
        out = self.conv2.forward_tokens(a2, n_img, h, w, residual=res, out_scale=1.0 / self.output_scale_factor,
                                        out=res if need_raw else None, stats_rows=ops.stats_rows_for(h * w, rows_pb))
        return _untokens(out, b, self.out_channels, f, h, w)


# =============================================================================================== orig_attention.py
class GEGLU(nn.Module):
    """orig_attention.py:806-827 — parameter container; the GEMM epilogue computes value * gelu_erf(gate)."""

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_PackedModule):
    """orig_attention.py:739-781 (diffusers FeedForward, activation_fn='geglu'): net = [GEGLU, Dropout, Linear]."""

    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4, dropout: float = 0.0,
                 activation_fn: str = "geglu"):
        super().__init__()
        if activation_fn != "geglu":
            raise NotImplementedError("FeedForward: only activation_fn='geglu' is on the hot path")
        inner = int(dim * mult)
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim)])

    def _pack(self):
        w1, b1 = ops.pack_geglu(self.net[0].proj.weight, self.net[0].proj.bias)
        return {"w1": w1, "b1": b1, "w2": ops.pack_linear(self.net[2].weight), "b2": _f32c(self.net[2].bias)}

    def run(self, a_bf16: torch.Tensor, residual: Optional[torch.Tensor], out=None, out_dtype=F32) -> torch.Tensor:
        p = self.pk
        mid = ops.gemm(a_bf16, p["w1"], bias=p["b1"], geglu=True, out_dtype=OP16)
        return ops.gemm(mid, p["w2"], bias=p["b2"], residual=residual, out=out, out_dtype=out_dtype)

    def forward(self, hidden_states):
        shp = hidden_states.shape
        a = ops.cast_bf16(hidden_states.float().contiguous().view(-1, shp[-1]))
        return self.run(a, None).view(*shp[:-1], -1)


class CrossAttention(_PackedModule):
    """orig_attention.py:516-736 (== diffusers `Attention` as used by attention.py:192-227): to_q/to_k/to_v without
    bias, to_out = [Linear(+bias), Dropout]; softmax(q k^T d^-1/2) v with `heads` heads."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 dropout: float = 0.0, bias=False, upcast_attention: bool = False, upcast_softmax: bool = False,
                 added_kv_proj_dim: Optional[int] = None, norm_num_groups: Optional[int] = None):
        super().__init__()
        if added_kv_proj_dim is not None or norm_num_groups is not None:
            raise NotImplementedError("CrossAttention: added_kv_proj_dim / group_norm variants are not on the hot path")
        inner = dim_head * heads
        self.is_self = cross_attention_dim is None
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention, self.upcast_softmax = upcast_attention, upcast_softmax
        self.scale = dim_head ** -0.5
        self.heads, self.dim_head, self.inner_dim = heads, dim_head, inner
        self.sliceable_head_dim = heads
        self._slice_size = None
        self._use_memory_efficient_attention_xformers = False
        self.added_kv_proj_dim = None
        self.group_norm = None
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])
        self._ctx_cache = None

    def set_attention_slice(self, slice_size):  # flash kernel never materialises scores; kept for API parity
        self._slice_size = slice_size

    def _pack(self):
        wq, wk, wv = self.to_q.weight, self.to_k.weight, self.to_v.weight
        p = {"wq": ops.pack_linear(wq), "wkv": ops.pack_linear(torch.cat([wk, wv], 0)),
             "wo": ops.pack_linear(self.to_out[0].weight), "bo": _f32c(self.to_out[0].bias)}
        if wq.shape[1] == wk.shape[1]:
            p["wqkv"] = ops.pack_linear(torch.cat([wq, wk, wv], 0))
        bias = [m.bias for m in (self.to_q, self.to_k, self.to_v)]
        if bias[0] is not None:
            p["bq"], p["bkv"] = _f32c(bias[0]), _f32c(torch.cat(bias[1:], 0))
            p["bqkv"] = _f32c(torch.cat(bias, 0))
        return p

    # -- kernel-level entry points used by the transformer blocks -------------------------------------------------
    def self_attention(self, a_bf16: torch.Tensor, batch: int, n: int, bank_kv: Optional[torch.Tensor] = None,
                       bank_n: int = 0, bank_div: int = 1, bank_first: int = 0) -> torch.Tensor:
        """a: LN output bf16 [batch*n, C] -> attention output bf16 [batch*n, inner] (before to_out)."""
        p, c = self.pk, self.inner_dim
        qkv = ops.gemm(a_bf16, p["wqkv"], bias=p.get("bqkv"), out_dtype=OP16)
        out = torch.empty((batch * n, c), dtype=OP16, device=a_bf16.device)
        kw = {}
        if bank_kv is not None:
            kw = dict(k1=bank_kv[:, :c], v1=bank_kv[:, c:], n1=bank_n, kv1_strides=(bank_n * 2 * c, 2 * c),
                      kv1_batch_div=bank_div, kv1_first_batch=bank_first)
        ops.attention(qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:], out, batch=batch, heads=self.heads,
                      head_dim=self.dim_head, nq=n, n0=n, q_strides=(n * 3 * c, 3 * c), kv0_strides=(n * 3 * c, 3 * c),
                      o_strides=(n * c, c), scale=self.scale, **kw)
        return out

    def project_kv(self, ctx: torch.Tensor) -> torch.Tensor:
        """context / bank [B, N, D] (any float dtype) -> bf16 [B*N, 2*inner] (k | v); cached per context tensor
        (the text / audio context is constant over the 50 denoising steps)."""
        # The cache holds a reference to the context tensor itself (identity + version): a data_ptr key would go
        # stale when the allocator hands the same address to a different tensor.
        cc = self._ctx_cache
        if CTX_KV_CACHE_ENABLED and cc is not None and cc[0] is ctx and cc[1] == ctx._version and self._pk is not None:
            return cc[2]
        p = self.pk
        flat = ctx.reshape(-1, ctx.shape[-1])
        a = flat if flat.dtype == OP16 and flat.is_contiguous() else ops.cast_bf16(flat.float().contiguous())
        kv = ops.gemm(a, p["wkv"], bias=p.get("bkv"), out_dtype=OP16)
        self._ctx_cache = (ctx, ctx._version, kv) if CTX_KV_CACHE_ENABLED else None
        return kv

    def cross_attention(self, a_bf16: torch.Tensor, batch: int, n: int, ctx: torch.Tensor) -> torch.Tensor:
        p, c = self.pk, self.inner_dim
        q = ops.gemm(a_bf16, p["wq"], bias=p.get("bq"), out_dtype=OP16)
        kv = self.project_kv(ctx)
        bc, nc = ctx.shape[0], ctx.shape[1]
        if batch % bc != 0:
            raise ValueError(f"context batch {bc} does not divide attention batch {batch}")
        out = torch.empty((batch * n, c), dtype=OP16, device=a_bf16.device)
        ops.attention(q, kv[:, :c], kv[:, c:], out, batch=batch, heads=self.heads, head_dim=self.dim_head, nq=n, n0=nc,
                      q_strides=(n * c, c), kv0_strides=(nc * 2 * c, 2 * c), o_strides=(n * c, c), scale=self.scale,
                      kv0_batch_div=batch // bc)
        return out

    def out_proj(self, attn_bf16: torch.Tensor, residual: Optional[torch.Tensor], out=None) -> torch.Tensor:
        p = self.pk
        return ops.gemm(attn_bf16, p["wo"], bias=p["bo"], residual=residual, out=out)

    # -- reference-compatible module call (orig_attention.py:598-653) ---------------------------------------------
    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        if attention_mask is not None:
            raise NotImplementedError("CrossAttention: attention_mask is not supported by the CUDA path")
        if not hidden_states.is_cuda:
            raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
        b, n, _ = hidden_states.shape
        a = ops.cast_bf16(hidden_states.float().contiguous().view(b * n, -1))
        if encoder_hidden_states is None:
            o = self.self_attention(a, b, n)
        else:
            o = self.cross_attention(a, b, n, encoder_hidden_states)
        return self.out_proj(o, None).view(b, n, -1)


# =============================================================================================== attention.py
@dataclass
class Transformer3DModelOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


class BasicTransformerBlock(nn.Module):
    """attention.py:164-320 with the ReferenceAttentionControl reader/writer semantics of
    mutual_self_attention.py:199-284 built in (no monkey patching): x += attn1(LN1 x [, bank]); x += attn2(LN2 x, ctx);
    x += FF(LN3 x).  Attributes touched by the reference hook (`norm1.normalized_shape`, `attn1`, `bank`, ...) exist."""

    def __init__(self, dim: int, num_attention_heads: int, attention_head_dim: int, dropout=0.0,
                 cross_attention_dim: Optional[int] = None, activation_fn: str = "geglu",
                 num_embeds_ada_norm: Optional[int] = None, attention_bias: bool = False,
                 only_cross_attention: bool = False, upcast_attention: bool = False,
                 unet_use_cross_frame_attention=None, unet_use_temporal_attention=None):
        super().__init__()
        assert unet_use_cross_frame_attention is not None and unet_use_temporal_attention is not None
        if num_embeds_ada_norm is not None or unet_use_cross_frame_attention or unet_use_temporal_attention:
            raise NotImplementedError("BasicTransformerBlock: AdaLayerNorm / cross-frame / attn_temp variants are cold paths")
        self.only_cross_attention = only_cross_attention
        self.use_ada_layer_norm = False
        self.use_ada_layer_norm_zero = False
        self.unet_use_cross_frame_attention = unet_use_cross_frame_attention
        self.unet_use_temporal_attention = unet_use_temporal_attention
        self.attn1 = CrossAttention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim,
                                    dropout=dropout, bias=attention_bias, upcast_attention=upcast_attention)
        self.norm1 = nn.LayerNorm(dim)
        if cross_attention_dim is not None:
            self.attn2 = CrossAttention(query_dim=dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
                                        dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                                        upcast_attention=upcast_attention)
            self.norm2 = nn.LayerNorm(dim)
        else:
            self.attn2, self.norm2 = None, None
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.norm3 = nn.LayerNorm(dim)
        # reference-attention state (set by ReferenceAttentionControl)
        self.bank: List[torch.Tensor] = []
        self._ref_mode: Optional[str] = None
        self._ref_cfg = False

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, attention_mask=None, video_length=None,
                _emit_bf16: bool = False):
        """`_emit_bf16` (internal, used by Transformer3DModel for its last block): the feed-forward epilogue writes the
        block output directly as the bf16 operand of proj_out instead of fp32 (no separate cast pass)."""
        if attention_mask is not None:
            raise NotImplementedError("BasicTransformerBlock: attention_mask is not supported by the CUDA path")
        if not hidden_states.is_cuda:
            raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
        bf, n, c = hidden_states.shape
        x = hidden_states.float().contiguous().view(bf * n, c)
        ln = self.norm1
        a = ops.layer_norm(x, ln.weight, ln.bias, ln.eps)
        bank_kv, bank_n, bank_first = None, 0, 0
        if self._ref_mode == "write":
            self.bank.append(a.view(bf, n, c).float())          # mutual_self_attention.py:230
        if self.attn1 is None:
            # AppearanceEncoderModel tail (appearance_encoder.py:613-621): the block was cut down to its norm1, whose
            # output feeds the bank above; the rest of the reference's block is identity stubs
            return hidden_states
        if self._ref_mode == "read" and len(self.bank) > 0:
            bank = self.bank[0] if len(self.bank) == 1 else torch.cat(list(self.bank), dim=1)
            bank_kv, bank_n = self.attn1.project_kv(bank), bank.shape[1]
            bank_first = bf // 2 if self._ref_cfg else 0
            self.bank = []  # mutual_self_attention.py:258
        o = self.attn1.self_attention(a, bf, n, bank_kv, bank_n, video_length or 1, bank_first)
        x = self.attn1.out_proj(o, x)  # new tensor: the caller's hidden_states is left untouched
        if self.attn2 is not None:
            ln = self.norm2
            a = ops.layer_norm(x, ln.weight, ln.bias, ln.eps)
            ctx = a.view(bf, n, c) if encoder_hidden_states is None else encoder_hidden_states
            o = self.attn2.cross_attention(a, bf, n, ctx)
            x = self.attn2.out_proj(o, x, out=x)
        ln = self.norm3
        a = ops.layer_norm(x, ln.weight, ln.bias, ln.eps)
        if _emit_bf16:
            return self.ff.run(a, x, out_dtype=OP16).view(bf, n, c)
        x = self.ff.run(a, x, out=x)
        return x.view(bf, n, c)


class Transformer3DModel(nn.Module):
    """attention.py:48-161 — per-frame GroupNorm -> proj_in -> transformer blocks -> proj_out -> + residual."""

    def __init__(self, num_attention_heads: int = 16, attention_head_dim: int = 88, in_channels: Optional[int] = None,
                 num_layers: int = 1, dropout: float = 0.0, norm_num_groups: int = 32,
                 cross_attention_dim: Optional[int] = None, attention_bias: bool = False, activation_fn: str = "geglu",
                 num_embeds_ada_norm: Optional[int] = None, use_linear_projection: bool = False,
                 only_cross_attention: bool = False, upcast_attention: bool = False,
                 unet_use_cross_frame_attention=None, unet_use_temporal_attention=None):
        super().__init__()
        self.config = AttrDict({k: v for k, v in locals().items() if k not in ("self", "__class__")})
        self.use_linear_projection = use_linear_projection
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        inner = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear_projection else nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                  num_embeds_ada_norm=num_embeds_ada_norm, attention_bias=attention_bias,
                                  only_cross_attention=only_cross_attention, upcast_attention=upcast_attention,
                                  unet_use_cross_frame_attention=unet_use_cross_frame_attention,
                                  unet_use_temporal_attention=unet_use_temporal_attention)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(in_channels, inner) if use_linear_projection else nn.Conv2d(inner, in_channels, 1)
        self._pk = None

    def _packed(self):
        tail = self.proj_out is None   # AppearanceEncoderModel's last transformer: GroupNorm -> proj_in -> norm1 only
        ver = _sig(self.proj_in.weight, self.proj_in.bias) + (() if tail else _sig(self.proj_out.weight, self.proj_out.bias))
        if self._pk is None or self._pk["ver"] != ver:
            self._pk = {"wi": ops.pack_linear(self.proj_in.weight), "bi": _f32c(self.proj_in.bias),
                        "wo": None if tail else ops.pack_linear(self.proj_out.weight),
                        "bo": None if tail else _f32c(self.proj_out.bias), "ver": ver}
        return self._pk

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, return_dict: bool = True):
        assert hidden_states.dim() == 5, f"Expected hidden_states to have ndim=5, but got ndim={hidden_states.dim()}."
        tok, (b, c, f, h, w) = _tokens(hidden_states)
        p, g = self._packed(), self.norm
        a, _ = ops.group_norm([tok], g.num_groups, h * w, b * f, g.weight, g.bias, g.eps, False)  # per frame
        x = ops.gemm(a, p["wi"], bias=p["bi"]).view(b * f, h * w, -1)
        if self.proj_out is None:
            self.transformer_blocks[0](x, encoder_hidden_states=encoder_hidden_states, timestep=timestep, video_length=f)
            return Transformer3DModelOutput(sample=hidden_states) if return_dict else (hidden_states,)
        nblk = len(self.transformer_blocks)
        for i, block in enumerate(self.transformer_blocks):
            # context is NOT repeated per frame (attention.py:118-119): the attention kernel indexes it by b = img // f
            x = block(x, encoder_hidden_states=encoder_hidden_states, timestep=timestep, video_length=f,
                      _emit_bf16=(i == nblk - 1))  # last FF epilogue emits the proj_out operand
        inner = x.shape[-1]
        xa = x.reshape(-1, inner)
        out = ops.gemm(xa if xa.dtype == OP16 else ops.cast_bf16(xa), p["wo"], bias=p["bo"], residual=tok,
                       stats_rows=ops.stats_rows_for(h * w, f * h * w))
        out = _untokens(out, b, c, f, h, w)
        return Transformer3DModelOutput(sample=out) if return_dict else (out,)


# =============================================================================================== motion_module.py
def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class PositionalEncoding(nn.Module):
    """motion_module.py:230-248 — sinusoidal table buffer `pe` [1, max_len, d_model]; added inside the LN kernel."""

    def __init__(self, d_model, dropout=0.0, max_len=24):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class VersatileAttention(CrossAttention):
    """motion_module.py:251-334 — temporal self-attention over the frame axis (attention_mode='Temporal')."""

    def __init__(self, attention_mode=None, cross_frame_attention_mode=None, temporal_position_encoding=False,
                 temporal_position_encoding_max_len=24, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert attention_mode == "Temporal"
        self.attention_mode = attention_mode
        self.is_cross_attention = kwargs["cross_attention_dim"] is not None
        if self.is_cross_attention:
            raise NotImplementedError("VersatileAttention: Temporal_Cross blocks are not on the hot path")
        self.pos_encoder = PositionalEncoding(kwargs["query_dim"], dropout=0.0, max_len=temporal_position_encoding_max_len) \
            if temporal_position_encoding else None

    def extra_repr(self):
        return f"(Module Info) Attention_Mode: {self.attention_mode}, Is_Cross_Attention: {self.is_cross_attention}"

    def run(self, x: torch.Tensor, norm: nn.LayerNorm, b: int, f: int, hw: int) -> torch.Tensor:
        """x fp32 [(b f hw), c] -> x + to_out(attn_over_frames(LN(x) + pe))   (in place on x)."""
        pe = None
        if self.pos_encoder is not None:
            if f > self.pos_encoder.pe.shape[1]:
                raise ValueError(f"video_length {f} exceeds temporal_position_encoding_max_len {self.pos_encoder.pe.shape[1]}")
            pe = self.pos_encoder.pe[0]
            if pe.dtype != F32 or not pe.is_contiguous():
                pe = pe.float().contiguous()
        a = ops.layer_norm(x, norm.weight, norm.bias, norm.eps, pe=pe, rows_per_frame=hw, frames=f)
        p = self.pk
        qkv = ops.gemm(a, p["wqkv"], bias=p.get("bqkv"), out_dtype=OP16)
        o = ops.temporal_attention(qkv, b, f, hw, self.heads, self.dim_head)
        return self.out_proj(o, x, out=x)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, video_length=None):
        # module-level call with the reference layout [(b f), d, c]; returns the attention branch only
        bf, d, c = hidden_states.shape
        x = hidden_states.float().contiguous().view(bf * d, c)
        pe = self.pos_encoder.pe[0].float().contiguous() if self.pos_encoder is not None else None
        a = ops.cast_bf16(x) if pe is None else None
        if a is None:  # add the table without a norm: identity LayerNorm is not available, so use torch for this cold path
            frame = (torch.arange(bf, device=x.device) % video_length).repeat_interleave(d)
            a = ops.cast_bf16((x + pe[frame]).contiguous())
        p = self.pk
        qkv = ops.gemm(a, p["wqkv"], bias=p.get("bqkv"), out_dtype=OP16)
        o = ops.temporal_attention(qkv, bf // video_length, video_length, d, self.heads, self.dim_head)
        return self.out_proj(o, None).view(bf, d, c)


class TemporalTransformerBlock(nn.Module):
    """motion_module.py:166-227 — [LN -> temporal attn -> +res] x len(attention_block_types), LN -> GEGLU FF -> +res."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, attention_block_types=("Temporal_Self", "Temporal_Self"),
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=768, activation_fn="geglu", attention_bias=False,
                 upcast_attention=False, cross_frame_attention_mode=None, temporal_position_encoding=False,
                 temporal_position_encoding_max_len=24):
        super().__init__()
        blocks, norms = [], []
        for name in attention_block_types:
            blocks.append(VersatileAttention(
                attention_mode=name.split("_")[0], cross_attention_dim=cross_attention_dim if name.endswith("_Cross") else None,
                query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                upcast_attention=upcast_attention, cross_frame_attention_mode=cross_frame_attention_mode,
                temporal_position_encoding=temporal_position_encoding,
                temporal_position_encoding_max_len=temporal_position_encoding_max_len))
            norms.append(nn.LayerNorm(dim))
        self.attention_blocks = nn.ModuleList(blocks)
        self.norms = nn.ModuleList(norms)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.ff_norm = nn.LayerNorm(dim)

    def run(self, x: torch.Tensor, b: int, f: int, hw: int, last_bf16: bool = False) -> torch.Tensor:
        for attn, norm in zip(self.attention_blocks, self.norms):
            x = attn.run(x, norm, b, f, hw)
        n = self.ff_norm
        a = ops.layer_norm(x, n.weight, n.bias, n.eps)
        if last_bf16:
            return self.ff.run(a, x, out_dtype=OP16)
        return self.ff.run(a, x, out=x)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, video_length=None):
        bf, d, c = hidden_states.shape
        x = hidden_states.float().contiguous().clone().view(bf * d, c)
        return self.run(x, bf // video_length, video_length, d).view(bf, d, c)


class TemporalTransformer3DModel(nn.Module):
    """motion_module.py:90-163 — per-frame GroupNorm -> Linear in -> blocks -> Linear out (zero-init) -> + residual."""

    def __init__(self, in_channels, num_attention_heads, attention_head_dim, num_layers,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), dropout=0.0, norm_num_groups=32,
                 cross_attention_dim=768, activation_fn="geglu", attention_bias=False, upcast_attention=False,
                 cross_frame_attention_mode=None, temporal_position_encoding=False, temporal_position_encoding_max_len=24):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            TemporalTransformerBlock(dim=inner, num_attention_heads=num_attention_heads, attention_head_dim=attention_head_dim,
                                     attention_block_types=attention_block_types, dropout=dropout,
                                     norm_num_groups=norm_num_groups, cross_attention_dim=cross_attention_dim,
                                     activation_fn=activation_fn, attention_bias=attention_bias,
                                     upcast_attention=upcast_attention, cross_frame_attention_mode=cross_frame_attention_mode,
                                     temporal_position_encoding=temporal_position_encoding,
                                     temporal_position_encoding_max_len=temporal_position_encoding_max_len)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels)
        self._pk = None

    def _packed(self):
        ver = _sig(self.proj_in.weight, self.proj_in.bias, self.proj_out.weight, self.proj_out.bias)
        if self._pk is None or self._pk["ver"] != ver:
            self._pk = {"wi": ops.pack_linear(self.proj_in.weight), "bi": _f32c(self.proj_in.bias),
                        "wo": ops.pack_linear(self.proj_out.weight), "bo": _f32c(self.proj_out.bias), "ver": ver}
        return self._pk

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        assert hidden_states.dim() == 5, f"Expected hidden_states to have ndim=5, but got ndim={hidden_states.dim()}."
        tok, (b, c, f, h, w) = _tokens(hidden_states)
        p, g = self._packed(), self.norm
        a, _ = ops.group_norm([tok], g.num_groups, h * w, b * f, g.weight, g.bias, g.eps, False)
        x = ops.gemm(a, p["wi"], bias=p["bi"])
        nblk = len(self.transformer_blocks)
        for i, block in enumerate(self.transformer_blocks):
            x = block.run(x, b, f, h * w, last_bf16=(i == nblk - 1))  # last FF epilogue emits the proj_out operand
        out = ops.gemm(x, p["wo"], bias=p["bo"], residual=tok, stats_rows=ops.stats_rows_for(h * w, f * h * w))
        return _untokens(out, b, c, f, h, w)


class VanillaTemporalModule(nn.Module):
    """motion_module.py:53-87."""

    def __init__(self, in_channels, num_attention_heads=8, num_transformer_block=2,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), cross_frame_attention_mode=None,
                 temporal_position_encoding=False, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1,
                 zero_initialize=True):
        super().__init__()
        self.temporal_transformer = TemporalTransformer3DModel(
            in_channels=in_channels, num_attention_heads=num_attention_heads,
            attention_head_dim=in_channels // num_attention_heads // temporal_attention_dim_div,
            num_layers=num_transformer_block, attention_block_types=attention_block_types,
            cross_frame_attention_mode=cross_frame_attention_mode, temporal_position_encoding=temporal_position_encoding,
            temporal_position_encoding_max_len=temporal_position_encoding_max_len)
        if zero_initialize:
            self.temporal_transformer.proj_out = zero_module(self.temporal_transformer.proj_out)

    def forward(self, input_tensor, temb, encoder_hidden_states, attention_mask=None, anchor_frame_idx=None):
        return self.temporal_transformer(input_tensor, encoder_hidden_states, attention_mask)


def get_motion_module(in_channels, motion_module_type: str, motion_module_kwargs: dict):
    if motion_module_type == "Vanilla":
        return VanillaTemporalModule(in_channels=in_channels, **motion_module_kwargs)
    raise ValueError


# =============================================================================================== unet_3d_blocks.py
class _Block3D(nn.Module):
    """Shared body of the five block containers (unet_3d_blocks.py:181-751): resnet -> [transformer] -> [motion]."""

    has_cross_attention = False

    def _make_resnet(self, cin, cout, kw):
        return ResnetBlock3D(in_channels=cin, out_channels=cout, temb_channels=kw["temb_channels"], eps=kw["resnet_eps"],
                             groups=kw["resnet_groups"], dropout=kw["dropout"], time_embedding_norm=kw["resnet_time_scale_shift"],
                             non_linearity=kw["resnet_act_fn"], output_scale_factor=kw["output_scale_factor"],
                             pre_norm=kw["resnet_pre_norm"])

    def _make_attn(self, ch, kw):
        if kw.get("dual_cross_attention"):
            raise NotImplementedError
        heads = kw["attn_num_head_channels"]
        return Transformer3DModel(heads, ch // heads, in_channels=ch, num_layers=1, cross_attention_dim=kw["cross_attention_dim"],
                                  norm_num_groups=kw["resnet_groups"], use_linear_projection=kw["use_linear_projection"],
                                  only_cross_attention=kw.get("only_cross_attention", False),
                                  upcast_attention=kw["upcast_attention"],
                                  unet_use_cross_frame_attention=kw["unet_use_cross_frame_attention"],
                                  unet_use_temporal_attention=kw["unet_use_temporal_attention"])

    def _make_motion(self, ch, kw):
        return get_motion_module(in_channels=ch, motion_module_type=kw["motion_module_type"],
                                 motion_module_kwargs=kw["motion_module_kwargs"]) if kw["use_motion_module"] else None

    def _layer(self, i, hidden_states, temb, encoder_hidden_states):
        hidden_states = self.resnets[i](hidden_states, temb)
        if self.has_cross_attention:
            hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states).sample
        mm = self.motion_modules[i]
        if mm is not None:
            hidden_states = mm(hidden_states, temb, encoder_hidden_states=encoder_hidden_states)
        return hidden_states


_COMMON = dict(dropout=0.0, num_layers=1, resnet_eps=1e-6, resnet_time_scale_shift="default", resnet_act_fn="swish",
               resnet_groups=32, resnet_pre_norm=True, output_scale_factor=1.0, use_motion_module=None,
               motion_module_type=None, motion_module_kwargs=None)
_ATTN = dict(attn_num_head_channels=1, cross_attention_dim=1280, dual_cross_attention=False, use_linear_projection=False,
             only_cross_attention=False, upcast_attention=False, unet_use_cross_frame_attention=None,
             unet_use_temporal_attention=None)


def _kw(defaults, given, required):
    unknown = set(given) - set(defaults) - set(required)
    if unknown:
        raise TypeError(f"unexpected keyword arguments: {sorted(unknown)}")
    missing = [r for r in required if r not in given]
    if missing:
        raise TypeError(f"missing required arguments: {missing}")
    out = dict(defaults)
    out.update(given)
    return out


class UNetMidBlock3DCrossAttn(_Block3D):
    """unet_3d_blocks.py:181-283."""
    has_cross_attention = True

    def __init__(self, **kwargs):
        super().__init__()
        kw = _kw({**_COMMON, **_ATTN}, kwargs, ("in_channels", "temb_channels"))
        if kw["resnet_groups"] is None:
            kw["resnet_groups"] = min(kw["in_channels"] // 4, 32)
        ch = kw["in_channels"]
        self.attn_num_head_channels = kw["attn_num_head_channels"]
        resnets, attns, motions = [self._make_resnet(ch, ch, kw)], [], []
        for _ in range(kw["num_layers"]):
            attns.append(self._make_attn(ch, kw))
            motions.append(self._make_motion(ch, kw))
            resnets.append(self._make_resnet(ch, ch, kw))
        self.attentions, self.resnets, self.motion_modules = nn.ModuleList(attns), nn.ModuleList(resnets), nn.ModuleList(motions)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet, mm in zip(self.attentions, self.resnets[1:], self.motion_modules):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states).sample
            if mm is not None:
                hidden_states = mm(hidden_states, temb, encoder_hidden_states=encoder_hidden_states)
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class _DownBase(_Block3D):
    def _build(self, kw, with_attn):
        cin, cout = kw["in_channels"], kw["out_channels"]
        resnets, attns, motions = [], [], []
        for i in range(kw["num_layers"]):
            resnets.append(self._make_resnet(cin if i == 0 else cout, cout, kw))
            if with_attn:
                attns.append(self._make_attn(cout, kw))
            motions.append(self._make_motion(cout, kw))
        if with_attn:
            self.attentions = nn.ModuleList(attns)
        self.resnets, self.motion_modules = nn.ModuleList(resnets), nn.ModuleList(motions)
        self.downsamplers = nn.ModuleList([Downsample3D(cout, use_conv=True, out_channels=cout,
                                                        padding=kw["downsample_padding"], name="op")]) \
            if kw["add_downsample"] else None
        self.gradient_checkpointing = False

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None):
        output_states = ()
        for i in range(len(self.resnets)):
            hidden_states = self._layer(i, hidden_states, temb, encoder_hidden_states)
            output_states += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            output_states += (hidden_states,)
        return hidden_states, output_states


class CrossAttnDownBlock3D(_DownBase):
    """unet_3d_blocks.py:286-423."""
    has_cross_attention = True

    def __init__(self, **kwargs):
        super().__init__()
        kw = _kw({**_COMMON, **_ATTN, "downsample_padding": 1, "add_downsample": True}, kwargs,
                 ("in_channels", "out_channels", "temb_channels"))
        self.attn_num_head_channels = kw["attn_num_head_channels"]
        self._build(kw, True)


class DownBlock3D(_DownBase):
    """unet_3d_blocks.py:426-519."""

    def __init__(self, **kwargs):
        super().__init__()
        kw = _kw({**_COMMON, "downsample_padding": 1, "add_downsample": True}, kwargs,
                 ("in_channels", "out_channels", "temb_channels"))
        self._build(kw, False)


class _UpBase(_Block3D):
    def _build(self, kw, with_attn):
        cin, cout, prev = kw["in_channels"], kw["out_channels"], kw["prev_output_channel"]
        n = kw["num_layers"]
        resnets, attns, motions = [], [], []
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = prev if i == 0 else cout
            resnets.append(self._make_resnet(rin + skip, cout, kw))
            if with_attn:
                attns.append(self._make_attn(cout, kw))
            motions.append(self._make_motion(cout, kw))
        if with_attn:
            self.attentions = nn.ModuleList(attns)
        self.resnets, self.motion_modules = nn.ModuleList(resnets), nn.ModuleList(motions)
        self.upsamplers = nn.ModuleList([Upsample3D(cout, use_conv=True, out_channels=cout)]) if kw["add_upsample"] else None
        self.gradient_checkpointing = False

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None, upsample_size=None,
                attention_mask=None):
        for i in range(len(self.resnets)):
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            # torch.cat([hidden, skip], dim=1) of the reference is folded into the resnet's norm / shortcut kernels
            hidden_states = self._layer(i, (hidden_states, skip), temb, encoder_hidden_states)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class CrossAttnUpBlock3D(_UpBase):
    """unet_3d_blocks.py:522-662."""
    has_cross_attention = True

    def __init__(self, **kwargs):
        super().__init__()
        kw = _kw({**_COMMON, **_ATTN, "add_upsample": True}, kwargs,
                 ("in_channels", "out_channels", "prev_output_channel", "temb_channels"))
        self.attn_num_head_channels = kw["attn_num_head_channels"]
        self._build(kw, True)


class UpBlock3D(_UpBase):
    """unet_3d_blocks.py:665-751."""

    def __init__(self, **kwargs):
        super().__init__()
        kw = _kw({**_COMMON, "add_upsample": True}, kwargs,
                 ("in_channels", "out_channels", "prev_output_channel", "temb_channels"))
        self._build(kw, False)


_DOWN_TYPES = {"DownBlock3D": DownBlock3D, "CrossAttnDownBlock3D": CrossAttnDownBlock3D}
_UP_TYPES = {"UpBlock3D": UpBlock3D, "CrossAttnUpBlock3D": CrossAttnUpBlock3D}


def _filter_kwargs(cls, kw):
    allowed = {**_COMMON, **(_ATTN if cls.has_cross_attention else {})}
    extra = {"in_channels", "out_channels", "temb_channels", "prev_output_channel", "add_downsample", "add_upsample",
             "downsample_padding"}
    return {k: v for k, v in kw.items() if k in allowed or k in extra}


def get_down_block(down_block_type, **kw):
    """unet_3d_blocks.py:30-103."""
    name = down_block_type[7:] if down_block_type.startswith("UNetRes") else down_block_type
    if name not in _DOWN_TYPES:
        raise ValueError(f"{name} does not exist.")
    cls = _DOWN_TYPES[name]
    if cls.has_cross_attention and kw.get("cross_attention_dim") is None:
        raise ValueError("cross_attention_dim must be specified for CrossAttnDownBlock3D")
    kw = _filter_kwargs(cls, kw)
    kw.pop("add_upsample", None), kw.pop("prev_output_channel", None)
    return cls(**kw)


def get_up_block(up_block_type, **kw):
    """unet_3d_blocks.py:106-178."""
    name = up_block_type[7:] if up_block_type.startswith("UNetRes") else up_block_type
    if name not in _UP_TYPES:
        raise ValueError(f"{name} does not exist.")
    cls = _UP_TYPES[name]
    if cls.has_cross_attention and kw.get("cross_attention_dim") is None:
        raise ValueError("cross_attention_dim must be specified for CrossAttnUpBlock3D")
    kw = _filter_kwargs(cls, kw)
    kw.pop("add_downsample", None), kw.pop("downsample_padding", None)
    return cls(**kw)


# =============================================================================================== embeddings.py
class Timesteps(nn.Module):
    """embeddings.py:221-235 — sinusoidal projection (kernel: emote_timestep_embedding, bf16 GEMM operand)."""

    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        return ops.timestep_embedding(timesteps.float().contiguous(), self.num_channels, self.flip_sin_to_cos,
                                      self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    """embeddings.py:161-218 — Linear -> SiLU -> Linear."""

    def __init__(self, in_channels: int, time_embed_dim: int, act_fn: str = "silu", out_dim: int = None):
        super().__init__()
        if act_fn != "silu":
            raise NotImplementedError
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)
        self._pk = None

    def forward(self, sample, condition=None):
        if condition is not None:
            raise NotImplementedError
        ver = _sig(self.linear_1.weight, self.linear_1.bias, self.linear_2.weight, self.linear_2.bias)
        if self._pk is None or self._pk["ver"] != ver:
            self._pk = {"w1": ops.pack_linear(self.linear_1.weight), "b1": _f32c(self.linear_1.bias),
                        "w2": ops.pack_linear(self.linear_2.weight), "b2": _f32c(self.linear_2.bias), "ver": ver}
        p = self._pk
        a = sample if sample.dtype == OP16 else ops.cast_bf16(sample.float().contiguous())
        h = ops.gemm(a, p["w1"], bias=p["b1"])
        return ops.gemm(ops.silu_bf16(h), p["w2"], bias=p["b2"])


# =============================================================================================== unet_controlnet.py
@dataclass
class UNet3DConditionOutput:
    sample: torch.Tensor

    def __getitem__(self, i):
        return (self.sample,)[i]


class UNet3DConditionModel(nn.Module):
    """unet_controlnet.py:54-525 — SD-1.5 UNet inflated to video with AnimateDiff motion modules."""

    _supports_gradient_checkpointing = True
    config_name = "config.json"

    def __init__(self, sample_size: Optional[int] = None, in_channels: int = 4, out_channels: int = 4,
                 center_input_sample: bool = False, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 down_block_types: Tuple[str] = ("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
                 mid_block_type: str = "UNetMidBlock3DCrossAttn",
                 up_block_types: Tuple[str] = ("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
                 only_cross_attention: Union[bool, Tuple[bool]] = False,
                 block_out_channels: Tuple[int] = (320, 640, 1280, 1280), layers_per_block: int = 2,
                 downsample_padding: int = 1, mid_block_scale_factor: float = 1, act_fn: str = "silu",
                 norm_num_groups: int = 32, norm_eps: float = 1e-5, cross_attention_dim: int = 1280,
                 attention_head_dim: Union[int, Tuple[int]] = 8, dual_cross_attention: bool = False,
                 use_linear_projection: bool = False, class_embed_type: Optional[str] = None,
                 num_class_embeds: Optional[int] = None, upcast_attention: bool = False,
                 resnet_time_scale_shift: str = "default",
                 use_motion_module=False, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False,
                 motion_module_decoder_only=False, motion_module_type=None, motion_module_kwargs={},
                 unet_use_cross_frame_attention=None, unet_use_temporal_attention=None):
        super().__init__()
        self.config = AttrDict({k: v for k, v in locals().items() if k not in ("self", "__class__")})
        self.sample_size = sample_size
        self.in_channels = in_channels
        time_embed_dim = block_out_channels[0] * 4
        self.conv_in = InflatedConv3d(in_channels, block_out_channels[0], kernel_size=3, padding=(1, 1))
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim)
        # class embedding (unet_controlnet.py:121-128)
        if class_embed_type is None and num_class_embeds is not None:
            self.class_embedding = nn.Embedding(num_class_embeds, time_embed_dim)
        elif class_embed_type == "timestep":
            self.class_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim)
        elif class_embed_type == "identity":
            self.class_embedding = nn.Identity(time_embed_dim, time_embed_dim)
        else:
            self.class_embedding = None
        n = len(down_block_types)
        if isinstance(only_cross_attention, bool):
            only_cross_attention = [only_cross_attention] * n
        if isinstance(attention_head_dim, int):
            attention_head_dim = (attention_head_dim,) * n
        shared = dict(temb_channels=time_embed_dim, resnet_eps=norm_eps, resnet_act_fn=act_fn, resnet_groups=norm_num_groups,
                      cross_attention_dim=cross_attention_dim, dual_cross_attention=dual_cross_attention,
                      use_linear_projection=use_linear_projection, upcast_attention=upcast_attention,
                      resnet_time_scale_shift=resnet_time_scale_shift,
                      unet_use_cross_frame_attention=unet_use_cross_frame_attention,
                      unet_use_temporal_attention=unet_use_temporal_attention,
                      motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs)
        self.down_blocks = nn.ModuleList()
        out_ch = block_out_channels[0]
        for i, t in enumerate(down_block_types):
            in_ch, out_ch = out_ch, block_out_channels[i]
            self.down_blocks.append(get_down_block(
                t, num_layers=layers_per_block, in_channels=in_ch, out_channels=out_ch, add_downsample=i != n - 1,
                attn_num_head_channels=attention_head_dim[i], downsample_padding=downsample_padding,
                only_cross_attention=only_cross_attention[i],
                use_motion_module=use_motion_module and (2 ** i in motion_module_resolutions) and not motion_module_decoder_only,
                **shared))
        if mid_block_type != "UNetMidBlock3DCrossAttn":
            raise ValueError(f"unknown mid_block_type : {mid_block_type}")
        mid_kw = {k: v for k, v in shared.items()}
        self.mid_block = UNetMidBlock3DCrossAttn(
            in_channels=block_out_channels[-1], output_scale_factor=mid_block_scale_factor,
            attn_num_head_channels=attention_head_dim[-1],
            use_motion_module=use_motion_module and motion_module_mid_block, **mid_kw)
        self.num_upsamplers = 0
        self.up_blocks = nn.ModuleList()
        rev_ch = list(reversed(block_out_channels))
        rev_heads = list(reversed(attention_head_dim))
        rev_oca = list(reversed(only_cross_attention))
        out_ch = rev_ch[0]
        for i, t in enumerate(up_block_types):
            prev, out_ch = out_ch, rev_ch[i]
            in_ch = rev_ch[min(i + 1, n - 1)]
            final = i == n - 1
            self.num_upsamplers += 0 if final else 1
            self.up_blocks.append(get_up_block(
                t, num_layers=layers_per_block + 1, in_channels=in_ch, out_channels=out_ch, prev_output_channel=prev,
                add_upsample=not final, attn_num_head_channels=rev_heads[i], only_cross_attention=rev_oca[i],
                use_motion_module=use_motion_module and (2 ** (3 - i) in motion_module_resolutions), **shared))
        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=norm_num_groups, eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = InflatedConv3d(block_out_channels[0], out_channels, kernel_size=3, padding=1)

    # -- diffusers-style conveniences the callers rely on ---------------------------------------------------------
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def from_config(cls, config, **kwargs):
        import inspect
        params = inspect.signature(cls.__init__).parameters
        merged = {k: v for k, v in dict(config).items() if k in params}
        merged.update({k: v for k, v in kwargs.items() if k in params})
        return cls(**merged)

    def set_attention_slice(self, slice_size):
        """unet_controlnet.py:259-322 — accepted for API parity; the flash kernel never materialises the scores."""
        for m in self.modules():
            if isinstance(m, CrossAttention):
                m.set_attention_slice(None if slice_size in ("auto", "max") or isinstance(slice_size, list) else slice_size)

    def _set_gradient_checkpointing(self, module, value=False):
        if isinstance(module, (_DownBase, _UpBase)):
            module.gradient_checkpointing = value

    @torch.no_grad()
    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int], encoder_hidden_states: torch.Tensor,
                class_labels: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None, return_dict: bool = True):
        """unet_controlnet.py:328-483."""
        if not sample.is_cuda:
            raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is not supported by the CUDA path")
        # the up blocks are told the size to produce when the latent is not a multiple of 2**num_upsamplers (:355-364)
        forward_upsample_size = any(s % (2 ** self.num_upsamplers) != 0 for s in sample.shape[-2:])
        in_dtype = sample.dtype
        sample = sample.float()
        if self.config.center_input_sample:
            sample = 2 * sample - 1.0
        # time embedding (unet_controlnet.py:376-398)
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.float32, device=sample.device)
        timesteps = timesteps.reshape(-1).to(device=sample.device, dtype=torch.float32).expand(sample.shape[0]).contiguous()
        emb = self.time_embedding(self.time_proj(timesteps))
        if self.class_embedding is not None:                     # unet_controlnet.py:400-408
            if class_labels is None:
                raise ValueError("class_labels should be provided when num_class_embeds > 0")
            if self.config.class_embed_type == "timestep":
                cl = class_labels.reshape(-1).to(device=sample.device, dtype=torch.float32).expand(sample.shape[0]).contiguous()
                class_emb = self.class_embedding(self.time_proj(cl))
            elif isinstance(self.class_embedding, nn.Embedding):   # a table row per sample: parameter indexing, no arithmetic
                class_emb = self.class_embedding.weight.detach()[class_labels.reshape(-1).long()].float()
                class_emb = class_emb.expand(sample.shape[0], -1).contiguous()
            else:
                class_emb = class_labels.to(device=sample.device, dtype=torch.float32).expand(sample.shape[0], -1).contiguous()
            emb = ops.add_f32(emb.contiguous(), class_emb)
        emb._emote_silu_bf16 = ops.silu_bf16(emb)  # shared by the 22 resnets' time_emb_proj

        sample = self.conv_in(sample.contiguous())
        is_controlnet = mid_block_additional_residual is not None and down_block_additional_residuals is not None
        down_res = (sample,)
        for blk in self.down_blocks:
            sample, res = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states)
            down_res += res
        if is_controlnet:
            down_res = tuple(self._add_residual(a, b) for a, b in zip(down_res, down_block_additional_residuals))
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        if is_controlnet:
            sample = self._add_residual(sample, mid_block_additional_residual)
        for i, blk in enumerate(self.up_blocks):
            k = len(blk.resnets)
            res, down_res = down_res[-k:], down_res[:-k]
            upsample_size = None
            if forward_upsample_size and i != len(self.up_blocks) - 1:
                upsample_size = down_res[-1].shape[2:]           # [f, h, w] of the skip the next block starts from (:458-460)
            sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                         encoder_hidden_states=encoder_hidden_states, upsample_size=upsample_size)
        tok, (b, c, f, h, w) = _tokens(sample)
        g = self.conv_norm_out
        a, _ = ops.group_norm([tok], g.num_groups, f * h * w, b, g.weight, g.bias, g.eps, True)
        co = self.conv_out.out_channels
        ld = (co + 3) // 4 * 4
        out_tok = torch.empty((b * f * h * w, ld), dtype=F32, device=tok.device)
        self.conv_out.forward_tokens(a, b * f, h, w, out=out_tok)
        if ld != co:
            out_tok = out_tok[:, :co].contiguous()
        out = ops.tokens_to_ncfhw(out_tok, b, co, f, h, w)
        if in_dtype != F32:
            out = out.to(in_dtype)
        return UNet3DConditionOutput(sample=out) if return_dict else (out,)

    @staticmethod
    def _add_residual(a: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
        ta, dims = _tokens(a)
        tr, _ = _tokens(r)
        return _untokens(ops.add_f32(ta.contiguous(), tr.contiguous()), *dims)

    @classmethod
    def from_pretrained_2d(cls, pretrained_model_path, subfolder=None, unet_additional_kwargs=None):
        """unet_controlnet.py:485-525 — build from a 2-D SD UNet folder (config.json + diffusion_pytorch_model.bin)."""
        if subfolder is not None:
            pretrained_model_path = os.path.join(pretrained_model_path, subfolder)
        config_file = os.path.join(pretrained_model_path, "config.json")
        if not os.path.isfile(config_file):
            raise RuntimeError(f"{config_file} does not exist")
        with open(config_file, "r") as fh:
            config = json.load(fh)
        config["_class_name"] = cls.__name__
        config["down_block_types"] = ["CrossAttnDownBlock3D"] * 3 + ["DownBlock3D"]
        config["up_block_types"] = ["UpBlock3D"] + ["CrossAttnUpBlock3D"] * 3
        model = cls.from_config(config, **(unet_additional_kwargs or {}))
        model_file = os.path.join(pretrained_model_path, "diffusion_pytorch_model.bin")
        if not os.path.isfile(model_file):
            raise RuntimeError(f"{model_file} does not exist")
        state_dict = torch.load(model_file, map_location="cpu")
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        print(f"### missing keys: {len(missing)}; \n### unexpected keys: {len(unexpected)};")
        n_temporal = sum(p.numel() for n, p in model.named_parameters() if "temporal" in n)
        print(f"### Temporal Module Parameters: {n_temporal / 1e6} M")
        return model


# =============================================================================================== mutual_self_attention.py
def torch_dfs(model: nn.Module):
    """stable_diffusion_controlnet_reference.py:65-69."""
    result = [model]
    for child in model.children():
        result += torch_dfs(child)
    return result


def reference_blocks(unet, fusion_blocks: str = "midup"):
    """The BasicTransformerBlocks a ReferenceAttentionControl hooks, in the reference's pairing order: depth-first over
    mid + up blocks (or the whole network), sorted by descending norm1 width (mutual_self_attention.py:534-543, 585-588;
    the sort is stable, so writer and reader lists pair up block by block)."""
    if fusion_blocks == "midup":
        mods = torch_dfs(unet.mid_block) + torch_dfs(unet.up_blocks)
    else:
        mods = torch_dfs(unet)
    mods = [m for m in mods if hasattr(m, "norm1") and hasattr(m, "attn1") and hasattr(m, "bank")]
    return sorted(mods, key=lambda x: -x.norm1.normalized_shape[0])


class ReferenceAttentionControl:
    """mutual_self_attention.py:128-641 — reference-attention reader/writer.  Instead of monkey-patching
    `BasicTransformerBlock.forward`, it flips a mode flag the blocks' own forward honours: writer blocks append
    LN1(x) to `.bank`; reader blocks attend to [self | bank] with the unconditional CFG half masked off the bank
    inside the attention kernel (replaces compute-twice-and-overwrite, :239-255)."""

    def __init__(self, unet, mode="write", do_classifier_free_guidance=False, attention_auto_machine_weight=float("inf"),
                 gn_auto_machine_weight=1.0, style_fidelity=1.0, reference_attn=True, reference_adain=False,
                 fusion_blocks="midup", batch_size=1):
        assert mode in ["read", "write"]
        assert fusion_blocks in ["midup", "full"]
        if reference_adain:
            raise NotImplementedError("reference_adain (GroupNorm statistics hacks, :319-530) is off by default and not implemented")
        self.unet, self.mode = unet, mode
        self.reference_attn, self.reference_adain, self.fusion_blocks = reference_attn, reference_adain, fusion_blocks
        if reference_attn:
            for i, m in enumerate(self._blocks(unet)):
                m._ref_mode, m._ref_cfg = mode, bool(do_classifier_free_guidance)
                m.bank = []
                m.attn_weight = float(i) / max(1, len(self._blocks(unet)))

    def _blocks(self, unet):
        return reference_blocks(unet, self.fusion_blocks)

    def update(self, writer, dtype=torch.float16):
        """:577-617 — copy the writer's banks into the reader blocks (paired by descending width)."""
        if not self.reference_attn:
            return
        src = writer.unet if hasattr(writer, "unet") else writer
        writer_blocks = reference_blocks(src, self.fusion_blocks)
        for r, w in zip(self._blocks(self.unet), writer_blocks):
            # the reference clones and casts to `dtype` (fp16, :587-588); the banks here stay fp32 until the attention
            # kernel's K/V projection rounds them to the operand type, so `dtype` only has to be a float type
            if dtype is not None and not torch.empty(0, dtype=dtype).is_floating_point():
                raise TypeError(f"ReferenceAttentionControl.update: dtype must be a floating type, got {dtype}")
            r.bank = [v.clone() for v in w.bank]

    def set_banks(self, banks: Dict[str, List[torch.Tensor]]):
        """Load banks by block name (e.g. 'mid_block.attentions.0.transformer_blocks.0') — for synthetic banks."""
        named = dict(self.unet.named_modules())
        for name, tensors in banks.items():
            named[name].bank = [t for t in tensors]

    def clear(self):
        """:619-641."""
        if self.reference_attn:
            for m in self._blocks(self.unet):
                m.bank = []

    def release(self):
        """clear() + switch the blocks back to plain self-attention (the reference's monkey patch stays installed for the
        life of the module; here the mode is a flag, so a control object can hand the network back)"""
        if self.reference_attn:
            for m in self._blocks(self.unet):
                m.bank = []
                m._ref_mode = None
