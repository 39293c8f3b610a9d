"""SD VAE on the sm_100a kernels — the `vae.decode(...)` the reference calls once per frame at
EMOAnimationPipeline.py:291-307 and the `vae.encode(...)` of `images2latents` (:402-414) that turns the reference image
into ReferenceNet latents (`diffusers.AutoencoderKL`, a third-party dependency absent from the reference tree; topology
restated in oracle/vae_decoder.py).

`AutoencoderKL` here mirrors the diffusers interface the pipeline touches (`decode(z).sample`, `config.scaling_factor`)
and diffusers' state_dict key names for `post_quant_conv.*` and `decoder.*` (both the legacy
`query/key/value/proj_attn` and the newer `to_q/to_k/to_v/to_out.0` attention names load).  All frames are decoded
in one batch instead of a Python loop of single-frame calls; `(x/2+0.5).clamp(0,1)` is fused with the final
layout change.  Kernel plan per stage: gn_stats/gn_apply(SiLU) -> implicit-GEMM conv (tcgen05) with fused bias /
residual; the 512-wide single-head mid attention is QK^T GEMM -> row softmax -> PV GEMM (V^T produced directly by
swapping GEMM operand roles), per frame.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from . import ops
from ._lib import EmoteKernelError
from .unet3d import AttrDict, _f32c, _sig

F32, OP16 = torch.float32, ops.OP16


class _Conv(nn.Conv2d):
    """parameter container + packed-weight cache for a 1x1 / 3x3 conv over NHWC tokens"""

    def __init__(self, cin, cout, k):
        super().__init__(cin, cout, k, padding=k // 2)
        self._pk = None

    def packed(self):
        sig = _sig(self.weight, self.bias)
        if self._pk is None or self._pk[2] != sig:
            with torch.no_grad():
                k = self.kernel_size[0]
                if k == 1:
                    w = ops.pack_linear(self.weight)
                elif self.in_channels <= 7:
                    w = ops.pack_conv3x3_small(self.weight)
                else:
                    w = ops.pack_conv3x3(self.weight)
            self._pk = (w, _f32c(self.bias), sig)
        return self._pk[0], self._pk[1]

    def run(self, a_bf16, n_img, h, w, **epi):
        wp, b = self.packed()
        if self.kernel_size[0] == 1:
            return ops.gemm(a_bf16, wp, bias=b, **epi)
        return ops.conv3x3(a_bf16, wp, n_img, h, w, self.in_channels, bias=b, **epi)


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D with temb=None (same arithmetic as reference resnet.py:177-207 with one frame):
    GN -> SiLU -> conv3x3 -> GN -> SiLU -> conv3x3 -> + (1x1 shortcut of) input."""

    def __init__(self, cin, cout, groups=32, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = _Conv(cin, cout, 3)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = _Conv(cout, cout, 3)
        self.conv_shortcut = _Conv(cin, cout, 1) if cin != cout else None

    def run(self, tok, n_img, h, w):
        g1, g2 = self.norm1, self.norm2
        raw_needed = self.conv_shortcut is not None
        a1, raw = ops.group_norm([tok], g1.num_groups, h * w, n_img, g1.weight, g1.bias, g1.eps, True, want_raw=raw_needed)
        sr = ops.stats_rows_for(h * w, h * w)   # per-image GroupNorm statistics ride on the conv epilogues
        h1 = self.conv1.run(a1, n_img, h, w, stats_rows=sr)
        a2, _ = ops.group_norm([h1], g2.num_groups, h * w, n_img, g2.weight, g2.bias, g2.eps, True)
        if raw_needed:
            res = self.conv_shortcut.run(raw, n_img, h, w)
            return self.conv2.run(a2, n_img, h, w, residual=res, out=res, stats_rows=sr)
        return self.conv2.run(a2, n_img, h, w, residual=tok, stats_rows=sr)


class AttentionBlock(nn.Module):
    """legacy diffusers AttentionBlock (vendored at reference orig_attention.py:253-385): GroupNorm -> q/k/v Linear
    (+bias) -> softmax(q k^T / sqrt(C)) v, one head -> proj_attn -> + residual."""

    def __init__(self, channels, groups=32, eps=1e-6):
        super().__init__()
        self.channels = channels
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.proj_attn = nn.Linear(channels, channels)
        self._pk = None
        self._register_load_state_dict_pre_hook(self._rename)

    @staticmethod
    def _rename(state_dict, prefix, *a):
        for new, old in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
            for suffix in ("weight", "bias"):
                k = f"{prefix}{new}.{suffix}"
                if k in state_dict:
                    state_dict[f"{prefix}{old}.{suffix}"] = state_dict.pop(k)

    def _packed(self):
        ver = _sig(*[t for m in (self.query, self.key, self.value, self.proj_attn) for t in (m.weight, m.bias)])
        if self._pk is None or self._pk["ver"] != ver:
            with torch.no_grad():
                self._pk = {
                    "wqkv": ops.pack_linear(torch.cat([self.query.weight, self.key.weight, self.value.weight], 0)),
                    "bqkv": _f32c(torch.cat([self.query.bias, self.key.bias, self.value.bias], 0)),
                    "wo": ops.pack_linear(self.proj_attn.weight), "bo": _f32c(self.proj_attn.bias), "ver": ver}
        return self._pk

    def run(self, tok, n_img, h, w):
        """GroupNorm -> fused q|k|v GEMM -> flash attention over the h*w positions of every image (one head of `channels`
        dims, all images in one launch) -> proj_attn + residual.  512 channels (the SD VAE): tcgen05 kernel with the
        accumulator in TMEM (`emote_attention_wide_bf16`); narrow test networks (<= 160 channels): the generic kernels."""
        c, n = self.channels, h * w
        g, p = self.group_norm, self._packed()
        a, _ = ops.group_norm([tok], g.num_groups, n, n_img, g.weight, g.bias, g.eps, False)
        qkv = ops.gemm(a, p["wqkv"], bias=p["bqkv"], out_dtype=OP16)      # [n_img*n, 3c]
        attn = torch.empty((n_img * n, c), dtype=OP16, device=tok.device)
        q, k, v = qkv[:, :c], qkv[:, c:2 * c], qkv[:, 2 * c:]
        if c == 512:
            ops.attention_wide(q, k, v, attn, batch=n_img, nq=n, nk=n, q_strides=(n * 3 * c, 3 * c),
                               kv_strides=(n * 3 * c, 3 * c), o_strides=(n * c, c), scale=c ** -0.5)
        elif c <= 160 and c % 8 == 0:
            ops.attention(q, k, v, attn, batch=n_img, heads=1, head_dim=c, nq=n, n0=n, q_strides=(n * 3 * c, 3 * c),
                          kv0_strides=(n * 3 * c, 3 * c), o_strides=(n * c, c), scale=c ** -0.5)
        else:
            raise NotImplementedError(f"AttentionBlock: {c} channels (supported: 512, or <= 160 in steps of 8)")
        return ops.gemm(attn, p["wo"], bias=p["bo"], residual=tok, stats_rows=ops.stats_rows_for(n, n))


class _Up(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = _Conv(ch, ch, 3)


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, n_layers, add_upsample, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups, eps) for i in range(n_layers)])
        self.upsamplers = nn.ModuleList([_Up(cout)]) if add_upsample else None


class _Mid(nn.Module):
    def __init__(self, ch, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, groups, eps), ResnetBlock2D(ch, ch, groups, eps)])
        self.attentions = nn.ModuleList([AttentionBlock(ch, groups, eps)])


class Decoder(nn.Module):
    def __init__(self, latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32, eps=1e-6):
        super().__init__()
        top = block_out_channels[-1]
        self.conv_in = _Conv(latent_channels, top, 3)
        self.mid_block = _Mid(top, norm_num_groups, eps)
        rev = list(reversed(block_out_channels))
        blocks, cin = [], top
        for i, cout in enumerate(rev):
            blocks.append(_UpBlock(cin, cout, layers_per_block + 1, i != len(rev) - 1, norm_num_groups, eps))
            cin = cout
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, rev[-1], eps=eps)
        self.conv_out = _Conv(rev[-1], out_channels, 3)


class _Down(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = _Conv(ch, ch, 3)   # applied with stride 2 and (0, 1) padding (diffusers Downsample2D(padding=0))


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, n_layers, add_downsample, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, groups, eps) for i in range(n_layers)])
        self.downsamplers = nn.ModuleList([_Down(cout)]) if add_downsample else None


class Encoder(nn.Module):
    """diffusers `Encoder` (SD VAE): conv_in -> 4 down blocks of `layers_per_block` resnets (stride-2 conv after the
    first three) -> mid (resnet, single-head attention, resnet) -> GroupNorm -> SiLU -> conv_out to 2 x latent channels."""

    def __init__(self, in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 norm_num_groups=32, eps=1e-6):
        super().__init__()
        self.conv_in = _Conv(in_channels, block_out_channels[0], 3)
        blocks, cin = [], block_out_channels[0]
        for i, cout in enumerate(block_out_channels):
            blocks.append(_DownBlock(cin, cout, layers_per_block, i != len(block_out_channels) - 1, norm_num_groups, eps))
            cin = cout
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _Mid(cin, norm_num_groups, eps)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, cin, eps=eps)
        self.conv_out = _Conv(cin, 2 * latent_channels, 3)


class DiagonalGaussianDistribution:
    """diffusers' posterior object as far as the reference touches it (`.mean`, EMOAnimationPipeline.py:412)."""

    def __init__(self, moments: torch.Tensor):
        self.mean, self.logvar = torch.chunk(moments, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise


class AutoencoderKLOutput(dict):
    """supports both `.latent_dist` and `['latent_dist']` (the reference indexes it, EMOAnimationPipeline.py:412)"""

    def __init__(self, latent_dist):
        super().__init__(latent_dist=latent_dist)
        self.latent_dist = latent_dist


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class AutoencoderKL(nn.Module):
    """diffusers.AutoencoderKL: `decode` (hot path) and `encode` (reference image -> ReferenceNet latents)."""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, scaling_factor=0.18215):
        super().__init__()
        self.config = AttrDict(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
                               layers_per_block=layers_per_block, latent_channels=latent_channels,
                               norm_num_groups=norm_num_groups, scaling_factor=scaling_factor)
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block, norm_num_groups)
        self.quant_conv = _Conv(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block, norm_num_groups)

    def load_state_dict(self, state_dict, strict=True, **kw):
        # a decoder-only state dict (decoder.* + post_quant_conv.*) stays loadable: the encoder keeps its own weights
        if strict and not any(k.startswith("encoder.") for k in state_dict):
            own = super().state_dict()
            state_dict = {**{k: v for k, v in own.items() if k.startswith(("encoder.", "quant_conv."))}, **state_dict}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def _decode_tokens(self, z: torch.Tensor, pre_scale: float):
        """z [n, 4, h, w] fp32 -> decoder output tokens [n*8h*8w, 4] fp32 (3 valid columns)."""
        if not z.is_cuda:
            raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
        n, cl, h, w = z.shape
        d = self.decoder
        pq = self.post_quant_conv
        cols = ops.latent_im2col(z.float().contiguous().view(n, cl, 1, h, w), pre_scale,
                                 _f32c(pq.weight).view(cl, cl), _f32c(pq.bias))
        wp, b = d.conv_in.packed()
        x = ops.gemm(cols, wp, bias=b, stats_rows=ops.stats_rows_for(h * w, h * w))
        x = d.mid_block.resnets[0].run(x, n, h, w)
        x = d.mid_block.attentions[0].run(x, n, h, w)
        x = d.mid_block.resnets[1].run(x, n, h, w)
        for blk in d.up_blocks:
            for r in blk.resnets:
                x = r.run(x, n, h, w)
            if blk.upsamplers is not None:
                c = x.shape[1]
                up = ops.upsample2x(x, n, h, w, c)
                h, w = 2 * h, 2 * w
                x = blk.upsamplers[0].conv.run(up, n, h, w, stats_rows=ops.stats_rows_for(h * w, h * w))
        g = d.conv_norm_out
        a, _ = ops.group_norm([x], g.num_groups, h * w, n, g.weight, g.bias, g.eps, True)
        out = torch.empty((n * h * w, 4), dtype=F32, device=x.device)
        d.conv_out.run(a, n, h, w, out=out)
        return out, h, w

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """diffusers `AutoencoderKL.encode`: images [n, 3, H, W] in [-1, 1] (H, W multiples of 8) -> posterior over
        [n, 4, H/8, W/8] latents (`.latent_dist.mean`; the caller applies the 0.18215 scaling)."""
        if not x.is_cuda:
            raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
        n, ci, h, w = x.shape
        e = self.encoder
        n_down = sum(1 for b in e.down_blocks if b.downsamplers is not None)
        if h % (2 ** n_down) or w % (2 ** n_down):
            raise ValueError(f"image height/width must be multiples of {2 ** n_down}")
        cols = ops.latent_im2col(x.float().contiguous().view(n, ci, 1, h, w))
        wp, b = e.conv_in.packed()
        t = ops.gemm(cols, wp, bias=b, stats_rows=ops.stats_rows_for(h * w, h * w))
        for blk in e.down_blocks:
            for r in blk.resnets:
                t = r.run(t, n, h, w)
            if blk.downsamplers is not None:
                conv = blk.downsamplers[0].conv
                c = t.shape[1]
                wp, b = conv.packed()
                h, w = h // 2, w // 2
                t = ops.gemm(ops.im2col_s2_pad01(t, n, 2 * h, 2 * w, c), wp, bias=b, stats_rows=ops.stats_rows_for(h * w, h * w))
        t = e.mid_block.resnets[0].run(t, n, h, w)
        t = e.mid_block.attentions[0].run(t, n, h, w)
        t = e.mid_block.resnets[1].run(t, n, h, w)
        g = e.conv_norm_out
        a, _ = ops.group_norm([t], g.num_groups, h * w, n, g.weight, g.bias, g.eps, True)
        m = e.conv_out.run(a, n, h, w)                                   # [n*h*w, 2*latent] fp32
        m = self.quant_conv.run(ops.cast_bf16(m), n, h, w)
        c2 = m.shape[1]
        moments = ops.tokens_to_ncfhw(m.contiguous(), n, c2, 1, h, w).view(n, c2, h, w)
        dist = DiagonalGaussianDistribution(moments)
        return AutoencoderKLOutput(dist) if return_dict else (dist,)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """diffusers `AutoencoderKL.decode`: latents (already divided by the scaling factor) -> images in [-1, 1]."""
        tok, h, w = self._decode_tokens(z, 1.0)
        n = z.shape[0]
        img = ops.tokens_to_ncfhw(tok[:, :3].contiguous(), n, 3, 1, h, w).view(n, 3, h, w)
        return DecoderOutput(sample=img) if return_dict else (img,)

    @torch.no_grad()
    def decode_video(self, latents: torch.Tensor, want_u8: bool = False, frame_chunk: Optional[int] = None):
        """[b, 4, f, h, w] scaled latents -> ([b, 3, f, 8h, 8w] fp32 in [0,1], optional uint8 copy).
        Fuses `1/0.18215 *`, the frame batching and `(x/2+0.5).clamp(0,1)` (EMOAnimationPipeline.py:293-304)."""
        b, c, f, h, w = latents.shape
        z = latents.float().permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        # frames per decoder pass: bounded so a long-form clip (240 frames) never holds more than one chunk of 512x512
        # decoder activations (~0.6 GB per frame); the reference decodes frame by frame (EMOAnimationPipeline.py:297-301)
        chunk = frame_chunk or min(b * f, 16)
        outs_f, outs_u = [], []
        for s in range(0, b * f, chunk):
            zc = z[s:s + chunk].contiguous()
            tok, ho, wo = self._decode_tokens(zc, 1.0 / self.config.scaling_factor)
            of, ou = ops.vae_postprocess(tok, zc.shape[0], ho, wo, want_f32=True, want_u8=want_u8)
            outs_f.append(of)
            if want_u8:
                outs_u.append(ou)
        vf = torch.cat(outs_f) if len(outs_f) > 1 else outs_f[0]
        video = vf.view(b, f, 3, vf.shape[-2], vf.shape[-1]).permute(0, 2, 1, 3, 4)
        vu = None
        if want_u8:
            vu = torch.cat(outs_u) if len(outs_u) > 1 else outs_u[0]
            vu = vu.view(b, f, 3, vu.shape[-2], vu.shape[-1]).permute(0, 2, 1, 3, 4)
        return video, vu
