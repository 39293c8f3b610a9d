"""TEST INFRASTRUCTURE ONLY — CPU/fp32 restatement ("port") of the reference UNet3DConditionModel forward.

A functional, state-dict driven re-derivation of the arithmetic of
  magicanimate/models/unet_controlnet.py:328-483   (UNet3DConditionModel.forward)
  magicanimate/models/unet_3d_blocks.py:276-283,384-423,491-519,616-662,726-751 (block forwards)
  magicanimate/models/resnet.py:31-38,56-84,102-110,177-207
  magicanimate/models/attention.py:112-161,276-320
  magicanimate/models/mutual_self_attention.py:199-284 (reader hook)
  magicanimate/models/motion_module.py:139-163,215-227,246-248,275-334
  magicanimate/models/orig_attention.py:598-684,778-781,825-827
  magicanimate/models/embeddings.py:28-68,206-218
It travels to the GPU box (the reference itself does not) and is the checker the CUDA path is compared with.
PINNED: tests/test_oracle.py checks it (a) directly against the untouched reference modules when
/root/reference is present and (b) against tests/golden/*.pt generated from the reference by oracle/make_golden.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this file.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def timestep_embedding(t: Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float) -> Tensor:
    # embeddings.py:28-68 with scale=1, max_period=10000
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - freq_shift))
    ang = t[:, None].float() * freqs[None]
    emb = torch.cat([ang.sin(), ang.cos()], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2:
        emb = F.pad(emb, (0, 1))
    return emb


class UNet3DOracle:
    """Evaluates the reference network from a reference-format state_dict (same key names) and its config."""

    def __init__(self, state_dict: Dict[str, Tensor], config: dict, dtype=torch.float32, device=None,
                 attention_slice_bytes: Optional[int] = None):
        # dtype=float32 is the oracle; bfloat16 is only used to calibrate "what eager PyTorch bf16 would give"
        self.dtype = dtype
        self.sd = {k: v.detach().to(device=device or v.device, dtype=dtype) for k, v in state_dict.items()}
        c = dict(config)
        self.groups = c.get("norm_num_groups", 32)
        self.eps = c.get("norm_eps", 1e-5)
        self.flip = c.get("flip_sin_to_cos", True)
        self.shift = c.get("freq_shift", 0)
        self.time_dim = c.get("block_out_channels", (320, 640, 1280, 1280))[0]
        hd = c.get("attention_head_dim", 8)
        n_blocks = len(c.get("block_out_channels", (320, 640, 1280, 1280)))
        self.heads_down = list(hd) if isinstance(hd, (list, tuple)) else [hd] * n_blocks
        self.mm_heads = (c.get("motion_module_kwargs") or {}).get("num_attention_heads", 8)
        self.mid_scale = c.get("mid_block_scale_factor", 1)
        self.center = c.get("center_input_sample", False)
        self.class_embed_type = c.get("class_embed_type")
        self.n_down = n_blocks
        self._collect = None
        # like unet.set_attention_slice (unet_controlnet.py:259-322 -> orig_attention.py:686-727): the score matrix is
        # evaluated in slices over the batch axis when it would exceed this many bytes; same arithmetic per slice
        self.attention_slice_bytes = attention_slice_bytes

    # ------------------------------------------------------------------ primitives
    def _has(self, key: str) -> bool:
        return key in self.sd

    def _lin(self, p: str, x: Tensor) -> Tensor:
        return F.linear(x, self.sd[p + ".weight"], self.sd.get(p + ".bias"))

    def _conv5(self, p: str, x: Tensor, stride: int = 1) -> Tensor:
        # InflatedConv3d: a 2-D conv applied to every frame (resnet.py:31-38)
        b, c, f, h, w = x.shape
        w_ = self.sd[p + ".weight"]
        y = F.conv2d(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w), w_, self.sd.get(p + ".bias"), stride=stride,
                     padding=w_.shape[-1] // 2)
        return y.reshape(b, f, *y.shape[1:]).permute(0, 2, 1, 3, 4)

    def _gn(self, p: str, x: Tensor, eps: float) -> Tensor:
        return F.group_norm(x, self.groups, self.sd[p + ".weight"], self.sd[p + ".bias"], eps)

    def _ln(self, p: str, x: Tensor) -> Tensor:
        return F.layer_norm(x, (x.shape[-1],), self.sd[p + ".weight"], self.sd[p + ".bias"], 1e-5)

    def _mha(self, p: str, x: Tensor, ctx: Tensor, heads: int) -> Tensor:
        # orig_attention.py:598-684: softmax(q k^T d^-1/2) v, biased out projection
        q, k, v = self._lin(p + ".to_q", x), self._lin(p + ".to_k", ctx), self._lin(p + ".to_v", ctx)
        b, n, c = q.shape
        d = c // heads
        sp = lambda t: t.reshape(t.shape[0], t.shape[1], heads, d).transpose(1, 2)
        qh, kh, vh = sp(q), sp(k), sp(v)
        step = b
        if self.attention_slice_bytes:
            per_batch = heads * n * k.shape[1] * q.element_size()
            step = max(1, min(b, self.attention_slice_bytes // max(1, per_batch)))
        outs = []
        for i in range(0, b, step):
            s = (qh[i:i + step] @ kh[i:i + step].transpose(-1, -2)) * d ** -0.5
            outs.append(s.softmax(-1) @ vh[i:i + step])
            del s
        o = (outs[0] if len(outs) == 1 else torch.cat(outs)).transpose(1, 2).reshape(b, n, c)
        return self._lin(p + ".to_out.0", o)

    def _ff(self, p: str, x: Tensor) -> Tensor:
        h, gate = self._lin(p + ".net.0.proj", x).chunk(2, dim=-1)  # GEGLU, exact erf GELU
        return self._lin(p + ".net.2", h * F.gelu(gate))

    # ------------------------------------------------------------------ blocks
    def _resnet(self, p: str, x: Tensor, emb: Tensor, scale: float = 1.0) -> Tensor:
        h = self._conv5(p + ".conv1", F.silu(self._gn(p + ".norm1", x, self.eps)))  # 5-D GroupNorm couples frames
        h = h + self._lin(p + ".time_emb_proj", F.silu(emb))[:, :, None, None, None]
        h = self._conv5(p + ".conv2", F.silu(self._gn(p + ".norm2", h, self.eps)))
        if self._has(p + ".conv_shortcut.weight"):
            x = self._conv5(p + ".conv_shortcut", x)
        return (x + h) / scale

    def _basic_block(self, p: str, x: Tensor, ctx: Tensor, heads: int, bank: Optional[Sequence[Tensor]], frames: int,
                     cfg: bool) -> Tensor:
        n1 = self._ln(p + ".norm1", x)
        if self._collect is not None:
            # writer hook (mutual_self_attention.py:226-232): the block's LayerNorm1 output is appended to its bank
            self._collect.setdefault(p, []).append(n1.clone())
        if not self._has(p + ".attn1.to_q.weight"):
            return x  # AppearanceEncoderModel's last block keeps only norm1 (appearance_encoder.py:613-621)
        if bank:
            # reader hook (mutual_self_attention.py:237-258): keys/values = [self | bank]; the unconditional first
            # half of the CFG batch is recomputed without the bank
            rep = [d.unsqueeze(1).repeat(1, frames, 1, 1).flatten(0, 1)[: x.shape[0]] for d in bank]
            with_ref = self._mha(p + ".attn1", n1, torch.cat([n1] + rep, dim=1), heads) + x
            if cfg:
                half = x.shape[0] // 2
                plain = self._mha(p + ".attn1", n1[:half], n1[:half], heads) + x[:half]
                x = torch.cat([plain, with_ref[half:]], dim=0)
            else:
                x = with_ref
        else:
            x = self._mha(p + ".attn1", n1, n1, heads) + x
        if self._has(p + ".attn2.to_q.weight"):
            x = self._mha(p + ".attn2", self._ln(p + ".norm2", x), ctx, heads) + x
        return self._ff(p + ".ff", self._ln(p + ".norm3", x)) + x

    def _transformer3d(self, p: str, x: Tensor, ctx: Tensor, heads: int, banks, cfg: bool) -> Tensor:
        b, c, f, h, w = x.shape
        x2 = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        if ctx.shape[0] != b * f:  # attention.py:118-119
            ctx = ctx.repeat_interleave(f, dim=0)
        t = F.group_norm(x2, self.groups, self.sd[p + ".norm.weight"], self.sd[p + ".norm.bias"], 1e-6)
        wi = self.sd[p + ".proj_in.weight"]
        if wi.dim() == 4:  # 1x1 conv projection
            t = F.conv2d(t, wi, self.sd[p + ".proj_in.bias"]).permute(0, 2, 3, 1).reshape(b * f, h * w, -1)
        else:
            t = self._lin(p + ".proj_in", t.permute(0, 2, 3, 1).reshape(b * f, h * w, c))
        i = 0
        while self._has(f"{p}.transformer_blocks.{i}.norm1.weight"):
            bp = f"{p}.transformer_blocks.{i}"
            t = self._basic_block(bp, t, ctx, heads, (banks or {}).get(bp), f, cfg)
            i += 1
        if not self._has(p + ".proj_out.weight"):
            return x  # trimmed writer tail: nothing downstream of the bank is computed
        wo = self.sd[p + ".proj_out.weight"]
        if wo.dim() == 4:
            t = F.conv2d(t.reshape(b * f, h, w, -1).permute(0, 3, 1, 2), wo, self.sd[p + ".proj_out.bias"])
        else:
            t = self._lin(p + ".proj_out", t).reshape(b * f, h, w, -1).permute(0, 3, 1, 2)
        return (t + x2).reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)

    def _motion(self, p: str, x: Tensor) -> Tensor:
        # VanillaTemporalModule -> TemporalTransformer3DModel (motion_module.py:139-163)
        p = p + ".temporal_transformer"
        b, c, f, h, w = x.shape
        x2 = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        t = F.group_norm(x2, self.groups, self.sd[p + ".norm.weight"], self.sd[p + ".norm.bias"], 1e-6)
        t = self._lin(p + ".proj_in", t.permute(0, 2, 3, 1).reshape(b * f, h * w, c))
        i = 0
        while self._has(f"{p}.transformer_blocks.{i}.ff_norm.weight"):
            bp = f"{p}.transformer_blocks.{i}"
            j = 0
            while self._has(f"{bp}.attention_blocks.{j}.to_q.weight"):
                ap = f"{bp}.attention_blocks.{j}"
                n = self._ln(f"{bp}.norms.{j}", t)
                # (b f) d c -> (b d) f c, add sinusoidal table, self-attend over frames (motion_module.py:280-332)
                seq = n.reshape(b, f, h * w, c).transpose(1, 2).reshape(b * h * w, f, c)
                if self._has(ap + ".pos_encoder.pe"):
                    seq = seq + self.sd[ap + ".pos_encoder.pe"][:, :f]
                o = self._mha(ap, seq, seq, self.mm_heads)
                t = o.reshape(b, h * w, f, c).transpose(1, 2).reshape(b * f, h * w, c) + t
                j += 1
            t = self._ff(bp + ".ff", self._ln(bp + ".ff_norm", t)) + t
            i += 1
        t = self._lin(p + ".proj_out", t).reshape(b * f, h, w, c).permute(0, 3, 1, 2)
        return (t + x2).reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)

    def _layers(self, p: str) -> int:
        i = 0
        while self._has(f"{p}.resnets.{i}.norm1.weight"):
            i += 1
        return i

    # ------------------------------------------------------------------ whole network
    @torch.no_grad()
    def forward(self, sample: Tensor, timestep, encoder_hidden_states: Tensor, banks: Optional[Dict[str, List[Tensor]]] = None,
                do_classifier_free_guidance: bool = True, down_block_additional_residuals=None,
                mid_block_additional_residual=None, collect_banks: Optional[Dict[str, List[Tensor]]] = None,
                class_labels: Optional[Tensor] = None) -> Tensor:
        """`collect_banks` (a dict, filled in place): run as a ReferenceNet WRITER — every BasicTransformerBlock appends
        its LayerNorm1 output under its module name (the caller keeps the mid / up entries, fusion_blocks='midup')."""
        self._collect = collect_banks
        try:
            return self._forward(sample, timestep, encoder_hidden_states, banks, do_classifier_free_guidance,
                                 down_block_additional_residuals, mid_block_additional_residual, class_labels)
        finally:
            self._collect = None

    def _forward(self, sample, timestep, encoder_hidden_states, banks, do_classifier_free_guidance,
                 down_block_additional_residuals, mid_block_additional_residual, class_labels=None) -> Tensor:
        x = sample.to(self.dtype)
        ctx = encoder_hidden_states.to(self.dtype)
        if self.center:
            x = 2 * x - 1.0
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=x.device)
        t = t.reshape(-1).to(x.device).expand(x.shape[0])
        emb = timestep_embedding(t, self.time_dim, self.flip, self.shift).to(self.dtype)
        emb = self._lin("time_embedding.linear_2", F.silu(self._lin("time_embedding.linear_1", emb)))
        # class embedding (unet_controlnet.py:400-408): Embedding table / TimestepEmbedding of the projected labels / identity
        if self._has("class_embedding.weight"):
            emb = emb + self.sd["class_embedding.weight"][class_labels.reshape(-1).long()]
        elif self._has("class_embedding.linear_1.weight"):
            ce = timestep_embedding(class_labels.reshape(-1).to(x.device).expand(x.shape[0]), self.time_dim, self.flip,
                                    self.shift).to(self.dtype)
            emb = emb + self._lin("class_embedding.linear_2", F.silu(self._lin("class_embedding.linear_1", ce)))
        elif self.class_embed_type == "identity":
            emb = emb + class_labels.to(self.dtype)
        # the interpolation output size is forced when the latent is not a multiple of 2**num_upsamplers (:355-364, 458-460)
        n_up = sum(1 for bi in range(self.n_down) if self._has(f"up_blocks.{bi}.upsamplers.0.conv.weight"))
        force_size = any(s % (2 ** n_up) != 0 for s in x.shape[-2:])

        x = self._conv5("conv_in", x)
        skips = [x]
        for bi in range(self.n_down):
            p = f"down_blocks.{bi}"
            for li in range(self._layers(p)):
                x = self._resnet(f"{p}.resnets.{li}", x, emb)
                if self._has(f"{p}.attentions.{li}.norm.weight"):
                    x = self._transformer3d(f"{p}.attentions.{li}", x, ctx, self.heads_down[bi], banks,
                                            do_classifier_free_guidance)
                if self._has(f"{p}.motion_modules.{li}.temporal_transformer.norm.weight"):
                    x = self._motion(f"{p}.motion_modules.{li}", x)
                skips.append(x)
            if self._has(f"{p}.downsamplers.0.conv.weight"):
                x = self._conv5(f"{p}.downsamplers.0.conv", x, stride=2)
                skips.append(x)
        if down_block_additional_residuals is not None and mid_block_additional_residual is not None:
            skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]

        p = "mid_block"
        x = self._resnet(f"{p}.resnets.0", x, emb, self.mid_scale)
        for li in range(self._layers(p) - 1):
            x = self._transformer3d(f"{p}.attentions.{li}", x, ctx, self.heads_down[-1], banks, do_classifier_free_guidance)
            if self._has(f"{p}.motion_modules.{li}.temporal_transformer.norm.weight"):
                x = self._motion(f"{p}.motion_modules.{li}", x)
            x = self._resnet(f"{p}.resnets.{li + 1}", x, emb, self.mid_scale)
        if down_block_additional_residuals is not None and mid_block_additional_residual is not None:
            x = x + mid_block_additional_residual

        heads_up = list(reversed(self.heads_down))
        for bi in range(self.n_down):
            p = f"up_blocks.{bi}"
            for li in range(self._layers(p)):
                x = torch.cat([x, skips.pop()], dim=1)
                x = self._resnet(f"{p}.resnets.{li}", x, emb)
                if self._has(f"{p}.attentions.{li}.norm.weight"):
                    x = self._transformer3d(f"{p}.attentions.{li}", x, ctx, heads_up[bi], banks, do_classifier_free_guidance)
                if self._has(f"{p}.motion_modules.{li}.temporal_transformer.norm.weight"):
                    x = self._motion(f"{p}.motion_modules.{li}", x)
            if self._has(f"{p}.upsamplers.0.conv.weight"):
                if force_size and bi != self.n_down - 1:
                    x = F.interpolate(x, size=skips[-1].shape[2:], mode="nearest")  # resnet.py:76
                else:
                    x = F.interpolate(x, scale_factor=(1.0, 2.0, 2.0), mode="nearest")  # resnet.py:74
                x = self._conv5(f"{p}.upsamplers.0.conv", x)

        if not self._has("conv_out.weight"):
            return x  # AppearanceEncoderModel has no output head
        x = F.silu(self._gn("conv_norm_out", x, self.eps))
        return self._conv5("conv_out", x)

    __call__ = forward
