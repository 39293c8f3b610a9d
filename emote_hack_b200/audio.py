"""Audio front-end of the hot path on the sm_100a kernels:

* `Wav2Vec2Model` — the wav2vec2-base forward the reference runs through `transformers.Wav2Vec2Model`
  (`Wav2VecFeatureExtractor.__init__`, Net.py:607-612; forward at :644): 7 strided 1-D convolutions (GroupNorm + GELU after
  the first, GELU after the others), LayerNorm + Linear feature projection, grouped positional convolution (weight norm,
  kernel 128, 16 groups) + GELU, LayerNorm, 12 post-LN transformer layers (12 heads x 64, GELU feed-forward 768 -> 3072).
  Same module tree / state_dict keys as the transformers class, so a `facebook/wav2vec2-base-960h` checkpoint loads
  unchanged (both the `parametrizations.weight.original0/1` and the older `weight_g/weight_v` spellings of the weight norm).
* `Wav2VecFeatureExtractor` — Net.py:607-667: waveform -> normalise (the processor's zero-mean / unit-variance) -> model ->
  the per-frame windows of m frames before / n after, as the reference's flat [T, (m+n+1)*768] or as [T, m+n+1, 768] tokens
  (the per-frame `encoder_hidden_states` of the audio cross-attention).
* `SpeedEncoder` — Net.py:198-258: tanh bucket encoding of head rotation speeds + Linear -> ReLU -> Linear.

Kernel plan: the strided convolutions are zero-copy GEMMs (row t of the operand = tokens stride*t .. stride*t + k - 1 of
the previous layer's [T, C] output: overlapping rows, `lda = stride*C < K`), GELU fused in the epilogue; the grouped
positional convolution is 16 such GEMMs over a group-major zero-padded copy; attention = `emote_attention_bf16` (d = 64).
There is no CPU path: everything raises on CPU tensors like the rest of the package.
"""
from __future__ import annotations

import math
import wave as _wave
from typing import List, Optional

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import EmoteKernelError
from .unet3d import AttrDict, _f32c, _sig

F32, OP16 = torch.float32, ops.OP16


class _ConvLayer(nn.Module):
    """transformers Wav2Vec2GroupNormConvLayer / Wav2Vec2NoLayerNormConvLayer (parameter container)"""

    def __init__(self, cin, cout, kernel, stride, bias, group_norm):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=kernel, stride=stride, bias=bias)
        if group_norm:
            self.layer_norm = nn.GroupNorm(num_groups=cout, num_channels=cout, affine=True)


class _FeatureEncoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        dims = (1,) + tuple(cfg.conv_dim)
        self.conv_layers = nn.ModuleList([
            _ConvLayer(dims[i], dims[i + 1], cfg.conv_kernel[i], cfg.conv_stride[i], cfg.conv_bias, group_norm=(i == 0))
            for i in range(len(cfg.conv_dim))])


class _FeatureProjection(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layer_norm = nn.LayerNorm(cfg.conv_dim[-1], eps=cfg.layer_norm_eps)
        self.projection = nn.Linear(cfg.conv_dim[-1], cfg.hidden_size)


class _PosConvEmbed(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        conv = nn.Conv1d(cfg.hidden_size, cfg.hidden_size, kernel_size=cfg.num_conv_pos_embeddings,
                         padding=cfg.num_conv_pos_embeddings // 2, groups=cfg.num_conv_pos_embedding_groups)
        self.conv = nn.utils.parametrizations.weight_norm(conv, name="weight", dim=2)


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (nn.Linear(dim, dim) for _ in range(4))


class _FeedForward(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.intermediate_dense = nn.Linear(cfg.hidden_size, cfg.intermediate_size)
        self.output_dense = nn.Linear(cfg.intermediate_size, cfg.hidden_size)


class _EncoderLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.attention = _Attention(cfg.hidden_size)
        self.layer_norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.feed_forward = _FeedForward(cfg)
        self.final_layer_norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.pos_conv_embed = _PosConvEmbed(cfg)
        self.layer_norm = nn.LayerNorm(cfg.hidden_size, eps=cfg.layer_norm_eps)
        self.layers = nn.ModuleList([_EncoderLayer(cfg) for _ in range(cfg.num_hidden_layers)])


class Wav2Vec2Model(nn.Module):
    """transformers.Wav2Vec2Model (wav2vec2-base geometry by default) for inference on one waveform."""

    def __init__(self, conv_dim=(512,) * 7, conv_stride=(5, 2, 2, 2, 2, 2, 2), conv_kernel=(10, 3, 3, 3, 3, 2, 2),
                 conv_bias: bool = False, feat_extract_norm: str = "group", hidden_size: int = 768,
                 num_hidden_layers: int = 12, num_attention_heads: int = 12, intermediate_size: int = 3072,
                 num_conv_pos_embeddings: int = 128, num_conv_pos_embedding_groups: int = 16, layer_norm_eps: float = 1e-5,
                 do_stable_layer_norm: bool = False):
        super().__init__()
        if feat_extract_norm != "group" or do_stable_layer_norm or conv_bias:
            raise NotImplementedError("Wav2Vec2Model: only the wav2vec2-base layout (group norm, post-LN, bias-free convs)")
        if hidden_size % num_attention_heads or (hidden_size // num_attention_heads) % 8 or hidden_size % 64 or conv_dim[-1] % 64:
            raise NotImplementedError("Wav2Vec2Model: widths must be multiples of 64 and head_dim a multiple of 8")
        self.config = AttrDict({k: v for k, v in locals().items() if k not in ("self", "__class__")})
        cfg = self.config
        self.masked_spec_embed = nn.Parameter(torch.zeros(hidden_size).uniform_())   # checkpoint key; unused at inference
        self.feature_extractor = _FeatureEncoder(cfg)
        self.feature_projection = _FeatureProjection(cfg)
        self.encoder = _Encoder(cfg)
        self._pk = None
        self._register_load_state_dict_pre_hook(self._rename_legacy_weight_norm)

    @staticmethod
    def _rename_legacy_weight_norm(state_dict, prefix, *a):
        for old, new in (("weight_g", "parametrizations.weight.original0"), ("weight_v", "parametrizations.weight.original1")):
            k = f"{prefix}encoder.pos_conv_embed.conv.{old}"
            if k in state_dict:
                state_dict[f"{prefix}encoder.pos_conv_embed.conv.{new}"] = state_dict.pop(k)

    @property
    def device(self):
        return self.masked_spec_embed.device

    def frames_for(self, n_samples: int) -> int:
        """number of output frames for a waveform of n samples (transformers `_get_feat_extract_output_lengths`)"""
        t = n_samples
        for k, s in zip(self.config.conv_kernel, self.config.conv_stride):
            t = (t - k) // s + 1
        return t

    # -- packed 16-bit weights (rebuilt when a parameter changed) ----------------------------------------------------
    def _packed(self):
        sig = _sig(*self.parameters())
        if self._pk is not None and self._pk["sig"] == sig:
            return self._pk
        cfg = self.config
        with torch.no_grad():
            p = {"sig": sig, "conv": [], "layers": []}
            w0 = self.feature_extractor.conv_layers[0].conv.weight                        # [C, 1, k]
            k0 = w0.shape[-1]
            kpad = (k0 + 7) // 8 * 8
            w0p = torch.zeros(w0.shape[0], kpad, dtype=OP16, device=w0.device)
            w0p[:, :k0] = w0[:, 0].to(OP16)
            p["w0"], p["k0pad"] = w0p, kpad
            for layer in self.feature_extractor.conv_layers[1:]:
                w = layer.conv.weight                                                     # [out, in, k] -> [out, (j, c)]
                p["conv"].append(w.permute(0, 2, 1).reshape(w.shape[0], -1).to(OP16).contiguous())
            fp = self.feature_projection
            p["wp"], p["bp"] = ops.pack_linear(fp.projection.weight), _f32c(fp.projection.bias)
            pc = self.encoder.pos_conv_embed.conv
            w = pc.weight                                                                 # weight norm applied: [C, C/G, k]
            g, cg = cfg.num_conv_pos_embedding_groups, cfg.hidden_size // cfg.num_conv_pos_embedding_groups
            p["wpos"] = [w[i * cg:(i + 1) * cg].permute(0, 2, 1).reshape(cg, -1).to(OP16).contiguous() for i in range(g)]
            p["bpos"] = _f32c(pc.bias)
            for lyr in self.encoder.layers:
                a, ff = lyr.attention, lyr.feed_forward
                p["layers"].append({
                    "wqkv": ops.pack_linear(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                    "bqkv": _f32c(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)),
                    "wo": ops.pack_linear(a.out_proj.weight), "bo": _f32c(a.out_proj.bias),
                    "w1": ops.pack_linear(ff.intermediate_dense.weight), "b1": _f32c(ff.intermediate_dense.bias),
                    "w2": ops.pack_linear(ff.output_dense.weight), "b2": _f32c(ff.output_dense.bias)})
        self._pk = p
        return p

    @torch.no_grad()
    def forward(self, input_values: torch.Tensor, normalize: bool = False, return_dict: bool = True):
        """input_values: waveform [n] or [1, n] fp32 on CUDA (already normalised, like the processor's `input_values`, unless
        normalize=True) -> last_hidden_state [1, T, hidden] fp32."""
        if not input_values.is_cuda:
            raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
        if input_values.dim() == 2:
            if input_values.shape[0] != 1:
                raise NotImplementedError("Wav2Vec2Model.forward: one waveform per call")
            input_values = input_values[0]
        wave = input_values.float().contiguous()
        cfg, p = self.config, self._packed()
        fe = self.feature_extractor.conv_layers
        if wave.numel() < cfg.conv_kernel[0] or self.frames_for(wave.numel()) < 1:
            raise ValueError("waveform too short for the convolutional feature encoder")
        stats = ops.wave_stats(wave) if normalize else None
        a0 = ops.wave_im2col(wave, stats, cfg.conv_kernel[0], cfg.conv_stride[0], p["k0pad"])
        h = ops.gemm(a0, p["w0"])                                                        # [T1, C] fp32
        gn = fe[0].layer_norm
        x = ops.channel_norm_gelu(h, gn.weight, gn.bias, gn.eps)                         # op16
        t, c = x.shape
        n_conv = len(p["conv"])
        for i, w in enumerate(p["conv"]):
            k, s = cfg.conv_kernel[i + 1], cfg.conv_stride[i + 1]
            t = (t - k) // s + 1
            # Conv1d(stride s, kernel k) over [T, C] tokens as a GEMM over overlapping rows — no im2col copy
            x = ops.gemm(x, w, M=t, lda=s * c, gelu=True, out_dtype=F32 if i == n_conv - 1 else OP16)
            c = w.shape[0]
        fp = self.feature_projection
        ln = ops.layer_norm(x, fp.layer_norm.weight, fp.layer_norm.bias, fp.layer_norm.eps)
        hidden = ops.gemm(ln, p["wp"], bias=p["bp"])                                     # [T, hidden] fp32
        # positional convolution: Conv1d(k, padding k/2, groups) -> drop the last frame (even k) -> GELU
        kpos, g = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
        cg = cfg.hidden_size // g
        xg = ops.tokens_to_groups(hidden, g, kpos // 2, kpos // 2)                       # [g, T + k, cg] op16
        pos = torch.empty_like(hidden)
        for i in range(g):
            ops.gemm(xg[i], p["wpos"][i], M=t, lda=cg, bias=p["bpos"][i * cg:(i + 1) * cg], gelu=True, out=pos, out_col=i * cg)
        hidden = ops.add_f32(hidden, pos)
        enc = self.encoder
        h16, h32 = ops.layer_norm_dual(hidden, enc.layer_norm.weight, enc.layer_norm.bias, enc.layer_norm.eps)
        heads, hd, d = cfg.num_attention_heads, cfg.hidden_size // cfg.num_attention_heads, cfg.hidden_size
        for lyr, lp in zip(enc.layers, p["layers"]):
            qkv = ops.gemm(h16, lp["wqkv"], bias=lp["bqkv"], out_dtype=OP16)
            att = torch.empty((t, d), dtype=OP16, device=wave.device)
            ops.attention(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], att, batch=1, heads=heads, head_dim=hd, nq=t, n0=t,
                          q_strides=(t * 3 * d, 3 * d), kv0_strides=(t * 3 * d, 3 * d), o_strides=(t * d, d), scale=hd ** -0.5)
            x = ops.gemm(att, lp["wo"], bias=lp["bo"], residual=h32)
            h16, h32 = ops.layer_norm_dual(x, lyr.layer_norm.weight, lyr.layer_norm.bias, lyr.layer_norm.eps)
            mid = ops.gemm(h16, lp["w1"], bias=lp["b1"], gelu=True, out_dtype=OP16)
            x = ops.gemm(mid, lp["w2"], bias=lp["b2"], residual=h32)
            h16, h32 = ops.layer_norm_dual(x, lyr.final_layer_norm.weight, lyr.final_layer_norm.bias, lyr.final_layer_norm.eps)
        out = h32[None]
        return AttrDict(last_hidden_state=out) if return_dict else (out,)


def window_features(hidden_states: torch.Tensor, m: int = 2, n: int = 2, as_tokens: bool = True) -> torch.Tensor:
    """Neighbour-frame windows of wav2vec2 hidden states (Net.py:646-667): frame f gets the features of frames f-m .. f+n,
    zero-padded past either end.  [T, d] (or [1, T, d]) -> [T, m+n+1, d] tokens, or the reference's flat [T, (m+n+1)*d]."""
    h = hidden_states[0] if hidden_states.dim() == 3 else hidden_states
    if h.dim() != 2 or m < 0 or n < 0:
        raise ValueError("window_features: expected hidden states [T, d] and m, n >= 0")
    t, d = h.shape
    padded = torch.cat([h.new_zeros(m, d), h, h.new_zeros(n, d)])
    idx = torch.arange(t, device=h.device)[:, None] + torch.arange(m + n + 1, device=h.device)[None]
    win = padded[idx]                                            # [T, m+n+1, d]
    return win if as_tokens else win.reshape(t, (m + n + 1) * d)


def read_wav(path: str):
    """PCM wav -> (float32 mono waveform in [-1, 1], sample rate); stdlib only (the reference uses soundfile)."""
    with _wave.open(path, "rb") as f:
        sr, nch, width, n = f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()
        raw = f.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f"unsupported sample width {width}")
    if nch > 1:
        x = x.reshape(-1, nch).mean(axis=1)                      # Net.py:633-634
    return x, sr


class Wav2VecFeatureExtractor:
    """Net.py:607-667.  `model`: an `audio.Wav2Vec2Model` (or None: built with the wav2vec2-base geometry; weights come
    from `model_name` when it is a local directory transformers can load offline, else stay randomly initialised)."""

    sampling_rate = 16000

    def __init__(self, model_name: str = "facebook/wav2vec2-base-960h", device="cuda", model: Optional[Wav2Vec2Model] = None):
        self.model_name, self.device = model_name, torch.device(device)
        if model is None:
            model = Wav2Vec2Model()
            try:
                import os
                if os.path.isdir(model_name):
                    import transformers
                    sd = transformers.Wav2Vec2Model.from_pretrained(model_name, local_files_only=True).state_dict()
                    model.load_state_dict(sd, strict=False)
            except Exception as e:  # no checkpoint on disk: the caller loads a state_dict later
                raise RuntimeError(f"could not load wav2vec2 weights from {model_name}: {e}") from e
        self.model = model.to(self.device).eval()

    def _waveform(self, audio, sample_rate: Optional[int]):
        if isinstance(audio, str):
            audio, sample_rate = read_wav(audio)
        x = torch.as_tensor(np.asarray(audio.detach().cpu()) if torch.is_tensor(audio) else np.asarray(audio), dtype=torch.float32)
        if x.dim() > 1:
            x = x.mean(dim=-1) if x.shape[-1] <= 8 else x.reshape(-1)
        sr = sample_rate or self.sampling_rate
        if sr != self.sampling_rate:
            # linear-interpolation resampling (the reference calls librosa.resample, Net.py:628-630: not available here)
            n_out = int(round(x.numel() * self.sampling_rate / sr))
            pos = torch.linspace(0, x.numel() - 1, n_out)
            lo = pos.floor().long().clamp(max=x.numel() - 1)
            hi = (lo + 1).clamp(max=x.numel() - 1)
            x = x[lo] + (pos - lo.float()) * (x[hi] - x[lo])
        return x.to(self.device)

    @torch.no_grad()
    def hidden_states(self, audio, sample_rate: Optional[int] = None) -> torch.Tensor:
        """waveform (array / tensor / wav path) -> wav2vec2 last_hidden_state [T, 768] (50 frames per second)"""
        return self.model(self._waveform(audio, sample_rate), normalize=True).last_hidden_state[0]

    def extract_tokens(self, audio, m: int = 2, n: int = 2, sample_rate: Optional[int] = None) -> torch.Tensor:
        """per-frame tokens [T, m+n+1, 768] = the audio cross-attention context of frame t"""
        return window_features(self.hidden_states(audio, sample_rate), m, n, as_tokens=True)

    def extract_features_from_wav(self, audio_path, m: int = 2, n: int = 2) -> torch.Tensor:
        """Net.py:614-667: flat per-frame features [T, (m+n+1)*768]"""
        return window_features(self.hidden_states(audio_path), m, n, as_tokens=False)

    def extract_features_from_mp4(self, video_path, m: int = 2, n: int = 2) -> torch.Tensor:
        """Net.py:670-732: uses the `.wav` next to the video (the reference extracts it with moviepy when missing — video
        I/O is out of scope here)."""
        import os
        audio_path = os.path.splitext(video_path)[0] + ".wav"
        if not os.path.exists(audio_path):
            raise FileNotFoundError(f"{audio_path}: extract the audio track first (moviepy is not part of this package)")
        return self.extract_features_from_wav(audio_path, m, n)


class SpeedEncoder(nn.Module):
    """Net.py:198-258 — head-rotation speed -> bucket vector tanh((s - c_i)/r_i * 3) -> Linear -> ReLU -> Linear.
    Same constructor checks as the reference (its 9 hard-coded bucket centres make num_speed_buckets = 9 the only value
    that passes, Net.py:225-229 vs :214)."""

    def __init__(self, num_speed_buckets: int, speed_embedding_dim: int):
        super().__init__()
        assert isinstance(num_speed_buckets, int), "num_speed_buckets must be an integer"
        assert num_speed_buckets > 0, "num_speed_buckets must be positive"
        assert isinstance(speed_embedding_dim, int), "speed_embedding_dim must be an integer"
        assert speed_embedding_dim > 0, "speed_embedding_dim must be positive"
        self.num_speed_buckets, self.speed_embedding_dim = num_speed_buckets, speed_embedding_dim
        self.bucket_centers = self.get_bucket_centers()
        self.bucket_radii = self.get_bucket_radii()
        assert len(self.bucket_centers) == self.num_speed_buckets, "bucket_centers length must match num_speed_buckets"
        assert len(self.bucket_radii) == self.num_speed_buckets, "bucket_radii length must match num_speed_buckets"
        self.mlp = nn.Sequential(nn.Linear(num_speed_buckets, speed_embedding_dim), nn.ReLU(),
                                 nn.Linear(speed_embedding_dim, speed_embedding_dim))

    def get_bucket_centers(self) -> List[float]:
        return [-1.0, -0.5, -0.2, -0.1, 0.0, 0.1, 0.2, 0.5, 1.0]

    def get_bucket_radii(self) -> List[float]:
        return [0.1] * self.num_speed_buckets

    @torch.no_grad()
    def forward(self, head_rotation_speeds) -> torch.Tensor:
        s = torch.as_tensor(head_rotation_speeds)
        assert s.ndim == 1, "head_rotation_speeds must be a 1D tensor"
        assert s.dtype == torch.float32, "head_rotation_speeds must be a tensor of floats"
        dev = self.mlp[0].weight.device
        if dev.type != "cuda":
            raise EmoteKernelError("emote_hack_b200 modules run on CUDA only (no CPU fallback)")
        s = s.to(dev).contiguous()
        centers = torch.tensor(self.bucket_centers, dtype=F32, device=dev)
        radii = torch.tensor(self.bucket_radii, dtype=F32, device=dev)
        l1, l2 = self.mlp[0], self.mlp[2]
        return ops.speed_encoder(s, centers, radii, _f32c(l1.weight), _f32c(l1.bias), _f32c(l2.weight), _f32c(l2.bias))
