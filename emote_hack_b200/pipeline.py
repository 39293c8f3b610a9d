"""Caller side of the hot path: DDIM scheduler, sliding context windows and the denoise + decode loop of
`EMOAnimationPipeline.__call__` (reference EMOAnimationPipeline.py:543-840), on the sm_100a kernels.

What is mirrored (same names / argument meaning): `DDIMScheduler.{set_timesteps, timesteps, init_noise_sigma,
scale_model_input, step(...).prev_sample}` (third-party diffusers class the reference instantiates at
EMOAnimationPipeline.py:908 / magicanimate/pipelines/animation.py:104), `get_context_scheduler('uniform')`
(magicanimate/pipelines/context.py:20-42), `EMOAnimationPipeline.{prepare_latents, decode_latents, __call__}`.
The ReferenceNet writer (`AppearanceEncoderModel`) is optional: pass `appearance_encoder` +
`ref_image_latents` and it runs once per timestep like the reference (:711-716), its banks feeding the reader blocks
through persistent device buffers (no per-window clone / cast / cat); or pass pre-computed `reference_banks`.
What is deliberately NOT here (out of scope for the path): CLIP text encoding, wav2vec feature extraction and the
pose ControlNet — their outputs enter as tensors (`text_embeddings` / per-frame audio tokens as
`encoder_hidden_states`, `down_block_additional_residuals`).

Multi-GPU: the unit of independent work is (sample | context window) x CFG branch.  `__call__` shards
the windows of one timestep across ranks (`global_context[rank::world_size]`, EMOAnimationPipeline.py:757) and — only
when more than one rank contributes to a step — replaces the reference's gather + broadcast + 2 barriers
(:796-802, :819-821) with ONE all-reduce of the accumulated noise prediction; every rank then runs the fused
CFG + DDIM update redundantly.  Decoded frames are sharded by frame and all-gathered once (uint8).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from . import unet3d as _unet3d
from .unet3d import ReferenceAttentionControl, reference_blocks


# =============================================================================================== scheduler
@dataclass
class DDIMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class DDIMScheduler:
    """diffusers.DDIMScheduler subset used by the reference (eta = 0, epsilon prediction, 'leading' spacing).
    Defaults = configs/inference.yaml:23-26 + the steps_offset=1 / clip_sample=False the pipeline forces
    (EMOAnimationPipeline.py:105-130)."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "linear", clip_sample: bool = False, set_alpha_to_one: bool = True,
                 steps_offset: int = 1, prediction_type: str = "epsilon"):
        if clip_sample or prediction_type != "epsilon":
            raise NotImplementedError("DDIMScheduler: only clip_sample=False / epsilon prediction (the reference's settings)")
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.config = type("Cfg", (), dict(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                                           clip_sample=clip_sample, beta_start=beta_start, beta_end=beta_end,
                                           beta_schedule=beta_schedule))()
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)  # kept on the host: the loop reads python ints (no device sync)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def alphas_for(self, timestep: int):
        prev = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[timestep])
        a_prev = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise: Optional[torch.Tensor] = None,
             return_dict: bool = True, **kwargs):
        """x_t -> x_{t-1} (`emote_ddim_step`).  eta > 0 adds the stochastic DDIM term sigma_t z with
        sigma_t = eta sqrt((1-a_prev)/(1-a_t) (1 - a_t/a_prev)); z = `variance_noise` or drawn from `generator`."""
        if use_clipped_model_output:
            raise NotImplementedError("DDIMScheduler.step: use_clipped_model_output (clip_sample is off in the reference)")
        a_t, a_prev = self.alphas_for(int(timestep))
        out = sample.float().contiguous().clone()
        eps = model_output.float().contiguous()
        sigma = ops.ddim_sigma(a_t, a_prev, float(eta))
        noise = None
        if sigma > 0.0:
            noise = variance_noise if variance_noise is not None else \
                torch.randn(out.shape, generator=generator, device=out.device, dtype=torch.float32)
            noise = noise.to(device=out.device, dtype=torch.float32).contiguous()
        ops.ddim_step(out, eps, a_t, a_prev, noise, sigma)
        out = out.to(sample.dtype)
        return DDIMSchedulerOutput(prev_sample=out) if return_dict else (out,)


# =============================================================================================== context windows
def _ordered_halving(val: int) -> float:
    return int(f"{val:064b}"[::-1], 2) / (1 << 64)


def uniform(step: int = ..., num_steps: Optional[int] = None, num_frames: int = ..., context_size: Optional[int] = None,
            context_stride: int = 3, context_overlap: int = 4, closed_loop: bool = True):
    """magicanimate/pipelines/context.py:20-42 — windows of `context_size` frames every size-overlap frames, wrapping
    modulo num_frames."""
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    context_stride = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    for context_step in 1 << np.arange(context_stride):
        pad = int(round(num_frames * _ordered_halving(step)))
        for j in range(int(_ordered_halving(step) * context_step) + pad,
                       num_frames + pad + (0 if closed_loop else -context_overlap),
                       (context_size * context_step - context_overlap)):
            yield [e % num_frames for e in range(j, j + context_size * context_step, context_step)]


def get_context_scheduler(name: str) -> Callable:
    if name == "uniform":
        return uniform
    raise ValueError(f"Unknown context_overlap policy {name}")


# =============================================================================================== audio tokens
def wav2vec_window_features(hidden_states: torch.Tensor, m: int = 2, n: int = 2, as_tokens: bool = True) -> torch.Tensor:
    """Neighbour-frame windows of wav2vec2 hidden states (Net.py:646-667) — see `audio.window_features`; the wav2vec2 forward
    that produces the hidden states is `audio.Wav2Vec2Model`."""
    from .audio import window_features
    if hidden_states.dim() not in (2, 3):
        raise ValueError("wav2vec_window_features: expected hidden states [T, d] and m, n >= 0")
    return window_features(hidden_states, m, n, as_tokens)


# =============================================================================================== CUDA graph
def _weights_fingerprint(module) -> int:
    """Changes whenever a parameter is rewritten in place (load_state_dict) or re-allocated (.to / .cuda): a captured
    graph reads the packed bf16 copies made from the parameters at capture time, so it must be rebuilt then."""
    return hash(tuple((p.data_ptr(), p._version) for p in module.parameters()))


class GraphedUNet:
    """One UNet3D step (all its kernel launches) captured in a CUDA graph and replayed per (timestep, unit): static input
    buffers (latents, timestep, context, banks), static output.  Removes the per-launch CPU cost (ctypes + tensor-map
    encoding + allocator) from the 50-step loop.

    mode: "pair" = both CFG branches of a window in one batch-2 call; "uncond" / "cond" = ONE branch (batch 1) — the unit
    of multi-GPU sharding.  The unconditional branch never sees the reference banks and the conditional
    branch reads bank row 1 (mutual_self_attention.py:237-255), so a lone branch needs no CFG masking."""

    replayed_kernels = 0  # kernels of libemote_b200 launched through graph replays (bench.py "gpu_launches")

    def __init__(self, unet, lat_shape, ctx: torch.Tensor, banks: Optional[Dict[str, List[torch.Tensor]]], dev,
                 alias_banks: bool = False, mode: str = "pair", writer=None):
        """alias_banks: read the given bank tensors in place (static outputs of a GraphedWriter) instead of private copies."""
        if mode not in ("pair", "uncond", "cond") or lat_shape[0] != (2 if mode == "pair" else 1):
            raise ValueError("GraphedUNet: mode must be pair (batch 2) / uncond / cond (batch 1)")
        self.unet, self.mode, self.writer = unet, mode, writer
        self.fingerprint = _weights_fingerprint(unet)
        self.lat = torch.zeros(lat_shape, dtype=torch.float32, device=dev)
        self.t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.ctx = torch.empty_like(ctx, dtype=torch.float32).copy_(ctx)
        self.bank_src = banks if (banks is not None and mode != "uncond") else None
        self.banks = None
        if self.bank_src is not None:
            sel = (lambda t: t) if mode == "pair" else (lambda t: t[1:2])
            self.banks = {k: [sel(t) if alias_banks else sel(t).clone() for t in v] for k, v in banks.items()}
        prev = _unet3d.CTX_KV_CACHE_ENABLED
        _unet3d.CTX_KV_CACHE_ENABLED = False
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._run()  # warm-up: packs weights, sets kernel attributes (not capturable work)
            torch.cuda.current_stream(dev).wait_stream(side)
            from . import _lib
            n0 = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._run()
            self.kernels_per_replay = _lib.launch_count() - n0  # kernel nodes of this library inside the graph
        finally:
            _unet3d.CTX_KV_CACHE_ENABLED = prev

    def _run(self):
        return _run_unet(self.unet, self.mode, self.lat, self.t, self.ctx, self.banks)

    def refresh(self, ctx: Optional[torch.Tensor], banks: Optional[Dict[str, List[torch.Tensor]]]):
        """new conditioning for a cached graph: contents are copied into the static buffers"""
        if ctx is not None and ctx is not self.ctx:
            self.ctx.copy_(ctx)
        if banks is not None and self.banks is not None:
            for k, v in banks.items():
                for dst, src in zip(self.banks[k], v):
                    src = src if self.mode == "pair" else src[1:2]
                    if dst.data_ptr() != src.data_ptr():
                        dst.copy_(src)

    def replay(self, t) -> torch.Tensor:
        """inputs were written into .lat / .ctx in place (ops.gather_frames); returns the static output"""
        ops.fill_f32(self.t, float(t))
        self.graph.replay()
        GraphedUNet.replayed_kernels += self.kernels_per_replay
        return self.out

    def __call__(self, lat: torch.Tensor, t, ctx: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.lat.copy_(lat)
        if ctx is not None and ctx is not self.ctx:
            self.ctx.copy_(ctx)
        return self.replay(t)


def _run_unet(unet, mode: str, lat, t, ctx, banks):
    """one UNet call of a unit: arms the reference-attention reader for the mode, runs, disarms"""
    reader = None
    if banks is not None and mode != "uncond":
        reader = ReferenceAttentionControl(unet, do_classifier_free_guidance=(mode == "pair"), mode="read",
                                           fusion_blocks="midup")
        reader.set_banks(banks)
    try:
        return unet(lat, t, encoder_hidden_states=ctx, return_dict=False)[0]
    finally:
        if reader is not None:
            reader.release()


class _EagerUNet:
    """same interface as GraphedUNet without capture (use_cuda_graph=False, debugging)"""

    def __init__(self, unet, lat_shape, ctx, dev, mode):
        self.unet, self.mode = unet, mode
        self.lat = torch.zeros(lat_shape, dtype=torch.float32, device=dev)
        self.ctx = torch.empty_like(ctx, dtype=torch.float32).copy_(ctx)
        self.banks = None

    def set_banks(self, banks):
        self.banks = None if (banks is None or self.mode == "uncond") else \
            {k: [t if self.mode == "pair" else t[1:2] for t in v] for k, v in banks.items()}

    def replay(self, t):
        return _run_unet(self.unet, self.mode, self.lat, float(t), self.ctx, self.banks)


def _writer_reader_pairs(unet, encoder):
    """(reader block name, writer block) pairs in the reference's order: both sides sorted by descending norm1 width
    (mutual_self_attention.py:585-588)."""
    names = {id(m): n for n, m in unet.named_modules()}
    return [(names[id(r)], w) for r, w in zip(reference_blocks(unet), reference_blocks(encoder))]


class GraphedWriter:
    """One ReferenceNet (AppearanceEncoderModel) pass captured in a CUDA graph: static reference latents / timestep /
    context in, the ten LayerNorm1 banks out at fixed addresses (`reader_banks`: UNet3D reader block name -> [bank]) that
    a GraphedUNet built with alias_banks=True reads in place — replaces the reference's per-window clone + cast + cat
    (`ReferenceAttentionControl.update`, mutual_self_attention.py:577-617)."""

    def __init__(self, encoder, unet, ref_latents: torch.Tensor, ctx: torch.Tensor, dev):
        self.encoder = encoder
        self.fingerprint = _weights_fingerprint(encoder)
        self.lat = ref_latents.float().contiguous().clone()
        self.t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.ctx = torch.empty_like(ctx, dtype=torch.float32).copy_(ctx)
        self.control = ReferenceAttentionControl(encoder, do_classifier_free_guidance=True, mode="write", fusion_blocks="midup")
        pairs = _writer_reader_pairs(unet, encoder)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                self._run()
            torch.cuda.current_stream(dev).wait_stream(side)
            from . import _lib
            n0 = _lib.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._run()
                self.reader_banks = {name: [w.bank[0]] for name, w in pairs}
            self.kernels_per_replay = _lib.launch_count() - n0
        finally:
            self.control.release()

    def _run(self):
        self.control.clear()
        self.encoder(self.lat, self.t, encoder_hidden_states=self.ctx, return_dict=False)

    def __call__(self, t):
        ops.fill_f32(self.t, float(t))
        self.graph.replay()
        GraphedUNet.replayed_kernels += self.kernels_per_replay
        return self.reader_banks


# =============================================================================================== work partition
def plan_units(n_windows: int, rank: int, world_size: int, shard: str = "units"):
    """The UNet calls one rank makes per timestep: [(window index, mode)], mode in {"pair", "uncond", "cond"}.

    shard="units" (default): the (window x CFG-branch) units, window-major with the unconditional branch
    first, are dealt to the ranks in equal contiguous blocks; two branches of one window that land on the same rank run as
    one batch-2 call.  240 frames = 20 windows = 40 units -> 5 per rank on 8 GPUs (ideal 8x); one 16-frame clip = 2 units
    (<= 2x).  shard="windows": the reference's split, whole windows round-robin — `global_context[rank::world_size]`
    (EMOAnimationPipeline.py:757), 3,3,3,3,2,2,2,2 windows at 240 frames (<= 6.67x)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} is not in [0, world_size = {world_size})")
    if shard == "windows":
        return [(w, "pair") for w in range(rank, n_windows, world_size)]
    if shard != "units":
        raise ValueError(f"unknown shard policy {shard!r}")
    n = 2 * n_windows
    lo, hi = rank * n // world_size, (rank + 1) * n // world_size
    calls, u = [], lo
    while u < hi:
        w, branch = divmod(u, 2)
        if branch == 0 and u + 1 < hi:
            calls.append((w, "pair"))
            u += 2
        else:
            calls.append((w, "cond" if branch else "uncond"))
            u += 1
    return calls


# =============================================================================================== pipeline
@dataclass
class AnimationPipelineOutput:
    videos: object


_CALL_KWARGS = {"prompt_embeddings", "reference_banks", "ref_image_latents", "appearance_context", "audio_features",
                "use_cuda_graph", "shard", "dist", "rank", "world_size"}


class EMOAnimationPipeline:
    """Denoise + decode loop of the reference pipeline (EMOAnimationPipeline.py:543-840) on the sm_100a kernels.

    Optional front-ends (all upstream of the hot path; each may be None): `text_encoder(prompts: List[str]) -> [n, 77, d]`
    embeddings (the reference's CLIP tokenizer + text encoder, :163-229), `audio_encoder` (`audio.Wav2VecFeatureExtractor`)
    and `speed_encoder` (`audio.SpeedEncoder`)."""

    MAX_GRAPHS = 12   # cached CUDA graphs (each pins its memory pool): least-recently-used ones are dropped

    def __init__(self, vae, unet, scheduler: DDIMScheduler, rank: int = 0, world_size: int = 1, process_group=None,
                 text_encoder: Optional[Callable] = None, audio_encoder=None, speed_encoder=None, appearance_encoder=None):
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self.text_encoder, self.audio_encoder, self.speed_encoder = text_encoder, audio_encoder, speed_encoder
        self.appearance_encoder = appearance_encoder
        self.vae_scale_factor = 8
        if world_size < 1 or not 0 <= rank < world_size:
            raise ValueError(f"rank {rank} is not in [0, world_size = {world_size})")
        self.rank, self.world_size, self.process_group = rank, world_size, process_group
        self._graphs: "Dict[tuple, object]" = {}
        self.last_speed_embeddings = None

    # -- graph cache -------------------------------------------------------------------------------------------------
    def _cache_get(self, key):
        g = self._graphs.pop(key, None)
        if g is not None:
            self._graphs[key] = g          # most recently used last
        return g

    def _cache_put(self, key, g):
        self._graphs[key] = g
        while len(self._graphs) > self.MAX_GRAPHS:
            self._graphs.pop(next(iter(self._graphs)))

    def _drop_graphs_of_writer(self, writer):
        for k in [k for k, g in self._graphs.items() if getattr(g, "writer", None) is writer]:
            del self._graphs[k]

    # -- EMOAnimationPipeline.py:341-368 ---------------------------------------------------------------------------
    def prepare_latents(self, batch_size, num_channels_latents, video_length, height, width, dtype, device, generator,
                        latents=None, clip_length=16):
        shape = (batch_size, num_channels_latents, clip_length, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=device, dtype=dtype)
            latents = latents.repeat(1, 1, max(1, video_length // clip_length), 1, 1)
        else:
            if latents.shape != shape:
                raise ValueError(f"Unexpected latents shape, got {latents.shape}, expected {shape}")
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    # -- EMOAnimationPipeline.py:402-414 ---------------------------------------------------------------------------
    @torch.no_grad()
    def images2latents(self, images, dtype=torch.float32):
        """RGB frames [f, h, w, 3] (uint8 numpy / tensor) -> scaled VAE latents [f, 4, h/8, w/8] on the VAE's device: the
        reference-image latents the ReferenceNet writer consumes.  All frames are encoded in one batch instead of the
        reference's per-frame loop."""
        images = torch.as_tensor(images)
        dev = self.vae.device
        x = (images.to(dev).float() / 127.5 - 1).permute(0, 3, 1, 2).contiguous()
        mean = self.vae.encode(x)["latent_dist"].mean
        return (mean * self.vae.config.scaling_factor).to(dtype)

    # -- EMOAnimationPipeline.py:479-512 ---------------------------------------------------------------------------
    def interpolate_latents(self, latents: torch.Tensor, interpolation_factor: int, device=None) -> torch.Tensor:
        """Insert `interpolation_factor - 1` blended frames between neighbouring latent frames with the method chosen by
        `magicanimate.utils.util.set_tensor_interpolation_method` (linear / slerp).  The reference fixes the factor to 1
        (:825), a no-op; larger factors are a cold path kept for API parity (a handful of tiny element-wise ops)."""
        if interpolation_factor < 2:
            return latents
        from .magicanimate.utils.util import get_tensor_interpolation_method, linear
        blend = get_tensor_interpolation_method() or linear
        k, f = int(interpolation_factor), latents.shape[2]
        out = latents.new_zeros((latents.shape[0], latents.shape[1], (f - 1) * k + 1) + tuple(latents.shape[3:]))
        for j in range(f - 1):
            v0, v1 = latents[:, :, j], latents[:, :, j + 1]
            out[:, :, j * k] = v0
            for r in range(1, k):
                out[:, :, j * k + r] = blend(v0, v1, r / k)
        out[:, :, (f - 1) * k] = latents[:, :, f - 1]
        return out

    # -- EMOAnimationPipeline.py:379-400 ---------------------------------------------------------------------------
    @torch.no_grad()
    def next_step(self, model_output: torch.Tensor, timestep: int, x: torch.Tensor, eta: float = 0.0, verbose: bool = False):
        """Inverse DDIM update x_{t - ratio} -> x_t used by `invert`; returns (x_next, pred_x0).  Same algebra as
        `scheduler.step` with the two alphas exchanged, so it runs in the same kernel (`emote_ddim_step`).  Like the
        reference function, `eta` is accepted and unused (the inversion is deterministic)."""
        sch = self.scheduler
        nxt = int(timestep)
        cur = min(nxt - sch.config.num_train_timesteps // sch.num_inference_steps, 999)
        a_cur = float(sch.alphas_cumprod[cur]) if cur >= 0 else float(sch.final_alpha_cumprod)
        a_next = float(sch.alphas_cumprod[nxt])
        eps = model_output.float().contiguous()
        x_next = x.float().contiguous().clone()
        pred_x0 = x_next.clone()
        ops.ddim_step(x_next, eps, a_cur, a_next)
        ops.ddim_step(pred_x0, eps, a_cur, 1.0)   # alpha_prev = 1: x0 itself
        return x_next.to(x.dtype), pred_x0.to(x.dtype)

    def _embed_prompt(self, prompt, negative_prompt=None, do_cfg: bool = True, device=None):
        """EMOAnimationPipeline._encode_prompt (:163-229): [uncond | cond] text embeddings.  A tensor passes through; strings
        need the optional `text_encoder` front-end (CLIP is upstream of the hot path and not part of this package)."""
        if torch.is_tensor(prompt):
            return prompt
        if self.text_encoder is None:
            raise NotImplementedError(
                "string prompts need a text encoder: construct the pipeline with text_encoder=callable(List[str]) -> "
                "[n, 77, d] embeddings, or pass prompt_embeddings=[uncond | cond] (the CLIP front-end is out of scope)")
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        cond = self.text_encoder(prompts)
        if not do_cfg:
            return cond
        neg = [""] * len(prompts) if negative_prompt is None else \
            ([negative_prompt] * len(prompts) if isinstance(negative_prompt, str) else list(negative_prompt))
        return torch.cat([self.text_encoder(neg), cond])

    # -- EMOAnimationPipeline.py:416-477 ---------------------------------------------------------------------------
    @torch.no_grad()
    def invert(self, image, prompt, num_inference_steps: int = 20, num_actual_inference_steps: Optional[int] = 10,
               eta: float = 0.0, return_intermediates: bool = False, **kwargs):
        """Deterministic DDIM inversion of real frames into a noise map.  `image`: uint8 frames [f, h, w, 3] (encoded with
        `images2latents`) or latents [f, 4, h/8, w/8]; `prompt`: text EMBEDDINGS [1, n, d], or a string when the pipeline
        has a `text_encoder`."""
        if isinstance(prompt, (str, list)):
            if self.text_encoder is None:
                raise TypeError("invert: pass the prompt's text embeddings [1, n, d] (no text_encoder attached; the CLIP "
                                "text encoder the reference calls here is upstream of the path)")
            prompt = self._embed_prompt(prompt, do_cfg=False)
        image = torch.as_tensor(image)
        if image.is_floating_point() and image.dim() == 4 and image.shape[1] == self.unet.in_channels:
            latents = image.to(prompt.device).float()
        else:
            latents = self.images2latents(image)
        self.scheduler.set_timesteps(num_inference_steps)
        latents_list = [latents]
        for i, t in enumerate(reversed(self.scheduler.timesteps.tolist())):
            if num_actual_inference_steps is not None and i >= num_actual_inference_steps:
                continue
            model_inputs = latents.permute(1, 0, 2, 3)[None].contiguous()            # f c h w -> 1 c f h w
            noise_pred = self.unet(model_inputs, t, encoder_hidden_states=prompt).sample
            noise_pred = noise_pred[0].permute(1, 0, 2, 3).contiguous()                # 1 c f h w -> f c h w
            latents, _ = self.next_step(noise_pred, t, latents, eta)
            latents_list.append(latents)
        return (latents, latents_list) if return_intermediates else latents

    # -- EMOAnimationPipeline.py:291-307 ---------------------------------------------------------------------------
    def decode_latents(self, latents, rank=0, decoder_consistency=None):
        """-> numpy fp32 [b, 3, f, H, W] in [0, 1] (host copy, like the reference)."""
        if decoder_consistency is not None:
            raise NotImplementedError("decoder_consistency is not supported")
        video, _ = self.decode_latents_device(latents)
        return video.cpu().float().numpy()

    def decode_latents_device(self, latents, want_u8: bool = False, shard: bool = False, frame_chunk: Optional[int] = None):
        """Device-side decode (frames decoded `frame_chunk` at a time, default 16: a long-form clip never holds more than
        one chunk of decoder activations).  With shard=True each rank decodes a contiguous block of frames and the uint8
        frames are all-gathered once over NCCL (the single collective of the path)."""
        if not shard or self.world_size == 1:
            return self.vae.decode_video(latents, want_u8=want_u8, frame_chunk=frame_chunk)
        import torch.distributed as dist
        b, c, f, h, w = latents.shape
        per = math.ceil(f / self.world_size)
        lo, hi = min(f, self.rank * per), min(f, (self.rank + 1) * per)
        pad = torch.zeros((b, 3, per, 8 * h, 8 * w), dtype=torch.uint8, device=latents.device)
        if hi > lo:
            _, u8 = self.vae.decode_video(latents[:, :, lo:hi].contiguous(), want_u8=True, frame_chunk=frame_chunk)
            pad[:, :, : hi - lo] = u8
        gathered = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(gathered, pad, group=self.process_group)
        video_u8 = torch.cat(gathered, dim=2)[:, :, :f]
        return None, video_u8

    # -- EMOAnimationPipeline.py:698-823 ---------------------------------------------------------------------------
    @torch.no_grad()
    def denoise(self, latents: torch.Tensor, text_embeddings: torch.Tensor, num_inference_steps: int = 50,
                guidance_scale: float = 7.5, context_frames: int = 16, context_stride: int = 1, context_overlap: int = 4,
                context_schedule: str = "uniform", reference_banks: Optional[Dict[str, List[torch.Tensor]]] = None,
                callback: Optional[Callable] = None, use_cuda_graph: bool = True, appearance_encoder=None,
                ref_image_latents: Optional[torch.Tensor] = None,
                appearance_context: Optional[torch.Tensor] = None, eta: float = 0.0, generator=None,
                num_actual_inference_steps: Optional[int] = None, callback_steps: int = 1,
                shard: str = "units") -> torch.Tensor:
        """latents [1, 4, F_total, h, w] fp32 (updated in place and returned); text_embeddings = cat([uncond, cond])
        of shape [2, n, d], or per-frame [2*F_total, n, d] audio tokens (uncond frames first).
        appearance_encoder + ref_image_latents [1, 4, h, w]: run the ReferenceNet writer once per timestep on the
        reference-image latents repeated over the CFG pair (EMOAnimationPipeline.py:711-716) and feed its banks to the
        reader blocks; its context is `appearance_context` [2, n, d] (default: text_embeddings when that is a CFG pair).
        Multi-GPU (world_size > 1): the (window x CFG-branch) units of a timestep are dealt to the ranks (`plan_units`),
        every rank accumulates its share into `noise_pred`, ONE all-reduce per step replaces the reference's gather +
        broadcast + barriers (:796-821) and every rank applies the fused CFG + DDIM update redundantly."""
        writer_ctx = None
        if appearance_encoder is not None:
            if reference_banks is not None:
                raise ValueError("pass either reference_banks or appearance_encoder, not both")
            if ref_image_latents is None or ref_image_latents.dim() != 4 or ref_image_latents.shape[0] != 1:
                raise ValueError("appearance_encoder needs ref_image_latents of shape [1, 4, h, w]")
            writer_ctx = text_embeddings if appearance_context is None else appearance_context
            if writer_ctx.shape[0] != 2:
                raise ValueError("the ReferenceNet writer needs a [2, n, d] context (pass appearance_context)")
        if not guidance_scale > 1.0:
            # the reference's own no-CFG path cannot run either: `pred_uc, pred_c = pred.chunk(2)` on a batch-1 prediction
            # and `control.chunk(2)` (EMOAnimationPipeline.py:651,787) fail to unpack, and its average over overlapping
            # windows (`noise_pred / counter`, :813) only happens inside the CFG branch
            raise NotImplementedError("guidance_scale <= 1: the fused sampler implements the classifier-free-guidance path "
                                      "(the only one the reference pipeline can execute)")
        if latents.dim() != 5 or latents.shape[0] != 1:
            raise ValueError("denoise() handles one sample [1, c, f, h, w] per call (run samples on different ranks / sequentially)")
        dev = latents.device
        latents = latents.float().contiguous()
        _, cl, f_total, h, w = latents.shape
        inner = h * w
        sch = self.scheduler
        sch.set_timesteps(num_inference_steps, device=dev)
        windows = [list(map(int, c)) for c in get_context_scheduler(context_schedule)(
            0, num_inference_steps, f_total, context_frames, context_stride, context_overlap)]
        counter = torch.zeros(f_total, dtype=torch.float32)
        for c in windows:
            counter[c] += 1
        counter = counter.to(dev)
        calls = plan_units(len(windows), self.rank, self.world_size, shard)
        need_reduce = self.world_size > 1      # ranks without a unit still contribute zeros and receive the sum
        ctx_all = text_embeddings.to(device=dev, dtype=torch.float32).contiguous()
        per_frame_ctx = ctx_all.shape[0] == 2 * f_total and f_total > 1
        if not per_frame_ctx and ctx_all.shape[0] != 2:
            raise ValueError("text_embeddings must be [2, n, d] (uncond | cond) or per-frame [2*F, n, d]")
        ctx_inner = ctx_all.shape[1] * ctx_all.shape[2]
        idx_dev = {wi: torch.tensor(windows[wi], dtype=torch.int32, device=dev) for wi in {wi for wi, _ in calls}}
        # direct: the one window is the whole clip in order and both branches are here — the UNet output IS the accumulator
        direct = (not need_reduce and calls == [(0, "pair")] and len(windows) == 1 and windows[0] == list(range(f_total)))
        noise_pred = None if direct else torch.zeros((2, cl, f_total, h, w), dtype=torch.float32, device=dev)

        # ---- ReferenceNet writer (once per timestep)
        gwriter, writer, ref_lat2 = None, None, None
        if appearance_encoder is not None and calls:
            ref_lat2 = ref_image_latents.to(dev).float().repeat(2, 1, 1, 1).contiguous()
            if use_cuda_graph:
                wkey = ("writer", id(appearance_encoder), tuple(ref_lat2.shape), tuple(writer_ctx.shape))
                gwriter = self._cache_get(wkey)
                if gwriter is not None and (gwriter.encoder is not appearance_encoder or
                                            gwriter.fingerprint != _weights_fingerprint(appearance_encoder)):
                    self._drop_graphs_of_writer(gwriter)          # UNet graphs alias the old writer's bank buffers
                    self._graphs.pop(wkey, None)
                    gwriter = None
                if gwriter is None:
                    gwriter = GraphedWriter(appearance_encoder, self.unet, ref_lat2, writer_ctx, dev)
                    self._cache_put(wkey, gwriter)
                else:
                    gwriter.lat.copy_(ref_lat2)
                    gwriter.ctx.copy_(writer_ctx)
                reference_banks = gwriter.reader_banks
            else:
                writer = ReferenceAttentionControl(appearance_encoder, do_classifier_free_guidance=True, mode="write",
                                                   fusion_blocks="midup")
                pairs = _writer_reader_pairs(self.unet, appearance_encoder)

        # ---- one runner per (mode, window length) this rank needs: static input buffers + captured step
        def mode_ctx(mode, wlen):
            if per_frame_ctx:
                return ctx_all.new_zeros(((2 if mode == "pair" else 1) * wlen, ctx_all.shape[1], ctx_all.shape[2]))
            return ctx_all if mode == "pair" else (ctx_all[0:1] if mode == "uncond" else ctx_all[1:2])

        runners = {}
        for wi, mode in calls:
            wlen = len(windows[wi])
            if (mode, wlen) in runners:
                continue
            nb = 2 if mode == "pair" else 1
            shape = (nb, cl, wlen, h, w)
            mctx = mode_ctx(mode, wlen)
            if use_cuda_graph:
                bsig = None if reference_banks is None else tuple(sorted((k, tuple(v[0].shape)) for k, v in reference_banks.items()))
                key = ("unet", mode, shape, tuple(mctx.shape), bsig)
                g = self._cache_get(key)
                if g is not None and (g.fingerprint != _weights_fingerprint(self.unet) or g.unet is not self.unet or
                                      g.writer is not gwriter):
                    self._graphs.pop(key, None)
                    g = None                                      # weights / writer changed since capture: rebuild
                if g is None:
                    g = GraphedUNet(self.unet, shape, mctx, reference_banks, dev, alias_banks=gwriter is not None, mode=mode,
                                    writer=gwriter)
                    self._cache_put(key, g)
                else:
                    g.refresh(None if per_frame_ctx else mctx, None if gwriter is not None else reference_banks)
                runners[(mode, wlen)] = g
            else:
                e = _EagerUNet(self.unet, shape, mctx, dev, mode)
                e.set_banks(reference_banks)
                runners[(mode, wlen)] = e

        timesteps = sch.timesteps.tolist()
        skip = 0 if num_actual_inference_steps is None else max(0, num_inference_steps - int(num_actual_inference_steps))
        try:
            for i, t in enumerate(timesteps):
                if i < skip:                                      # img2img setting (EMOAnimationPipeline.py:699-700)
                    continue
                if gwriter is not None:
                    gwriter(t)                                    # banks of this timestep land in the UNet graphs' inputs
                elif writer is not None:
                    writer.clear()
                    appearance_encoder(ref_lat2, t, encoder_hidden_states=writer_ctx, return_dict=False)
                    step_banks = {name: [wblk.bank[0]] for name, wblk in pairs}   # EMOAnimationPipeline.py:774
                    for r in runners.values():
                        r.set_banks(step_banks)
                for wi, mode in calls:
                    r = runners[(mode, len(windows[wi]))]
                    nb = 2 if mode == "pair" else 1
                    b0 = 1 if mode == "cond" else 0
                    # latents[:, :, c].repeat(nb) (:759-763) and the per-frame audio tokens of the window, written
                    # straight into the step's static inputs
                    ops.gather_frames(latents, r.lat, idx_dev[wi], nb * cl, f_total, inner, src_mod=cl)
                    if per_frame_ctx:
                        ops.gather_frames(ctx_all, r.ctx, idx_dev[wi], nb, f_total, ctx_inner, src_mod=2, src_off=b0)
                    pred = r.replay(t)
                    if direct:
                        noise_pred = pred   # static graph output: consumed by the fused update below before the next replay
                    else:                   # noise_pred[b0:b0+nb, :, c] += pred (:790-794)
                        ops.scatter_add_frames(pred, noise_pred, idx_dev[wi], nb * cl, f_total, inner, dst_off=b0 * cl)
                if need_reduce:
                    import torch.distributed as dist
                    dist.all_reduce(noise_pred, group=self.process_group)
                a_t, a_prev = sch.alphas_for(int(t))
                sigma = ops.ddim_sigma(a_t, a_prev, float(eta))
                z = None
                if sigma > 0.0:
                    z = torch.randn(latents.shape, generator=generator, device=dev, dtype=torch.float32)
                ops.cfg_ddim_step(latents, noise_pred, counter, guidance_scale, a_t, a_prev, zero_noise_pred=not direct,
                                  noise=z, sigma=sigma)
                if callback is not None and i % max(1, callback_steps) == 0:
                    callback(i, t, latents)
        finally:
            if writer is not None:
                writer.release()                                  # EMOAnimationPipeline.py:823
        return latents

    @torch.no_grad()
    def __call__(self, prompt, video_length: Optional[int], height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, negative_prompt=None,
                 num_videos_per_prompt: Optional[int] = 1, eta: float = 0.0, generator=None,
                 latents: Optional[torch.Tensor] = None, output_type: Optional[str] = "tensor", return_dict: bool = True,
                 callback: Optional[Callable] = None, callback_steps: Optional[int] = 1, controlnet_condition=None,
                 controlnet_conditioning_scale: float = 1.0, context_frames: int = 16, context_stride: int = 1,
                 context_overlap: int = 4, context_batch_size: int = 1, context_schedule: str = "uniform",
                 init_latents: Optional[torch.Tensor] = None, num_actual_inference_steps: Optional[int] = None,
                 appearance_encoder=None, reference_control_writer=None, reference_control_reader=None,
                 source_image=None, decoder_consistency=None, audio=None, head_rotation_speeds=None, **kwargs):
        """The reference's call surface (EMOAnimationPipeline.py:543-578), same argument names and meaning.

        `prompt`: a string / list (needs the pipeline's `text_encoder`) or the [uncond | cond] embeddings tensor; the keyword
        `prompt_embeddings=` does the same.  `audio`: a waveform tensor / wav path for the pipeline's `audio_encoder`
        (wav2vec2 -> per-frame [T, 5, 768] tokens used as `encoder_hidden_states`, Net.py:614-667), or pass ready tokens as
        `audio_features=` [T, n, d].  `source_image`: uint8 RGB array [H, W, 3] (or an image path) -> ReferenceNet latents
        through `images2latents`; `ref_image_latents=` passes them directly.  `reference_control_writer` / `_reader` are
        accepted and ignored exactly like the reference, which overwrites both (:633-634).
        Not on this path (raise): `controlnet_condition` (pose ControlNet, out of scope), `decoder_consistency`,
        context_batch_size != 1 and num_videos_per_prompt != 1 (the reference asserts both, :641-642).
        Extra keywords: reference_banks, appearance_context, use_cuda_graph, shard, and the reference's dist / rank /
        world_size (:636-638; the pipeline's own rank / world_size are used)."""
        unknown = set(kwargs) - _CALL_KWARGS
        if unknown:
            raise TypeError(f"EMOAnimationPipeline.__call__: unexpected keyword arguments {sorted(unknown)}")
        if controlnet_condition is not None:
            raise NotImplementedError("controlnet_condition: the pose ControlNet is not part of this path (EMO has none)")
        if context_batch_size != 1 or num_videos_per_prompt != 1:
            raise NotImplementedError("context_batch_size and num_videos_per_prompt must be 1 (as the reference asserts)")
        if video_length is None:
            raise ValueError("video_length is required")
        emb = kwargs.get("prompt_embeddings")
        if emb is None:
            emb = self._embed_prompt(prompt, negative_prompt, do_cfg=True)
        dev = emb.device
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        # audio cross-attention context: per-frame wav2vec tokens replace the text embeddings as encoder_hidden_states
        tokens = kwargs.get("audio_features")
        if tokens is None and audio is not None:
            if self.audio_encoder is None:
                raise NotImplementedError("audio= needs the pipeline's audio_encoder (audio.Wav2VecFeatureExtractor), or pass "
                                          "ready per-frame tokens as audio_features=[T, n, d]")
            tokens = self.audio_encoder.extract_tokens(audio, m=2, n=2)
        appearance_context = kwargs.get("appearance_context")
        ctx = emb
        if tokens is not None:
            tokens = tokens.to(dev).float()
            if tokens.shape[0] < video_length:
                raise ValueError(f"audio features cover {tokens.shape[0]} frames, video_length is {video_length}")
            tokens = tokens[:video_length]
            ctx = torch.cat([torch.zeros_like(tokens), tokens])   # unconditional branch: silence (zero tokens)
            if appearance_context is None:
                appearance_context = emb
        if head_rotation_speeds is not None:
            # the reference computes these and hands them to a UNet signature that has no such argument
            # (EMOAnimationPipeline.py:597-601,784 vs unet_controlnet.py:328-339): computed and exposed, not consumed
            if self.speed_encoder is None:
                raise NotImplementedError("head_rotation_speeds needs the pipeline's speed_encoder (audio.SpeedEncoder)")
            self.last_speed_embeddings = self.speed_encoder(head_rotation_speeds)
        # latents
        if init_latents is not None:                              # "(b f) c h w -> b c f h w" (:656-657)
            lat = init_latents.to(dev).float().reshape(-1, video_length, *init_latents.shape[1:]).permute(0, 2, 1, 3, 4)
        else:
            lat = self.prepare_latents(1, self.unet.in_channels, video_length, height, width, torch.float32, dev, generator,
                                       latents, clip_length=video_length if latents is not None else min(context_frames, video_length))
        lat = lat[:, :, :video_length].contiguous()
        # ReferenceNet
        appearance_encoder = appearance_encoder if appearance_encoder is not None else self.appearance_encoder
        ref_lat = kwargs.get("ref_image_latents")
        if ref_lat is None and source_image is not None and appearance_encoder is not None:
            if isinstance(source_image, str):
                from PIL import Image   # optional dependency, only for the path form of the argument
                source_image = np.array(Image.open(source_image).convert("RGB").resize((width, height)))
            ref_lat = self.images2latents(np.asarray(source_image)[None], torch.float32)
        banks = kwargs.get("reference_banks")
        if appearance_encoder is not None and ref_lat is None and banks is None:
            raise ValueError("appearance_encoder needs source_image= (or ref_image_latents=)")
        lat = self.denoise(lat, ctx, num_inference_steps, guidance_scale, context_frames, context_stride, context_overlap,
                           context_schedule, banks, callback, use_cuda_graph=kwargs.get("use_cuda_graph", True),
                           appearance_encoder=appearance_encoder if banks is None else None, ref_image_latents=ref_lat,
                           appearance_context=appearance_context, eta=eta, generator=generator,
                           num_actual_inference_steps=num_actual_inference_steps, callback_steps=callback_steps or 1,
                           shard=kwargs.get("shard", "units"))
        video = self.decode_latents(lat, self.rank, decoder_consistency=decoder_consistency)
        if output_type == "tensor":
            video = torch.from_numpy(video)
        return AnimationPipelineOutput(videos=video) if return_dict else video
