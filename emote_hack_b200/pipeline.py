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

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0, **kwargs):
        """x_t -> x_{t-1}; the arithmetic runs in emote_cfg_ddim_step (guidance folded out: eps is used as is)."""
        if eta != 0.0:
            raise NotImplementedError("DDIMScheduler.step: eta != 0 (stochastic DDIM) is not used by the reference")
        a_t, a_prev = self.alphas_for(int(timestep))
        out = sample.float().contiguous().clone()
        eps = model_output.float().contiguous().reshape(1, 1, 1, 1, -1)
        # guidance 1.0 with identical halves == plain epsilon; the update is elementwise, so any sample rank works
        ops.cfg_ddim_step(out.view(1, 1, 1, 1, -1), torch.cat([eps, eps]), None, 1.0, a_t, a_prev)
        return DDIMSchedulerOutput(prev_sample=out.to(sample.dtype))


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
    """Neighbour-frame windows of wav2vec2 hidden states (Net.py:646-667, `extract_features_from_wav`): frame f gets the
    features of frames f-m .. f+n, zero-padded past either end.  hidden_states [T, d] (or [1, T, d]) ->
    [T, m+n+1, d] tokens (as_tokens, the per-frame `encoder_hidden_states` of the audio cross-attention) or the
    reference's flattened [T, (m+n+1)*d].  A gather only — the wav2vec2 forward itself is upstream of the path."""
    h = hidden_states[0] if hidden_states.dim() == 3 else hidden_states
    if h.dim() != 2 or m < 0 or n < 0:
        raise ValueError("wav2vec_window_features: expected hidden states [T, d] and m, n >= 0")
    t, d = h.shape
    padded = torch.cat([h.new_zeros(m, d), h, h.new_zeros(n, d)])
    idx = torch.arange(t, device=h.device)[:, None] + torch.arange(m + n + 1, device=h.device)[None]
    win = padded[idx]                                            # [T, m+n+1, d]
    return win if as_tokens else win.reshape(t, (m + n + 1) * d)


# =============================================================================================== CUDA graph
def _weights_fingerprint(module) -> int:
    """Changes whenever a parameter is rewritten in place (load_state_dict) or re-allocated (.to / .cuda): a captured
    graph reads the packed bf16 copies made from the parameters at capture time, so it must be rebuilt then."""
    return hash(tuple((p.data_ptr(), p._version) for p in module.parameters()))


class GraphedUNet:
    """One UNet3D step (all ~650 kernel launches) captured in a CUDA graph and replayed per (timestep, window):
    static input buffers (latents, timestep, context, banks), static output.  Removes the per-launch CPU cost
    (ctypes + tensor-map encoding + allocator) from the 50-step loop."""

    replayed_kernels = 0  # kernels of libemote_b200 launched through graph replays (bench.py "gpu_launches")

    def __init__(self, unet, lat_shape, ctx: torch.Tensor, banks: Optional[Dict[str, List[torch.Tensor]]], dev,
                 alias_banks: bool = False):
        """alias_banks: read the given bank tensors in place (static outputs of a GraphedWriter) instead of private copies."""
        self.unet = unet
        self.fingerprint = _weights_fingerprint(unet)
        self.lat = torch.zeros(lat_shape, dtype=torch.float32, device=dev)
        self.t = torch.zeros(1, dtype=torch.float32, device=dev)
        self.ctx = torch.empty_like(ctx, dtype=torch.float32).copy_(ctx)
        self.banks = None if banks is None else {k: [t if alias_banks else t.clone() for t in v] for k, v in banks.items()}
        self.reader = None
        if self.banks is not None:
            self.reader = ReferenceAttentionControl(unet, do_classifier_free_guidance=True, mode="read", fusion_blocks="midup")
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
            if self.reader is not None:
                self.reader.clear()
                for blk in self.reader._blocks(unet):
                    blk._ref_mode = None

    def _run(self):
        if self.reader is not None:
            self.reader.set_banks(self.banks)
        return self.unet(self.lat, self.t, encoder_hidden_states=self.ctx, return_dict=False)[0]

    def __call__(self, lat: torch.Tensor, t: int, ctx: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.lat.copy_(lat)
        self.t.fill_(float(t))
        if ctx is not None and ctx is not self.ctx:
            self.ctx.copy_(ctx)
        self.graph.replay()
        GraphedUNet.replayed_kernels += self.kernels_per_replay
        return self.out


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
            self.control.clear()
            for blk in self.control._blocks(encoder):
                blk._ref_mode = None

    def _run(self):
        self.control.clear()
        self.encoder(self.lat, self.t, encoder_hidden_states=self.ctx, return_dict=False)

    def __call__(self, t: int):
        self.t.fill_(float(t))
        self.graph.replay()
        GraphedUNet.replayed_kernels += self.kernels_per_replay
        return self.reader_banks


# =============================================================================================== pipeline
@dataclass
class AnimationPipelineOutput:
    videos: object


class EMOAnimationPipeline:
    """Denoise + decode loop of the reference pipeline for pre-computed conditioning."""

    def __init__(self, vae, unet, scheduler: DDIMScheduler, rank: int = 0, world_size: int = 1,
                 process_group=None):
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self.vae_scale_factor = 8
        self.rank, self.world_size, self.process_group = rank, world_size, process_group
        self._graphs: Dict[tuple, GraphedUNet] = {}

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

    # -- EMOAnimationPipeline.py:379-400 ---------------------------------------------------------------------------
    @torch.no_grad()
    def next_step(self, model_output: torch.Tensor, timestep: int, x: torch.Tensor, eta: float = 0.0, verbose: bool = False):
        """Inverse DDIM update x_{t - ratio} -> x_t used by `invert`; returns (x_next, pred_x0).  Same algebra as
        `scheduler.step` with the two alphas swapped, so it runs in the same fused kernel (`emote_cfg_ddim_step`)."""
        if eta != 0.0:
            raise NotImplementedError("next_step: eta != 0 is not used by the reference")
        sch = self.scheduler
        nxt = int(timestep)
        cur = min(nxt - sch.config.num_train_timesteps // sch.num_inference_steps, 999)
        a_cur = float(sch.alphas_cumprod[cur]) if cur >= 0 else float(sch.final_alpha_cumprod)
        a_next = float(sch.alphas_cumprod[nxt])
        eps = model_output.float().contiguous().reshape(1, 1, 1, 1, -1)
        pair = torch.cat([eps, eps])                         # guidance 1.0 on identical halves == plain epsilon
        x_next = x.float().contiguous().clone()
        pred_x0 = x_next.clone()
        ops.cfg_ddim_step(x_next.view(1, 1, 1, 1, -1), pair, None, 1.0, a_cur, a_next)
        ops.cfg_ddim_step(pred_x0.view(1, 1, 1, 1, -1), pair, None, 1.0, a_cur, 1.0)   # alpha_prev = 1: x0 itself
        return x_next.to(x.dtype), pred_x0.to(x.dtype)

    # -- EMOAnimationPipeline.py:416-477 ---------------------------------------------------------------------------
    @torch.no_grad()
    def invert(self, image, prompt, num_inference_steps: int = 20, num_actual_inference_steps: Optional[int] = 10,
               eta: float = 0.0, return_intermediates: bool = False, **kwargs):
        """Deterministic DDIM inversion of real frames into a noise map.  `image`: uint8 frames [f, h, w, 3] (encoded with
        `images2latents`) or latents [f, 4, h/8, w/8]; `prompt`: the text EMBEDDINGS [1, n, d] (the CLIP text encoder
        the reference calls here is upstream of the path)."""
        if isinstance(prompt, (str, list)):
            raise TypeError("invert: pass the prompt's text embeddings [1, n, d]; the CLIP text encoder is out of scope")
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

    def decode_latents_device(self, latents, want_u8: bool = False, shard: bool = False):
        """Device-side decode.  With shard=True each rank decodes frames rank::world_size and the uint8 frames are
        all-gathered once over NCCL (the single collective of the path)."""
        if not shard or self.world_size == 1:
            return self.vae.decode_video(latents, want_u8=want_u8)
        import torch.distributed as dist
        b, c, f, h, w = latents.shape
        per = math.ceil(f / self.world_size)
        lo, hi = min(f, self.rank * per), min(f, (self.rank + 1) * per)
        pad = torch.zeros((b, 3, per, 8 * h, 8 * w), dtype=torch.uint8, device=latents.device)
        if hi > lo:
            _, u8 = self.vae.decode_video(latents[:, :, lo:hi].contiguous(), want_u8=True)
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
                appearance_context: Optional[torch.Tensor] = None) -> torch.Tensor:
        """latents [1, 4, F_total, h, w] fp32 (updated in place and returned); text_embeddings = cat([uncond, cond])
        of shape [2, n, d], or per-frame [2*F_total, n, d] audio tokens (uncond frames first).
        appearance_encoder + ref_image_latents [1, 4, h, w]: run the ReferenceNet writer once per timestep on the
        reference-image latents repeated over the CFG pair (EMOAnimationPipeline.py:711-716) and feed its banks to the
        reader blocks; its context is `appearance_context` [2, n, d] (default: text_embeddings when that is a CFG pair)."""
        writer_ctx = None
        if appearance_encoder is not None:
            if reference_banks is not None:
                raise ValueError("pass either reference_banks or appearance_encoder, not both")
            if ref_image_latents is None or ref_image_latents.dim() != 4 or ref_image_latents.shape[0] != 1:
                raise ValueError("appearance_encoder needs ref_image_latents of shape [1, 4, h, w]")
            writer_ctx = text_embeddings if appearance_context is None else appearance_context
            if writer_ctx.shape[0] != 2:
                raise ValueError("the ReferenceNet writer needs a [2, n, d] context (pass appearance_context)")
        do_cfg = guidance_scale > 1.0
        if not do_cfg:
            raise NotImplementedError("the fused sampler implements the classifier-free-guidance path the reference runs")
        if latents.shape[0] != 1:
            raise ValueError("denoise() handles one sample per call (run samples on different ranks / sequentially)")
        dev = latents.device
        latents = latents.float().contiguous()
        f_total = latents.shape[2]
        self.scheduler.set_timesteps(num_inference_steps, device=dev)
        windows = list(get_context_scheduler(context_schedule)(0, num_inference_steps, f_total, context_frames,
                                                               context_stride, context_overlap))
        counter = torch.zeros(f_total, dtype=torch.float32)
        for c in windows:
            counter[c] += 1
        counter = counter.to(dev)
        my_windows = windows[self.rank::self.world_size]
        need_reduce = self.world_size > 1 and len(windows) > 1
        per_frame_ctx = text_embeddings.shape[0] == 2 * f_total and f_total > 1
        reader = None
        single_window = len(windows) == 1 and windows[0] == list(range(f_total))
        noise_pred = torch.zeros((2,) + tuple(latents.shape[1:]), dtype=torch.float32, device=dev)
        graphed, gwriter, writer, ref_lat2 = None, None, None, None
        if appearance_encoder is not None:
            ref_lat2 = ref_image_latents.to(dev).float().repeat(2, 1, 1, 1).contiguous()
            if use_cuda_graph and not per_frame_ctx and len(my_windows) > 0:
                wkey = ("writer", id(appearance_encoder), tuple(ref_lat2.shape), tuple(writer_ctx.shape))
                gwriter = self._graphs.get(wkey)
                if gwriter is not None and gwriter.fingerprint != _weights_fingerprint(appearance_encoder):
                    gwriter = None                                # weights changed since capture: rebuild
                if gwriter is None:
                    gwriter = GraphedWriter(appearance_encoder, self.unet, ref_lat2, writer_ctx, dev)
                    self._graphs[wkey] = gwriter
                else:
                    gwriter.lat.copy_(ref_lat2)
                    gwriter.ctx.copy_(writer_ctx)
                reference_banks = gwriter.reader_banks
            else:
                writer = ReferenceAttentionControl(appearance_encoder, do_classifier_free_guidance=True, mode="write",
                                                   fusion_blocks="midup")
        if use_cuda_graph and not per_frame_ctx and len(my_windows) > 0:
            wlen = len(my_windows[0])
            key = (2, latents.shape[1], wlen, latents.shape[3], latents.shape[4], tuple(text_embeddings.shape),
                   None if reference_banks is None else tuple(sorted((k, tuple(v[0].shape)) for k, v in reference_banks.items())),
                   None if gwriter is None else id(gwriter))
            if all(len(w) == wlen for w in my_windows):
                graphed = self._graphs.get(key)
                if graphed is not None and graphed.fingerprint != _weights_fingerprint(self.unet):
                    graphed = None                                # weights changed since capture: rebuild
                if graphed is None:
                    graphed = GraphedUNet(self.unet, (2, latents.shape[1], wlen, latents.shape[3], latents.shape[4]),
                                          text_embeddings, reference_banks, dev, alias_banks=gwriter is not None)
                    self._graphs[key] = graphed
                else:
                    graphed.ctx.copy_(text_embeddings)
                    if reference_banks is not None:
                        for k, v in reference_banks.items():
                            for dst, src in zip(graphed.banks[k], v):
                                if dst is not src:
                                    dst.copy_(src)
        if graphed is None and (reference_banks is not None or writer is not None):  # eager path: banks are re-armed before every UNet call
            reader = ReferenceAttentionControl(self.unet, do_classifier_free_guidance=True, mode="read",
                                               fusion_blocks="midup")
        try:
            for i, t in enumerate(self.scheduler.timesteps.tolist()):
                if not single_window:
                    noise_pred.zero_()
                if gwriter is not None:
                    gwriter(t)                                   # banks of this timestep land in the UNet graph's inputs
                elif writer is not None:
                    writer.clear()
                    appearance_encoder(ref_lat2, t, encoder_hidden_states=writer_ctx, return_dict=False)
                for c in my_windows:
                    lat_in = latents if single_window else latents[:, :, c]
                    lat_in = lat_in.expand(2, -1, -1, -1, -1) if single_window else lat_in.repeat(2, 1, 1, 1, 1)
                    if per_frame_ctx:
                        idx = torch.as_tensor(c, device=dev)
                        ctx = torch.cat([text_embeddings[:f_total][idx], text_embeddings[f_total:][idx]])
                    else:
                        ctx = text_embeddings
                    if graphed is not None:
                        pred = graphed(lat_in, t)
                    else:
                        if reader is not None and writer is not None:
                            reader.update(writer)            # EMOAnimationPipeline.py:774
                        elif reader is not None:
                            reader.set_banks(reference_banks)
                        pred = self.unet(lat_in.contiguous(), t, encoder_hidden_states=ctx, return_dict=False)[0]
                    if single_window:
                        noise_pred = pred  # (a static graph output: consumed by the fused update below before the next replay)
                    else:
                        noise_pred[:, :, c] += pred
                if need_reduce:
                    import torch.distributed as dist
                    dist.all_reduce(noise_pred, group=self.process_group)
                a_t, a_prev = self.scheduler.alphas_for(int(t))
                ops.cfg_ddim_step(latents, noise_pred.contiguous(), counter, guidance_scale, a_t, a_prev)
                if callback is not None:
                    callback(i, t, latents)
        finally:
            if reader is not None:
                reader.clear()
                for blk in reader._blocks(self.unet):
                    blk._ref_mode = None
            if writer is not None:
                writer.clear()                                   # EMOAnimationPipeline.py:823
                for blk in writer._blocks(appearance_encoder):
                    blk._ref_mode = None
        return latents

    @torch.no_grad()
    def __call__(self, text_embeddings: torch.Tensor, video_length: int, height: int = 512, width: int = 512,
                 num_inference_steps: int = 50, guidance_scale: float = 7.5, generator=None, latents=None,
                 output_type: str = "tensor", return_dict: bool = True, context_frames: int = 16,
                 context_stride: int = 1, context_overlap: int = 4, context_schedule: str = "uniform",
                 reference_banks=None, callback=None, appearance_encoder=None, ref_image_latents=None,
                 appearance_context=None, **unused):
        dev = text_embeddings.device
        lat = self.prepare_latents(1, self.unet.in_channels, video_length, height, width, torch.float32, dev, generator,
                                   latents, clip_length=min(context_frames, video_length))
        lat = lat[:, :, :video_length].contiguous()
        lat = self.denoise(lat, text_embeddings, num_inference_steps, guidance_scale, context_frames, context_stride,
                           context_overlap, context_schedule, reference_banks, callback,
                           appearance_encoder=appearance_encoder, ref_image_latents=ref_image_latents,
                           appearance_context=appearance_context)
        video = self.decode_latents(lat, self.rank)
        if output_type == "tensor":
            video = torch.from_numpy(video)
        return AnimationPipelineOutput(videos=video) if return_dict else video
