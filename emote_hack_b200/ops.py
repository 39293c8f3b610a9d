"""Host-side operator layer: torch tensors in, C-ABI kernel launches out.

PyTorch is used here for device memory (allocation, views) and the current CUDA stream only; all arithmetic
on the hot path runs in libemote_b200.so.  Activations are "tokens-major": a [b, c, f, h, w] tensor stored
channels-last (torch.channels_last_3d), which is byte-identical to a row-major [(b f h w), c] matrix.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import EmoteAttnArgs, EmoteGemmArgs, check

F32, OP16 = torch.float32, _lib.op16_torch_dtype()   # OP16: fp16 (default) or bf16, see _lib.OPERAND
EPI_LINEAR, EPI_GEGLU, EPI_GELU = 0, 1, 2


def _stream() -> int:
    # one process drives one GPU (torch.cuda.set_device(local_rank), like the reference's launcher): kernels go to the
    # current stream of the current device, and _req() rejects tensors that live on another device
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.EmoteKernelError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.device.index != torch.cuda.current_device():
        raise _lib.EmoteKernelError(f"{name}: tensor on {t.device} but the current device is cuda:{torch.cuda.current_device()} "
                                    "(one process per GPU: call torch.cuda.set_device first)")
    if t.dtype != dtype:
        raise _lib.EmoteKernelError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.EmoteKernelError(f"{name}: expected a contiguous tensor")
    return t


def auto_block_n(N: int) -> int:
    return 160 if N % 160 == 0 else 128


# ----------------------------------------------------------------------------------------------- weight packing
def pack_linear(w: torch.Tensor) -> torch.Tensor:
    """nn.Linear / 1x1-conv weight [N, K(,1,1)] -> bf16 [N, K] (K-major B operand)."""
    return w.detach().reshape(w.shape[0], -1).to(OP16).contiguous()


def pack_conv3x3(w: torch.Tensor) -> torch.Tensor:
    """conv weight [N, C, 3, 3] -> bf16 [N, 9*C] with column (ky*3+kx)*C + c (tap-major, matches the TMA tap loop)."""
    n, c = w.shape[0], w.shape[1]
    return w.detach().permute(0, 2, 3, 1).reshape(n, 9 * c).to(OP16).contiguous()


def pack_conv3x3_small(w: torch.Tensor) -> torch.Tensor:
    """conv weight [N, Cl<=7, 3, 3] -> bf16 [N, 64] matching emote_latent_im2col columns, zero padded."""
    n, c = w.shape[0], w.shape[1]
    out = torch.zeros(n, 64, dtype=OP16, device=w.device)
    out[:, : 9 * c] = w.detach().permute(0, 2, 3, 1).reshape(n, 9 * c).to(OP16)
    return out


def pack_geglu(w: torch.Tensor, b: torch.Tensor):
    """GEGLU proj [2*inner, K]: rows [0,inner) = value, [inner,2*inner) = gate (orig_attention.py:825-827).
    Re-ordered per block_n tile as [value rows | gate rows] so one accumulator tile sees both halves."""
    n2, k = w.shape
    inner = n2 // 2
    bn = auto_block_n(n2)
    half = bn // 2
    if inner % half != 0:
        raise _lib.EmoteKernelError(f"GEGLU inner dim {inner} not divisible by {half}")
    wv, wg = w.detach()[:inner], w.detach()[inner:]
    wp = torch.stack([wv.reshape(-1, half, k), wg.reshape(-1, half, k)], dim=1).reshape(n2, k)
    bv, bg = b.detach()[:inner], b.detach()[inner:]
    bp = torch.stack([bv.reshape(-1, half), bg.reshape(-1, half)], dim=1).reshape(n2)
    return wp.to(OP16).contiguous(), bp.to(F32).contiguous()


# ----------------------------------------------------------------------------------------------- GEMM / conv
def gemm(a: torch.Tensor, w: torch.Tensor, *, bias=None, row_bias=None, rows_per_group: int = 0, residual=None,
         out_scale: float = 1.0, geglu: bool = False, out_dtype=F32, out: Optional[torch.Tensor] = None,
         conv: Optional[tuple] = None, M: Optional[int] = None, pair_mode: int = 0, tma_store: int = 0,
         stats_rows: int = 0, gelu: bool = False, lda: Optional[int] = None, out_col: int = 0) -> torch.Tensor:
    """out = epilogue(a @ w.T).  a: bf16 [M, K] (or NHWC [n_img, H, W, C] when conv=(n_img, H, W, C)); w: bf16 [N, K].

    stats_rows > 0 (fp32 outputs): the epilogue also accumulates per-column (sum, sum of squares) of the output per
    block of `stats_rows` rows; they ride on the returned tensor (`_emote_colstats`) and let group_norm() skip its
    statistics pass over that tensor.
    gelu: out = gelu_erf(a @ w.T + bias) (+ residual).  lda: row pitch of `a` in elements when it differs from K — may be
    smaller than K (overlapping rows: the zero-copy operand of a strided 1-D convolution; pass M).  out_col: first column
    of `out` (a wider [M, ld] tensor) the N outputs are written to."""
    _req(a, OP16, "gemm.a"), _req(w, OP16, "gemm.w")
    N, K = w.shape
    args = EmoteGemmArgs()
    if conv is not None:
        n_img, H, W_, Cc = conv
        M = n_img * H * W_
        args.conv_taps, args.n_img, args.H, args.W, args.C = 9, n_img, H, W_, Cc
        args.lda = Cc
    else:
        if M is None:
            M = a.numel() // K
        args.conv_taps = 1
        args.lda = K if lda is None else lda
        if (M - 1) * args.lda + K > a.numel():
            raise _lib.EmoteKernelError("gemm: the A operand (M, lda, K) reaches past the end of its buffer")
    args.M, args.N, args.K = M, N, K
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), dtype=out_dtype, device=a.device)
    else:
        _req(out, out_dtype, "gemm.out")
    if bias is not None:
        _req(bias, F32, "gemm.bias")
    if row_bias is not None:
        _req(row_bias, F32, "gemm.row_bias")
    if residual is not None:
        _req(residual, F32, "gemm.residual")
    args.bias = _ptr(bias)
    args.row_bias = _ptr(row_bias)
    args.rows_per_group = rows_per_group
    args.residual = _ptr(residual)
    args.ldr = n_out
    args.out_scale = out_scale
    args.epilogue = EPI_GEGLU if geglu else (EPI_GELU if gelu else EPI_LINEAR)
    if out_dtype not in (F32, OP16):
        raise _lib.EmoteKernelError(f"gemm: out_dtype must be float32 or the operand type {OP16}, got {out_dtype}")
    args.out_dtype = 1 if out_dtype == OP16 else 0
    args.ldc = out.shape[-1]
    args.block_n = FORCE_BLOCK_N
    args.pair_mode = pair_mode
    args.tma_store = tma_store
    colstats = None
    if (stats_rows > 0 and FUSED_GN_STATS and out_dtype == F32 and not geglu and not gelu and stats_rows % 32 == 0
            and M % stats_rows == 0 and n_out % 2 == 0 and _conv_stats_ok(conv, stats_rows)):
        # one fp32 (sum, sum of squares) slot per 32-row quarter and column: every slot is written, nothing to clear
        colstats = torch.empty((-(-M // 128) * 4 if conv is None else _conv_subtiles(conv) * 4, n_out, 2), dtype=F32,
                               device=a.device)
        args.colstats, args.stats_rows = colstats.data_ptr(), stats_rows
    if out_col:
        if out_col + n_out > out.shape[-1] or (out_col * out.element_size()) % 16 != 0:
            raise _lib.EmoteKernelError("gemm: out_col must keep the output window inside `out` and 16-byte aligned")
    check(_lib.load().emote_gemm_bf16(a.data_ptr(), w.data_ptr(), out.data_ptr() + out_col * out.element_size(),
                                      C.byref(args), _stream()), "emote_gemm_bf16")
    if colstats is not None:
        out._emote_colstats = (colstats, stats_rows, out._version)
    elif getattr(out, "_emote_colstats", None) is not None:
        out._emote_colstats = None   # overwritten in place: statistics of the old contents no longer apply
    return out


FORCE_BLOCK_N = 0       # dev knob (scripts/bench_gemm.py): 0 = library default, else 128 / 160
FUSED_GN_STATS = True   # GEMM epilogues that feed a GroupNorm also produce its statistics


def set_pdl(enabled: bool) -> None:
    """Programmatic dependent launch of the library's kernels (default off; also EMOTE_PDL=1)."""
    _lib.load().emote_set_pdl(1 if enabled else 0)


def set_tuning(key: str, value: int) -> None:
    """Measurement knobs of the library (`emote_set_tuning`): "gn_reduce" 1 = flat statistics fold (default), 0 = per-slot walk."""
    check(_lib.load().emote_set_tuning(key.encode(), int(value)), "emote_set_tuning")


def stats_rows_for(rows_per_frame: int, rows_per_sample: int) -> int:
    """Finest statistics granularity a 32-row quarter of a GEMM tile never straddles: per frame, else per sample, else
    none (0)."""
    if rows_per_frame % 32 == 0:
        return rows_per_frame
    if rows_per_sample % 32 == 0:
        return rows_per_sample
    return 0


def _conv_geometry(conv):
    """(bw, bh, images) of the 128-pixel TMA box of the implicit-GEMM conv and whether it is a 2-D patch (gemm_tcgen05.cu)"""
    n_img, H, W_, _ = conv
    bw = math.gcd(W_, 128)
    bh = math.gcd(H, 128 // bw)
    bn = 128 // (bw * bh)
    contiguous = bw == 128 or (bw == W_ and (bn == 1 or bh == H))
    return bw, bh, bn, not contiguous


def _conv_subtiles(conv) -> int:
    n_img, H, W_, _ = conv
    bw, bh, bn, patch = _conv_geometry(conv)
    if not patch:
        return -(-(n_img * H * W_) // 128)
    return -(-n_img // bn) * (W_ // bw) * (H // bh)


def _conv_stats_ok(conv, stats_rows: int) -> bool:
    """patch-mode convs: a sub-tile touches `images` whole-image groups, which must not straddle a statistics batch"""
    if conv is None:
        return True
    n_img, H, W_, _ = conv
    bw, bh, bn, patch = _conv_geometry(conv)
    return (not patch) or stats_rows % (H * W_ * bn) == 0


def conv_tile_ok(H: int, W: int) -> bool:
    """Whether the TMA implicit-GEMM conv covers an H x W map: always — a 128-pixel sub-tile is the box {gcd(W, 128) pixels
    of a row, rows, images}, a 2-D patch when rows do not pack into 128-pixel runs (96 / 48 / 24 / 12 latents)."""
    return H >= 1 and W >= 1


def conv3x3(x_bf16: torch.Tensor, w_packed: torch.Tensor, n_img: int, H: int, W: int, Cc: int, **epi) -> torch.Tensor:
    """3x3 / stride 1 / pad 1 conv over NHWC bf16.  Implicit GEMM when the tile geometry allows, else explicit im2col."""
    if conv_tile_ok(H, W) and Cc % 64 == 0:
        return gemm(x_bf16, w_packed, conv=(n_img, H, W, Cc), **epi)
    cols = torch.empty((n_img * H * W, 9 * Cc), dtype=OP16, device=x_bf16.device)
    check(_lib.load().emote_im2col3x3_bf16(x_bf16.data_ptr(), n_img, H, W, Cc, 1, cols.data_ptr(), _stream()),
          "emote_im2col3x3_bf16")
    return gemm(cols, w_packed, **epi)


def im2col_s2(x: torch.Tensor, n_img: int, H: int, W: int, Cc: int) -> torch.Tensor:
    _req(x, F32, "im2col_s2.x")
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    cols = torch.empty((n_img * Ho * Wo, 9 * Cc), dtype=OP16, device=x.device)
    check(_lib.load().emote_im2col3x3(x.data_ptr(), n_img, H, W, Cc, 2, cols.data_ptr(), _stream()), "emote_im2col3x3")
    return cols


def im2col_s2_pad01(x: torch.Tensor, n_img: int, H: int, W: int, Cc: int) -> torch.Tensor:
    """stride-2 3x3 operand with (0, 1) padding (VAE encoder downsample) -> bf16 [n_img*(H/2)*(W/2), 9*Cc]"""
    _req(x, F32, "im2col_s2_pad01.x")
    cols = torch.empty((n_img * (H // 2) * (W // 2), 9 * Cc), dtype=OP16, device=x.device)
    check(_lib.load().emote_im2col3x3_s2_pad01(x.data_ptr(), n_img, H, W, Cc, cols.data_ptr(), _stream()),
          "emote_im2col3x3_s2_pad01")
    return cols


def upsample2x(x: torch.Tensor, n_img: int, H: int, W: int, Cc: int) -> torch.Tensor:
    _req(x, F32, "upsample2x.x")
    out = torch.empty((n_img * 4 * H * W, Cc), dtype=OP16, device=x.device)
    check(_lib.load().emote_upsample2x(x.data_ptr(), n_img, H, W, Cc, out.data_ptr(), _stream()), "emote_upsample2x")
    return out


def upsample_nearest(x: torch.Tensor, n_img: int, H: int, W: int, Cc: int, Ho: int, Wo: int) -> torch.Tensor:
    """nearest-neighbour resize of tokens-major fp32 [n_img, H, W, C] to Ho x Wo -> op16 conv operand"""
    _req(x, F32, "upsample_nearest.x")
    out = torch.empty((n_img * Ho * Wo, Cc), dtype=OP16, device=x.device)
    check(_lib.load().emote_upsample_nearest(x.data_ptr(), n_img, H, W, Cc, Ho, Wo, out.data_ptr(), _stream()),
          "emote_upsample_nearest")
    return out


def latent_im2col(lat: torch.Tensor, pre_scale: float = 1.0, pw_weight=None, pw_bias=None) -> torch.Tensor:
    """lat [B, Cl, F, H, W] fp32 (standard NCFHW contiguous) -> bf16 [B*F*H*W, 64]."""
    _req(lat, F32, "latent_im2col.lat")
    B, Cl, F_, H, W = lat.shape
    out = torch.empty((B * F_ * H * W, 64), dtype=OP16, device=lat.device)
    check(_lib.load().emote_latent_im2col(lat.data_ptr(), B, Cl, F_, H, W, pre_scale, _ptr(pw_weight), _ptr(pw_bias),
                                          out.data_ptr(), _stream()), "emote_latent_im2col")
    return out


# ----------------------------------------------------------------------------------------------- norms
def group_norm(sources: Sequence[torch.Tensor], groups: int, rows_per_batch: int, n_batches: int, gamma, beta,
               eps: float, silu: bool, want_raw: bool = False):
    """GroupNorm over the channel-concatenation of `sources` (each fp32 [rows, C_i]); returns bf16 [rows, sum C_i]
    (and the raw bf16 concatenation when want_raw)."""
    lib = _lib.load()
    rows = rows_per_batch * n_batches
    c_total = sum(int(s.shape[-1]) for s in sources)
    dev = sources[0].device
    sums = torch.empty((n_batches, groups, 2), dtype=torch.float64, device=dev)
    out = torch.empty((rows, c_total), dtype=OP16, device=dev)
    raw = torch.empty((rows, c_total), dtype=OP16, device=dev) if want_raw else None
    st = _stream()
    off = 0
    for i, s in enumerate(sources):
        _req(s, F32, "group_norm.source")
        cs = int(s.shape[-1])
        info = getattr(s, "_emote_colstats", None)
        if (info is not None and FUSED_GN_STATS and info[2] == s._version and rows_per_batch % info[1] == 0
                and rows_per_batch % 32 == 0 and info[0].shape[0] * 32 >= rows and info[0].shape[1] == cs):
            # statistics were stored by the GEMM epilogue that wrote this source (32-row slots)
            check(lib.emote_gn_colstats_reduce(info[0].data_ptr(), cs, off, c_total, groups, rows_per_batch // 32,
                                               n_batches, sums.data_ptr(), 1 if i == 0 else 0, st),
                  "emote_gn_colstats_reduce")
        else:
            check(lib.emote_gn_stats(s.data_ptr(), cs, off, c_total, groups, rows_per_batch, n_batches, sums.data_ptr(),
                                     1 if i == 0 else 0, st), "emote_gn_stats")
        off += cs
    off = 0
    for s in sources:
        cs = int(s.shape[-1])
        check(lib.emote_gn_apply(s.data_ptr(), cs, off, c_total, groups, rows_per_batch, n_batches, sums.data_ptr(),
                                 gamma.data_ptr(), beta.data_ptr(), eps, 1 if silu else 0, out.data_ptr(), _ptr(raw),
                                 st), "emote_gn_apply")
        off += cs
    return out, raw


def layer_norm_dual(x: torch.Tensor, gamma, beta, eps: float = 1e-5, want_op16: bool = True):
    """LayerNorm of fp32 rows -> (op16 copy or None, fp32 copy): post-LN transformers continue their fp32 residual stream
    from the normalised value (wav2vec2 encoder layers)."""
    _req(x, F32, "layer_norm_dual.x")
    Cc = x.shape[-1]
    M = x.numel() // Cc
    o16 = torch.empty((M, Cc), dtype=OP16, device=x.device) if want_op16 else None
    o32 = torch.empty((M, Cc), dtype=F32, device=x.device)
    check(_lib.load().emote_layernorm_dual(x.data_ptr(), M, Cc, gamma.data_ptr(), beta.data_ptr(), eps, _ptr(o16),
                                           o32.data_ptr(), _stream()), "emote_layernorm_dual")
    return o16, o32


def layer_norm(x: torch.Tensor, gamma, beta, eps: float = 1e-5, pe: Optional[torch.Tensor] = None,
               rows_per_frame: int = 0, frames: int = 0) -> torch.Tensor:
    _req(x, F32, "layer_norm.x")
    Cc = x.shape[-1]
    M = x.numel() // Cc
    out = torch.empty((M, Cc), dtype=OP16, device=x.device)
    check(_lib.load().emote_layernorm(x.data_ptr(), M, Cc, gamma.data_ptr(), beta.data_ptr(), eps, _ptr(pe),
                                      rows_per_frame, frames, out.data_ptr(), _stream()), "emote_layernorm")
    return out


# ----------------------------------------------------------------------------------------------- attention
def attention(q: torch.Tensor, k0: torch.Tensor, v0: torch.Tensor, out: torch.Tensor, *, batch: int, heads: int,
              head_dim: int, nq: int, n0: int, q_strides, kv0_strides, o_strides, scale: float, kv0_batch_div: int = 1,
              k1=None, v1=None, n1: int = 0, kv1_strides=(0, 0), kv1_batch_div: int = 1, kv1_first_batch: int = 0,
              impl: str = "auto"):
    """Strided flash attention; q/k/v/out are bf16 views into (possibly fused) projection outputs.
    *_strides = (batch_stride, row_stride) in elements."""
    a = EmoteAttnArgs()
    a.q, a.k0, a.v0, a.out = q.data_ptr(), k0.data_ptr(), v0.data_ptr(), out.data_ptr()
    a.k1, a.v1 = _ptr(k1), _ptr(v1)
    a.batch, a.heads, a.head_dim = batch, heads, head_dim
    a.nq, a.n0, a.n1 = nq, n0, n1
    a.q_batch_stride, a.q_row_stride = q_strides
    a.kv0_batch_stride, a.kv0_row_stride = kv0_strides
    a.kv1_batch_stride, a.kv1_row_stride = kv1_strides
    a.o_batch_stride, a.o_row_stride = o_strides
    a.kv0_batch_div, a.kv1_batch_div, a.kv1_first_batch = kv0_batch_div, kv1_batch_div, kv1_first_batch
    a.scale = scale
    lib = _lib.load()
    # tcgen05 kernel for the long-sequence self-attention shapes; warp-level mma kernel for short key sets (text /
    # audio context) and the other head dims
    use_tc = impl == "tc" or (impl == "auto" and n0 >= 256 and lib.emote_attention_tc_supported(head_dim) == 1)
    if use_tc:
        check(lib.emote_attention_tc_bf16(C.byref(a), _stream()), "emote_attention_tc_bf16")
    else:
        check(lib.emote_attention_bf16(C.byref(a), _stream()), "emote_attention_bf16")
    return out


def attention_wide(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, batch: int, nq: int, nk: int,
                   q_strides, kv_strides, o_strides, scale: float):
    """single 512-wide head (VAE mid block): tcgen05 flash kernel batched over images; *_strides = (batch, row) elements"""
    check(_lib.load().emote_attention_wide_bf16(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), batch, nq, nk, 512,
                                                q_strides[0], q_strides[1], kv_strides[0], kv_strides[1], o_strides[0],
                                                o_strides[1], scale, _stream()), "emote_attention_wide_bf16")
    return out


def temporal_attention(qkv: torch.Tensor, B: int, F_: int, HW: int, heads: int, head_dim: int) -> torch.Tensor:
    _req(qkv, OP16, "temporal_attention.qkv")
    out = torch.empty((B * F_ * HW, heads * head_dim), dtype=OP16, device=qkv.device)
    check(_lib.load().emote_temporal_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, F_, HW, heads, head_dim,
                                                    head_dim ** -0.5, _stream()), "emote_temporal_attention_bf16")
    return out


def softmax_rows(scores: torch.Tensor, scale: float) -> torch.Tensor:
    _req(scores, F32, "softmax_rows.scores")
    N = scores.shape[-1]
    R = scores.numel() // N
    out = torch.empty(scores.shape, dtype=OP16, device=scores.device)
    check(_lib.load().emote_softmax_rows_bf16(scores.data_ptr(), R, N, scale, out.data_ptr(), _stream()),
          "emote_softmax_rows_bf16")
    return out


# ----------------------------------------------------------------------------------------------- misc
def cast_bf16(x: torch.Tensor, out: Optional[torch.Tensor] = None, c_offset: int = 0) -> torch.Tensor:
    _req(x, F32, "cast_bf16.x")
    cs = x.shape[-1]
    rows = x.numel() // cs
    if out is None:
        out = torch.empty((rows, cs), dtype=OP16, device=x.device)
    check(_lib.load().emote_cast_bf16(x.data_ptr(), rows, cs, c_offset, out.shape[-1], out.data_ptr(), _stream()),
          "emote_cast_bf16")
    return out


def silu_bf16(x: torch.Tensor) -> torch.Tensor:
    _req(x, F32, "silu_bf16.x")
    out = torch.empty(x.shape, dtype=OP16, device=x.device)
    check(_lib.load().emote_silu_bf16(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "emote_silu_bf16")
    return out


def tokens_to_ncfhw(tok: torch.Tensor, B: int, Cc: int, F_: int, H: int, W: int, ld: Optional[int] = None):
    _req(tok, F32, "tokens_to_ncfhw.tok")
    out = torch.empty((B, Cc, F_, H, W), dtype=F32, device=tok.device)
    check(_lib.load().emote_tokens_to_ncfhw(tok.data_ptr(), B, Cc, F_, H * W, out.data_ptr(), _stream()),
          "emote_tokens_to_ncfhw")
    return out


def ncfhw_to_tokens(x: torch.Tensor) -> torch.Tensor:
    _req(x, F32, "ncfhw_to_tokens.x")
    B, Cc, F_, H, W = x.shape
    out = torch.empty((B * F_ * H * W, Cc), dtype=F32, device=x.device)
    check(_lib.load().emote_ncfhw_to_tokens(x.data_ptr(), B, Cc, F_, H * W, out.data_ptr(), _stream()),
          "emote_ncfhw_to_tokens")
    return out


def add_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _req(a, F32, "add.a"), _req(b, F32, "add.b")
    out = torch.empty_like(a)
    check(_lib.load().emote_add_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "emote_add_f32")
    return out


def timestep_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool, freq_shift: float) -> torch.Tensor:
    _req(t, F32, "timestep_embedding.t")
    out = torch.empty((t.numel(), dim), dtype=OP16, device=t.device)
    check(_lib.load().emote_timestep_embedding(t.data_ptr(), t.numel(), dim, 1 if flip_sin_to_cos else 0,
                                               float(freq_shift), out.data_ptr(), _stream()), "emote_timestep_embedding")
    return out


def gather_frames(src: torch.Tensor, dst: torch.Tensor, frame_idx: torch.Tensor, n_outer: int, f_src: int, inner: int,
                  src_mod: int, src_off: int = 0) -> torch.Tensor:
    """dst[o, j, :] = src[(o % src_mod) + src_off, frame_idx[j], :] over [outer, frames, inner] views (fp32)."""
    _req(src, F32, "gather_frames.src"), _req(dst, F32, "gather_frames.dst")
    _req(frame_idx, torch.int32, "gather_frames.frame_idx")
    wlen = frame_idx.numel()
    if dst.numel() != n_outer * wlen * inner or (min(n_outer, src_mod) + src_off) * f_src * inner > src.numel():
        raise _lib.EmoteKernelError("gather_frames: buffer sizes do not match the [outer, frames, inner] geometry")
    check(_lib.load().emote_gather_frames(src.data_ptr(), dst.data_ptr(), frame_idx.data_ptr(), n_outer, wlen, f_src,
                                          inner, src_mod, src_off, _stream()), "emote_gather_frames")
    return dst


def scatter_add_frames(src: torch.Tensor, dst: torch.Tensor, frame_idx: torch.Tensor, n_outer: int, f_dst: int,
                       inner: int, dst_off: int = 0) -> torch.Tensor:
    """dst[o + dst_off, frame_idx[j], :] += src[o, j, :] over [outer, frames, inner] views (fp32)."""
    _req(src, F32, "scatter_add_frames.src"), _req(dst, F32, "scatter_add_frames.dst")
    _req(frame_idx, torch.int32, "scatter_add_frames.frame_idx")
    wlen = frame_idx.numel()
    if src.numel() != n_outer * wlen * inner or (n_outer + dst_off) * f_dst * inner > dst.numel():
        raise _lib.EmoteKernelError("scatter_add_frames: buffer sizes do not match the [outer, frames, inner] geometry")
    check(_lib.load().emote_scatter_add_frames(src.data_ptr(), dst.data_ptr(), frame_idx.data_ptr(), n_outer, wlen, f_dst,
                                               inner, dst_off, _stream()), "emote_scatter_add_frames")
    return dst


def fill_f32(t: torch.Tensor, value: float) -> torch.Tensor:
    _req(t, F32, "fill_f32.t")
    check(_lib.load().emote_fill_f32(t.data_ptr(), float(value), t.numel(), _stream()), "emote_fill_f32")
    return t


def ddim_sigma(alpha_t: float, alpha_prev: float, eta: float) -> float:
    """std of the stochastic DDIM term (Song et al. 2021 eq. 16; diffusers `DDIMScheduler._get_variance`)"""
    if eta == 0.0:
        return 0.0
    var = (1.0 - alpha_prev) / (1.0 - alpha_t) * (1.0 - alpha_t / alpha_prev)
    return eta * math.sqrt(max(var, 0.0))


def cfg_ddim_step(latents: torch.Tensor, noise_pred: torch.Tensor, counter: Optional[torch.Tensor], guidance: float,
                  alpha_t: float, alpha_prev: float, zero_noise_pred: bool = False, noise: Optional[torch.Tensor] = None,
                  sigma: float = 0.0) -> torch.Tensor:
    """latents [1|B, C, F, H, W] fp32 updated in place; noise_pred [2*B, C, F, H, W] (uncond first), cleared as it is
    consumed when zero_noise_pred; noise / sigma: the stochastic term of DDIM with eta > 0."""
    _req(latents, F32, "cfg_ddim_step.latents"), _req(noise_pred, F32, "cfg_ddim_step.noise_pred")
    n = latents.numel()
    if noise_pred.numel() != 2 * n:
        raise _lib.EmoteKernelError("cfg_ddim_step: noise_pred must hold the (uncond, cond) pair of the latents")
    if noise is not None:
        _req(noise, F32, "cfg_ddim_step.noise")
    n_frames = latents.shape[2]
    inner = latents.shape[3] * latents.shape[4]
    check(_lib.load().emote_cfg_ddim_step(latents.data_ptr(), noise_pred.data_ptr(), _ptr(counter), n, n_frames, inner,
                                          guidance, alpha_t, alpha_prev, _ptr(noise), sigma, 1 if zero_noise_pred else 0,
                                          _stream()), "emote_cfg_ddim_step")
    return latents


def ddim_step(latents: torch.Tensor, eps: torch.Tensor, alpha_t: float, alpha_prev: float,
              noise: Optional[torch.Tensor] = None, sigma: float = 0.0) -> torch.Tensor:
    """plain DDIM update of `latents` (in place) with a ready epsilon of the same size"""
    _req(latents, F32, "ddim_step.latents"), _req(eps, F32, "ddim_step.eps")
    if eps.numel() != latents.numel():
        raise _lib.EmoteKernelError("ddim_step: eps and latents differ in size")
    if noise is not None:
        _req(noise, F32, "ddim_step.noise")
    check(_lib.load().emote_ddim_step(latents.data_ptr(), eps.data_ptr(), latents.numel(), alpha_t, alpha_prev, _ptr(noise),
                                      sigma, _stream()), "emote_ddim_step")
    return latents


def vae_postprocess(tok: torch.Tensor, n_img: int, H: int, W: int, want_f32: bool = True, want_u8: bool = False):
    _req(tok, F32, "vae_postprocess.tok")
    ld = tok.shape[-1]
    of = torch.empty((n_img, 3, H, W), dtype=F32, device=tok.device) if want_f32 else None
    ou = torch.empty((n_img, 3, H, W), dtype=torch.uint8, device=tok.device) if want_u8 else None
    check(_lib.load().emote_vae_postprocess(tok.data_ptr(), n_img, H * W, ld, _ptr(of), _ptr(ou), _stream()),
          "emote_vae_postprocess")
    return of, ou


def video_grid_u8(videos: torch.Tensor, nrow: int = 6, padding: int = 2, rescale: bool = False) -> torch.Tensor:
    """videos [b, c, t, h, w] fp32 -> uint8 [t, Hg, Wg, 3] frames (make_grid + truncating cast, utils/util.py:21-30)."""
    _req(videos, F32, "video_grid_u8.videos")
    if videos.dim() != 5 or videos.shape[1] not in (1, 3):
        raise _lib.EmoteKernelError("video_grid_u8: videos must be [b, 1|3, t, h, w]")
    b, c, t, h, w = videos.shape
    xmaps = min(nrow, b)
    ymaps = -(-b // xmaps)
    hg, wg = (h, w) if b == 1 else ((h + padding) * ymaps + padding, (w + padding) * xmaps + padding)
    out = torch.empty((t, hg, wg, 3), dtype=torch.uint8, device=videos.device)
    check(_lib.load().emote_video_grid_u8(videos.data_ptr(), b, c, t, h, w, nrow, padding, int(bool(rescale)), out.data_ptr(),
                                          _stream()), "emote_video_grid_u8")
    return out


# ----------------------------------------------------------------------------------------------- audio front-end
def wave_stats(wave: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    _req(wave, F32, "wave_stats.wave")
    stats = torch.empty(2, dtype=F32, device=wave.device)
    check(_lib.load().emote_wave_stats(wave.data_ptr(), wave.numel(), eps, stats.data_ptr(), _stream()), "emote_wave_stats")
    return stats


def wave_im2col(wave: torch.Tensor, stats: Optional[torch.Tensor], kernel: int, stride: int, kpad: int) -> torch.Tensor:
    _req(wave, F32, "wave_im2col.wave")
    t_out = (wave.numel() - kernel) // stride + 1
    out = torch.empty((t_out, kpad), dtype=OP16, device=wave.device)
    check(_lib.load().emote_wave_im2col(wave.data_ptr(), wave.numel(), _ptr(stats), kernel, stride, kpad, out.data_ptr(),
                                        _stream()), "emote_wave_im2col")
    return out


def channel_norm_gelu(x: torch.Tensor, gamma, beta, eps: float = 1e-5) -> torch.Tensor:
    """GroupNorm(num_groups == channels) over the rows of x [T, C] + GELU -> op16"""
    _req(x, F32, "channel_norm_gelu.x")
    T, Cc = x.shape
    scratch = torch.empty(2 * Cc, dtype=torch.float64, device=x.device)
    out = torch.empty((T, Cc), dtype=OP16, device=x.device)
    check(_lib.load().emote_channel_norm_gelu(x.data_ptr(), T, Cc, gamma.data_ptr(), beta.data_ptr(), eps,
                                              scratch.data_ptr(), out.data_ptr(), _stream()), "emote_channel_norm_gelu")
    return out


def tokens_to_groups(x: torch.Tensor, groups: int, pad_front: int, pad_back: int) -> torch.Tensor:
    _req(x, F32, "tokens_to_groups.x")
    T, Cc = x.shape
    out = torch.empty((groups, T + pad_front + pad_back, Cc // groups), dtype=OP16, device=x.device)
    check(_lib.load().emote_tokens_to_groups(x.data_ptr(), T, Cc, groups, pad_front, pad_back, out.data_ptr(), _stream()),
          "emote_tokens_to_groups")
    return out


def speed_encoder(speeds, centers, radii, w1, b1, w2, b2) -> torch.Tensor:
    for t, n in ((speeds, "speeds"), (centers, "centers"), (radii, "radii"), (w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2")):
        _req(t, F32, f"speed_encoder.{n}")
    out = torch.empty((speeds.numel(), w2.shape[0]), dtype=F32, device=speeds.device)
    check(_lib.load().emote_speed_encoder(speeds.data_ptr(), speeds.numel(), centers.data_ptr(), radii.data_ptr(),
                                          centers.numel(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                          w2.shape[0], out.data_ptr(), _stream()), "emote_speed_encoder")
    return out


# ----------------------------------------------------------------------------------------------- profiling
class KernelProfiler:
    """Brackets every C-ABI launch with CUDA events on the launching stream (bench.py roofline pass, dev profiling).

    with KernelProfiler() as prof: model(...)
    prof.summary() -> {entry point: (launches, total ms)};  prof.gemm_flops / prof.gemm_ms for the tcgen05 GEMM.
    """

    def __init__(self):
        self.records = []
        self._orig = {}

    def __enter__(self):
        lib = _lib.load()
        for name in _lib.SIGNATURES:
            fn = getattr(lib, name)
            self._orig[name] = fn

            def wrapped(*args, _fn=fn, _name=name):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                meta = None
                if _name == "emote_gemm_bf16":
                    a = args[3]._obj
                    meta = (a.M, a.N, a.K, a.conv_taps)
                elif _name in ("emote_attention_bf16", "emote_attention_tc_bf16"):
                    a = args[0]._obj
                    meta = (a.batch, a.heads, a.head_dim, a.nq, a.n0, a.n1)
                else:
                    meta = tuple(x for x in args if isinstance(x, int) and x < (1 << 40))[:6]
                e0.record()
                rc = _fn(*args)
                e1.record()
                self.records.append((_name, e0, e1, meta))
                return rc

            setattr(lib, name, wrapped)
        return self

    def __exit__(self, *exc):
        lib = _lib.load()
        for name, fn in self._orig.items():
            setattr(lib, name, fn)
        torch.cuda.synchronize()
        self.times = [(n, e0.elapsed_time(e1), meta) for n, e0, e1, meta in self.records]
        return False

    def summary(self):
        out = {}
        for n, ms, _ in self.times:
            c, t = out.get(n, (0, 0.0))
            out[n] = (c + 1, t + ms)
        return out

    @property
    def gemm_ms(self):
        return sum(ms for n, ms, _ in self.times if n == "emote_gemm_bf16")

    @property
    def gemm_flops(self):
        return sum(2.0 * m[0] * m[1] * m[2] for n, _, m in self.times if n == "emote_gemm_bf16")

    def by_shape(self, name="emote_gemm_bf16"):
        out = {}
        for n, ms, m in self.times:
            if n == name:
                c, t = out.get(m, (0, 0.0))
                out[m] = (c + 1, t + ms)
        return out
