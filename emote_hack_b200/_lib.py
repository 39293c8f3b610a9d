"""ctypes binding of libemote_b200.so (the C ABI declared in include/emote_b200.h).

The library is the product: there is no CPU or eager-PyTorch fallback.  `load()` raises when the shared
object is missing, and every op raises `EmoteKernelError` when a kernel call reports an error.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# Tensor-core operand type of the process, fixed at import: EMOTE_OPERAND=fp16 (default; libemote_b200.so — the reference
# pipeline itself runs fp16 and 10 mantissa bits keep the whole network within ~1e-3 of the fp32 reference) or bf16
# (libemote_b200_bf16.so: same kernels built with -DEMOTE_OPERAND_BF16, ~8x the rounding error, wider exponent range).
OPERAND = os.environ.get("EMOTE_OPERAND", "fp16").lower()
if OPERAND not in ("fp16", "bf16"):
    raise ValueError(f"EMOTE_OPERAND must be 'fp16' or 'bf16', got {OPERAND!r}")
# EMOTE_B200_LIB: dev override (same-box A/B of two builds of the library); the default is the in-tree build
LIB_PATH = Path(os.environ.get("EMOTE_B200_LIB") or
                (_HERE / "lib" / ("libemote_b200.so" if OPERAND == "fp16" else "libemote_b200_bf16.so")))
CSRC = _HERE / "csrc"


class EmoteKernelError(RuntimeError):
    pass


class EmoteGemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("lda", C.c_int32),
        ("conv_taps", C.c_int32),
        ("n_img", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("bias", C.c_void_p),
        ("row_bias", C.c_void_p),
        ("rows_per_group", C.c_int32),
        ("residual", C.c_void_p),
        ("ldr", C.c_int32),
        ("out_scale", C.c_float),
        ("epilogue", C.c_int32),
        ("out_dtype", C.c_int32),
        ("ldc", C.c_int32),
        ("block_n", C.c_int32),
        ("pair_mode", C.c_int32),
        ("tma_store", C.c_int32),
        ("colstats", C.c_void_p),
        ("stats_rows", C.c_int32),
    ]


class EmoteAttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k0", C.c_void_p), ("v0", C.c_void_p), ("k1", C.c_void_p), ("v1", C.c_void_p),
        ("out", C.c_void_p),
        ("batch", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("nq", C.c_int32), ("n0", C.c_int32), ("n1", C.c_int32),
        ("q_batch_stride", C.c_int64), ("q_row_stride", C.c_int64),
        ("kv0_batch_stride", C.c_int64), ("kv0_row_stride", C.c_int64),
        ("kv1_batch_stride", C.c_int64), ("kv1_row_stride", C.c_int64),
        ("o_batch_stride", C.c_int64), ("o_row_stride", C.c_int64),
        ("kv0_batch_div", C.c_int32), ("kv1_batch_div", C.c_int32), ("kv1_first_batch", C.c_int32),
        ("scale", C.c_float),
    ]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> argtypes; every entry returns int except the three introspection calls.  Must list every symbol
# declared in include/emote_b200.h (tests/test_abi.py checks that against the header text).
SIGNATURES = {
    "emote_gemm_bf16": [_vp, _vp, _vp, C.POINTER(EmoteGemmArgs), _vp],
    "emote_gn_stats": [_vp, _i32, _i32, _i32, _i32, _i64, _i32, _vp, _i32, _vp],
    "emote_gn_colstats_reduce": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "emote_gn_apply": [_vp, _i32, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _f32, _i32, _vp, _vp, _vp],
    "emote_layernorm": [_vp, _i64, _i32, _vp, _vp, _f32, _vp, _i32, _i32, _vp, _vp],
    "emote_layernorm_dual": [_vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _vp],
    "emote_wave_stats": [_vp, _i64, _f32, _vp, _vp],
    "emote_wave_im2col": [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp],
    "emote_channel_norm_gelu": [_vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _vp],
    "emote_tokens_to_groups": [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp],
    "emote_speed_encoder": [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp],
    "emote_attention_bf16": [C.POINTER(EmoteAttnArgs), _vp],
    "emote_attention_tc_bf16": [C.POINTER(EmoteAttnArgs), _vp],
    "emote_attention_tc_supported": [_i32],
    "emote_attention_wide_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _i64, _i64, _i64, _f32, _vp],
    "emote_temporal_attention_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "emote_softmax_rows_bf16": [_vp, _i64, _i32, _f32, _vp, _vp],
    "emote_latent_im2col": [_vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _vp],
    "emote_im2col3x3": [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp],
    "emote_im2col3x3_bf16": [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp],
    "emote_im2col3x3_s2_pad01": [_vp, _i32, _i32, _i32, _i32, _vp, _vp],
    "emote_upsample2x": [_vp, _i32, _i32, _i32, _i32, _vp, _vp],
    "emote_upsample_nearest": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp],
    "emote_cast_bf16": [_vp, _i64, _i32, _i32, _i32, _vp, _vp],
    "emote_silu_bf16": [_vp, _i64, _vp, _vp],
    "emote_tokens_to_ncfhw": [_vp, _i32, _i32, _i32, _i32, _vp, _vp],
    "emote_ncfhw_to_tokens": [_vp, _i32, _i32, _i32, _i32, _vp, _vp],
    "emote_add_f32": [_vp, _vp, _vp, _i64, _vp],
    "emote_timestep_embedding": [_vp, _i32, _i32, _i32, _f32, _vp, _vp],
    "emote_cfg_ddim_step": [_vp, _vp, _vp, _i64, _i32, _i64, _f32, _f32, _f32, _vp, _f32, _i32, _vp],
    "emote_ddim_step": [_vp, _vp, _i64, _f32, _f32, _vp, _f32, _vp],
    "emote_gather_frames": [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i32, _i32, _vp],
    "emote_scatter_add_frames": [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i32, _vp],
    "emote_fill_f32": [_vp, _f32, _i64, _vp],
    "emote_vae_postprocess": [_vp, _i32, _i32, _i32, _vp, _vp, _vp],
    "emote_video_grid_u8": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp],
}
INTROSPECTION = {
    "emote_last_error": (C.c_char_p, []),
    "emote_launch_count": (C.c_longlong, []),
    "emote_abi_version": (C.c_int, []),
    "emote_operand_dtype": (C.c_int, []),
    "emote_set_pdl": (None, [C.c_int]),
    "emote_set_tuning": (C.c_int, [C.c_char_p, C.c_int32]),
}

_lib = None


def build(verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into lib/libemote_b200.so (fp16 operands) and lib/libemote_b200_bf16.so (nvcc
    cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(CSRC), "-j8"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building libemote_b200.so failed:\n{r.stdout[-4000:]}\n{r.stderr[-4000:]}")
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise EmoteKernelError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / eager fallback for the hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (restype, argtypes) in INTROSPECTION.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    want = 2 if OPERAND == "fp16" else 1   # EMOTE_OP_F16 / EMOTE_OP_BF16
    if "EMOTE_B200_LIB" not in os.environ and lib.emote_operand_dtype() != want:
        raise EmoteKernelError(f"{LIB_PATH} was built for a different operand type than EMOTE_OPERAND={OPERAND}")
    _lib = lib
    return lib


def op16_torch_dtype():
    """torch dtype of the 16-bit tensor-core operand buffers of this process"""
    import torch
    return torch.float16 if OPERAND == "fp16" else torch.bfloat16


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().emote_last_error().decode("utf-8", "replace")
        raise EmoteKernelError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(load().emote_launch_count())
