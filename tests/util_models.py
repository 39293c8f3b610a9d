"""Shared test fixtures: seeded configs / weights / inputs for the UNet3D path (no reference needed)."""
import json
import os
from pathlib import Path

import torch

# ------------------------------------------------------------------------------------------------ parity limits
# Every parity check of the GPU tests goes through check_parity(key, err, default): the limit is 1.5 x the error MEASURED
# on a B200 for this operand type (tests/parity_limits.json, written by scripts/update_parity_limits.py from the log of a
# `pytest -m gpu` run with EMOTE_PARITY_LOG set), or `default` for a key that has no measurement yet.
_LIMITS_FILE = Path(__file__).resolve().parent / "parity_limits.json"
_LIMITS = json.loads(_LIMITS_FILE.read_text()) if _LIMITS_FILE.exists() else {}


def operand() -> str:
    from emote_hack_b200 import _lib
    return _lib.OPERAND


def parity_limit(key: str, default: float) -> float:
    return float(_LIMITS.get(operand(), {}).get(key, {}).get("limit", default))


_auto_counts = {}


def chk(err: float, default: float) -> float:
    """check_parity keyed by the running test's id (+ a per-test counter): for parametrised kernel tests"""
    cur = os.environ.get("PYTEST_CURRENT_TEST", "unknown").split(" ")[0].split("::", 1)[-1]
    n = _auto_counts.get(cur, 0)
    _auto_counts[cur] = n + 1
    return check_parity(f"k.{cur}#{n}".replace(" ", ""), err, default)


def record_only(key: str, value: float) -> float:
    """log a context figure (e.g. what eager PyTorch fp16 gives on the same network) without gating on it"""
    log = os.environ.get("EMOTE_PARITY_LOG")
    if log:
        with open(log, "a") as fh:
            fh.write(f"{operand()} info.{key} {value:.6e} nan\n")
    print(f"info[{operand()}] {key}: {value:.3e}")
    return value


def check_parity(key: str, err: float, default: float) -> float:
    """assert err < limit(key); logs `operand key err limit` to $EMOTE_PARITY_LOG (one line per check)"""
    lim = parity_limit(key, default)
    log = os.environ.get("EMOTE_PARITY_LOG")
    if log:
        with open(log, "a") as fh:
            fh.write(f"{operand()} {key} {err:.6e} {lim:.6e}\n")
    print(f"parity[{operand()}] {key}: {err:.3e} (limit {lim:.3e})")
    assert err < lim, f"{key}: rel-L2 {err:.3e} exceeds the limit {lim:.3e} ({operand()} operands)"
    return err


MM_KW = dict(num_attention_heads=4, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
             temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1)

TINY_CFG = dict(sample_size=8, cross_attention_dim=64, block_out_channels=(64, 128, 128, 128), attention_head_dim=4,
                norm_num_groups=32, use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8),
                motion_module_type="Vanilla", motion_module_kwargs=MM_KW, unet_use_cross_frame_attention=False,
                unet_use_temporal_attention=False)

# SD-1.5 widths + configs/inference.yaml kwargs with motion modules on (BASELINE config #2 network)
FULL_MM_KW = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
                  temporal_position_encoding=True, temporal_position_encoding_max_len=24, temporal_attention_dim_div=1)
FULL_CFG = dict(sample_size=64, cross_attention_dim=768, use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8),
                motion_module_mid_block=False, motion_module_decoder_only=False, motion_module_type="Vanilla",
                motion_module_kwargs=FULL_MM_KW, unet_use_cross_frame_attention=False, unet_use_temporal_attention=False)


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def rerandomise_zero_inits(model, seed=1, std=0.05):
    """temporal_transformer.proj_out is zero-initialised (motion_module.py:79-80): without this the temporal path
    contributes exactly 0 and temporal bugs are invisible (SURVEY.md §7 hard part 6)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "temporal_transformer.proj_out" in n:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return model


def make_inputs(b=2, f=4, hw=8, ctx_tokens=7, ctx_dim=64, seed=1234, per_frame_ctx=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(b, 4, f, hw, hw, generator=g)
    ctx = torch.randn(b * f if per_frame_ctx else b, ctx_tokens, ctx_dim, generator=g)
    return x, ctx


def reader_block_names(model):
    """names of the mid+up BasicTransformerBlocks, i.e. the ReferenceAttentionControl(fusion_blocks='midup') readers"""
    return [n for n, m in model.named_modules()
            if (n.startswith("mid_block") or n.startswith("up_blocks")) and n.endswith("transformer_blocks.0")
            and "temporal" not in n]


def make_banks(model, latent_hw, seed=7):
    """synthetic ReferenceNet banks [2, HW_level, C_level] per reader block; seeded per block
    NAME so the reference model, the oracle port and the CUDA modules get identical banks."""
    import zlib
    mods = dict(model.named_modules())
    banks = {}
    for name in reader_block_names(model):
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 104729 * seed) % (2 ** 31))
        c = mods[name].norm1.normalized_shape[0]
        if name.startswith("mid_block"):
            side = latent_hw // 8
        else:
            side = latent_hw // (8 >> int(name.split(".")[1]))
        banks[name] = [torch.randn(2, side * side, c, generator=g)]
    return banks


def seeded_state_dict(shapes, seed=0, keep=None):
    """Deterministic, name-keyed random weights: the reference (in the build container), the oracle port and the CUDA
    modules all get IDENTICAL parameters without shipping a checkpoint.  `shapes`: {key: shape}; `keep`: tensors to
    copy verbatim (deterministic buffers such as the sinusoidal `pos_encoder.pe`)."""
    import zlib
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if keep is not None and name in keep:
            out[name] = keep[name].clone()
            continue
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
        if name.endswith(".pe"):
            raise KeyError(f"{name}: positional tables must be passed through `keep`")
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            out[name] = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        elif name.endswith("weight") and ("norm" in name):
            out[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            out[name] = 0.05 * torch.randn(shape, generator=g)
    return out


def sinusoid_pe(max_len, d_model):
    """motion_module.py:239-244 (deterministic buffer)"""
    import math
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def seeded_unet_state_dict(shapes, seed=0):
    keep = {k: sinusoid_pe(s[1], s[2]) for k, s in shapes.items() if k.endswith(".pe")}
    return seeded_state_dict(shapes, seed, keep)


def controlnet_residuals(seed=9, frames=2):
    """ControlNet residuals for TINY_CFG on an 8x8 latent (unet_controlnet.py:336-337, 414-448): 12 down + 1 mid"""
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 8), (64, 8), (64, 8), (64, 4), (128, 4), (128, 4), (128, 2), (128, 2), (128, 2), (128, 1), (128, 1), (128, 1)]
    down = [torch.randn(2, c, frames, s, s, generator=g) * 0.1 for c, s in shapes]
    mid = torch.randn(2, 128, frames, 1, 1, generator=g) * 0.1
    return down, mid


def writer_cfg():
    """tiny ReferenceNet writer = TINY_CFG without motion modules (a 2-D SD UNet run as a one-frame UNet3D)"""
    cfg = dict(TINY_CFG)
    cfg.update(use_motion_module=False, motion_module_type=None, motion_module_kwargs={})
    return cfg


def writer_inputs(hw=16, seed=5):
    """reference-image latents [2, 4, hw, hw] (CFG pair) and text context for the writer fixtures"""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(2, 4, hw, hw, generator=g), torch.randn(2, 7, 64, generator=g)


APPEARANCE_TRIMMED = ("conv_norm_out.", "conv_out.", "up_blocks.3.attentions.2.proj_out.",
                      "up_blocks.3.attentions.2.transformer_blocks.0.attn1.", "up_blocks.3.attentions.2.transformer_blocks.0.attn2.",
                      "up_blocks.3.attentions.2.transformer_blocks.0.norm2.", "up_blocks.3.attentions.2.transformer_blocks.0.norm3.",
                      "up_blocks.3.attentions.2.transformer_blocks.0.ff.")


def appearance_cfg():
    """AppearanceEncoderModel kwargs matching writer_cfg()"""
    return dict(sample_size=8, cross_attention_dim=64, block_out_channels=(64, 128, 128, 128), attention_head_dim=4,
                norm_num_groups=32)
