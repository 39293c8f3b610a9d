"""TEST INFRASTRUCTURE ONLY — loader for the *untouched* reference modules under /root/reference.

The reference imports `diffusers`, which is not installed in this image.  This module registers a fake
`diffusers` package in `sys.modules` that contains NO arithmetic: config/mixin plumbing only, plus aliases that
point `diffusers.models.attention.*` / `diffusers.models.embeddings.*` at the reference's OWN vendored copies
(`magicanimate/models/orig_attention.py`, `magicanimate/models/embeddings.py`).  With that, the reference's
`UNet3DConditionModel` and `ReferenceAttentionControl` run on CPU and serve to (a) pin `oracle/unet3d_port.py`
and (b) generate the golden vectors in `tests/golden/` (see `oracle/make_golden.py`).

/root/reference exists only in the build container; nothing that runs on the GPU box may import this file.
"""
from __future__ import annotations

import functools
import inspect
import sys
import types
from collections import OrderedDict
from dataclasses import fields, is_dataclass
from pathlib import Path

import torch
from torch import nn

REFERENCE_ROOT = Path("/root/reference")


def reference_available() -> bool:
    return (REFERENCE_ROOT / "magicanimate" / "models" / "unet_controlnet.py").is_file()


class _Config(dict):
    """attribute-style dict; unknown attributes raise AttributeError (deepcopy/pickle rely on that)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _register_to_config(init):
    @functools.wraps(init)
    def wrapped(self, *args, **kwargs):
        bound = inspect.signature(init).bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        self._internal_dict = _Config(cfg)
        init(self, *args, **kwargs)

    return wrapped


class _ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self._internal_dict

    @classmethod
    def from_config(cls, config, **kwargs):
        params = inspect.signature(cls.__init__).parameters
        merged = {k: v for k, v in dict(config).items() if k in params}
        merged.update({k: v for k, v in kwargs.items() if k in params})
        return cls(**merged)


class _ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class _BaseOutput(OrderedDict):
    def __post_init__(self):
        if is_dataclass(self):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    self[f.name] = v

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())

    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]


class _Logger:
    def __getattr__(self, name):
        return lambda *a, **k: None


def _mod(name: str, **attrs) -> types.ModuleType:
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave as a package so sub-imports resolve through sys.modules
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(_mod(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_installed = False


def install() -> None:
    """Register the fake diffusers package and put /root/reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("/root/reference is not present (the reference only exists in the build container)")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))

    _mod("diffusers", __version__="0.0.shim")
    _mod("diffusers.configuration_utils", ConfigMixin=_ConfigMixin, register_to_config=_register_to_config,
         FrozenDict=_Config)
    _mod("diffusers.models")
    _mod("diffusers.models.modeling_utils", ModelMixin=_ModelMixin)
    _mod("diffusers.utils", BaseOutput=_BaseOutput, logging=types.SimpleNamespace(get_logger=lambda *_: _Logger()),
         WEIGHTS_NAME="diffusion_pytorch_model.bin", deprecate=lambda *a, **k: None, is_accelerate_available=lambda: False)
    _mod("diffusers.utils.import_utils", is_xformers_available=lambda: False)
    _mod("diffusers.modeling_utils", ModelMixin=_ModelMixin)

    # the reference's own vendored arithmetic stands in for the diffusers classes it was copied from
    import magicanimate.models.embeddings as ref_emb  # noqa: E402

    _mod("diffusers.models.embeddings", Timesteps=ref_emb.Timesteps, TimestepEmbedding=ref_emb.TimestepEmbedding,
         ImagePositionalEmbeddings=ref_emb.ImagePositionalEmbeddings)
    import magicanimate.models.orig_attention as ref_attn  # noqa: E402

    _mod("diffusers.models.attention", Attention=ref_attn.CrossAttention, CrossAttention=ref_attn.CrossAttention,
         FeedForward=ref_attn.FeedForward, AdaLayerNorm=ref_attn.AdaLayerNorm,
         BasicTransformerBlock=ref_attn.BasicTransformerBlock)

    # empty stand-ins needed only so `mutual_self_attention.py` imports
    class _Stub:  # noqa: D401
        pass

    _mod("diffusers", StableDiffusionControlNetPipeline=type("StableDiffusionControlNetPipeline", (_Stub,), {}))
    _mod("diffusers.models", ControlNetModel=type("ControlNetModel", (_Stub,), {}))
    _mod("diffusers.models.unet_2d_blocks", **{n: type(n, (nn.Module,), {}) for n in
                                               ("CrossAttnDownBlock2D", "CrossAttnUpBlock2D", "DownBlock2D", "UpBlock2D")})
    _mod("diffusers.pipelines")
    _mod("diffusers.pipelines.controlnet")
    _mod("diffusers.pipelines.controlnet.multicontrolnet", MultiControlNetModel=type("MultiControlNetModel", (_Stub,), {}))
    _mod("diffusers.pipelines.stable_diffusion", StableDiffusionPipelineOutput=type("StableDiffusionPipelineOutput", (_Stub,), {}))
    _mod("diffusers.utils.torch_utils", is_compiled_module=lambda m: False, randn_tensor=None)
    _mod("diffusers.image_processor", VaeImageProcessor=type("VaeImageProcessor", (_Stub,), {}))
    _installed = True


def load_reference_unet_class():
    install()
    from magicanimate.models.unet_controlnet import UNet3DConditionModel

    return UNet3DConditionModel


def load_reference_control_class():
    install()
    from magicanimate.models.mutual_self_attention import ReferenceAttentionControl

    return ReferenceAttentionControl


def load_reference_context_uniform():
    """magicanimate/pipelines/context.py imports nothing from diffusers."""
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    sys.dont_write_bytecode = True
    from magicanimate.pipelines.context import uniform

    return uniform


def build_reference_vae_decoder(state_dict, block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2,
                                groups: int = 32, eps: float = 1e-6):
    """SD-VAE decoder wired out of the reference's OWN leaf modules (no arithmetic here, topology only).

    `diffusers.AutoencoderKL` is absent, but every arithmetic leaf of its decoder has a twin inside /root/reference:
    the 2-D resnet is `ResnetBlock3D` with one frame and no time embedding (resnet.py:114-207), the mid-block attention
    is the legacy `AttentionBlock` (orig_attention.py:253-385), the upsampler is `Upsample3D` (resnet.py:42-84) and
    the plain convs are `InflatedConv3d` (resnet.py:30-39).  Composing them in the published decoder order gives a
    decoder whose every multiply-add is executed by untouched reference code; `oracle/vae_decoder.py` is pinned
    against it (tests/golden/vae_decoder_outputs.pt).  What remains anchored only on the public definition is the
    ORDER of the blocks.  Input / output: [n, 4, 1, h, w] -> [n, 3, 1, 8h, 8w].
    """
    install()
    from magicanimate.models.orig_attention import AttentionBlock
    from magicanimate.models.resnet import InflatedConv3d, ResnetBlock3D, Upsample3D

    def res(cin, cout):
        return ResnetBlock3D(in_channels=cin, out_channels=cout, temb_channels=None, groups=groups, eps=eps)

    class _Attn5D(nn.Module):
        """[n, c, 1, h, w] view around the reference's 4-D AttentionBlock (a reshape, no arithmetic)."""

        def __init__(self, c):
            super().__init__()
            self.inner = AttentionBlock(c, None, groups, 1.0, eps)

        def forward(self, x):
            return self.inner(x[:, :, 0])[:, :, None]

    top = block_out_channels[-1]
    mods = OrderedDict()
    mods["post_quant_conv"] = InflatedConv3d(4, 4, 1)
    mods["decoder.conv_in"] = InflatedConv3d(4, top, 3, padding=1)
    mods["decoder.mid_block.resnets.0"] = res(top, top)
    mods["decoder.mid_block.attentions.0"] = _Attn5D(top)
    mods["decoder.mid_block.resnets.1"] = res(top, top)
    rev = list(reversed(block_out_channels))
    cin = top
    for bi, cout in enumerate(rev):
        for li in range(layers_per_block + 1):
            mods[f"decoder.up_blocks.{bi}.resnets.{li}"] = res(cin, cout)
            cin = cout
        if bi != len(rev) - 1:
            mods[f"decoder.up_blocks.{bi}.upsamplers.0"] = Upsample3D(cout, use_conv=True)
    mods["decoder.conv_norm_out"] = nn.GroupNorm(groups, rev[-1], eps=eps)
    mods["decoder.conv_out"] = InflatedConv3d(rev[-1], 3, 3, padding=1)

    for name, m in mods.items():
        sd = {}
        target = m.inner if isinstance(m, _Attn5D) else m
        for k in target.state_dict():
            sd[k] = state_dict[f"{name}.{k}"]
        target.load_state_dict(sd, strict=True)
        m.eval()

    def decode(z5):
        x = z5
        with torch.no_grad():
            for name, m in mods.items():
                if isinstance(m, ResnetBlock3D):
                    x = m(x, None)
                elif name == "decoder.conv_norm_out":
                    x = torch.nn.functional.silu(m(x))
                else:
                    x = m(x)
        return x

    return decode


def build_reference_vae_encoder(state_dict, block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2,
                                groups: int = 32, eps: float = 1e-6):
    """SD-VAE encoder (+ quant_conv) wired out of the reference's own leaf modules, like build_reference_vae_decoder.
    The stride-2 downsample uses the reference's InflatedConv3d after a (0, 1) zero pad (diffusers Downsample2D with
    padding=0; the reference's Downsample3D refuses padding=0, resnet.py:104-105, so only the conv leaf is reference
    code there).  Input / output: [n, 3, 1, H, W] -> moments [n, 8, 1, H/8, W/8]."""
    install()
    from magicanimate.models.orig_attention import AttentionBlock
    from magicanimate.models.resnet import InflatedConv3d, ResnetBlock3D

    def res(cin, cout):
        return ResnetBlock3D(in_channels=cin, out_channels=cout, temb_channels=None, groups=groups, eps=eps)

    top = block_out_channels[-1]
    mods = OrderedDict()
    mods["encoder.conv_in"] = InflatedConv3d(3, block_out_channels[0], 3, padding=1)
    cin = block_out_channels[0]
    for bi, cout in enumerate(block_out_channels):
        for li in range(layers_per_block):
            mods[f"encoder.down_blocks.{bi}.resnets.{li}"] = res(cin, cout)
            cin = cout
        if bi != len(block_out_channels) - 1:
            mods[f"encoder.down_blocks.{bi}.downsamplers.0.conv"] = InflatedConv3d(cout, cout, 3, stride=2, padding=0)
    mods["encoder.mid_block.resnets.0"] = res(top, top)
    mods["encoder.mid_block.attentions.0"] = AttentionBlock(top, None, groups, 1.0, eps)
    mods["encoder.mid_block.resnets.1"] = res(top, top)
    mods["encoder.conv_norm_out"] = nn.GroupNorm(groups, top, eps=eps)
    mods["encoder.conv_out"] = InflatedConv3d(top, 8, 3, padding=1)
    mods["quant_conv"] = InflatedConv3d(8, 8, 1)
    for name, m in mods.items():
        m.load_state_dict({k: state_dict[f"{name}.{k}"] for k in m.state_dict()}, strict=True)
        m.eval()

    def encode(x5):
        x = x5
        with torch.no_grad():
            for name, m in mods.items():
                if isinstance(m, ResnetBlock3D):
                    x = m(x, None)
                elif isinstance(m, AttentionBlock):
                    x = m(x[:, :, 0])[:, :, None]
                elif name.endswith("downsamplers.0.conv"):
                    x = m(torch.nn.functional.pad(x, (0, 1, 0, 1)))
                elif name == "encoder.conv_norm_out":
                    x = torch.nn.functional.silu(m(x))
                else:
                    x = m(x)
        return x

    return encode
