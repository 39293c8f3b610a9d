"""Pins the BLOCK ORDER of oracle/vae_decoder.py (decoder and encoder) on an independent implementation of the published
VAE architecture.

`diffusers.AutoencoderKL` — the class the reference calls (EMOAnimationPipeline.py:291-307, 402-414) — is not installed and
not vendored in the reference.  It is a re-implementation of the CompVis latent-diffusion `Encoder` / `Decoder`
(conv_in -> [down blocks] -> mid (resnet, attention, resnet) -> [up blocks] -> norm -> swish -> conv_out), and SD-VAE
checkpoints are converted between the two by a pure key renaming (diffusers' `convert_ldm_vae_checkpoint`).  This image
ships that original architecture as third-party code: torchtitan/experiments/flux/model/autoencoder.py (the Flux
autoencoder = the same Encoder / Decoder with 16 latent channels and no quant convs).  The test builds it at SD-VAE
geometry (4 latent channels, ch_mult 1-2-4-4, 2 resnets per level), renames its random weights to diffusers keys with the
published mapping, and requires the oracle to reproduce its outputs.  Leaf arithmetic was already pinned on the
reference's own modules (tests/test_oracle.py); with this the VAE oracle has no unpinned part left.
"""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

from oracle.vae_decoder import VAEDecoderOracle


def _load_ldm_autoencoder():
    spec = importlib.util.find_spec("torchtitan")
    if spec is None or not spec.submodule_search_locations:
        pytest.skip("torchtitan (third-party copy of the CompVis autoencoder) is not in this image")
    path = Path(list(spec.submodule_search_locations)[0]) / "experiments" / "flux" / "model" / "autoencoder.py"
    if not path.exists():
        pytest.skip(f"{path} not found")
    mspec = importlib.util.spec_from_file_location("_ldm_autoencoder", path)   # the file alone: no torchtitan package import
    mod = importlib.util.module_from_spec(mspec)
    sys.modules[mspec.name] = mod          # dataclasses resolve the module by name while the file executes
    mspec.loader.exec_module(mod)
    return mod


def _ldm_to_diffusers(sd, prefix, n_levels):
    """the published key renaming (convert_ldm_vae_checkpoint): levels of the decoder are stored low-to-high resolution in
    the LDM module list and high-to-low in diffusers' up_blocks; 1x1-conv attention projections become Linear weights"""
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if parts[0] == "mid":
            if parts[1].startswith("block_"):
                nk = f"mid_block.resnets.{int(parts[1][-1]) - 1}." + ".".join(parts[2:])
            else:
                name = {"norm": "group_norm", "q": "query", "k": "key", "v": "value", "proj_out": "proj_attn"}[parts[2]]
                nk = f"mid_block.attentions.0.{name}.{parts[3]}"
                if v.dim() == 4:
                    v = v[:, :, 0, 0]
        elif parts[0] in ("up", "down"):
            lvl = int(parts[1])
            blk = f"up_blocks.{n_levels - 1 - lvl}" if parts[0] == "up" else f"down_blocks.{lvl}"
            if parts[2] == "block":
                nk = f"{blk}.resnets.{parts[3]}." + ".".join(parts[4:])
            else:
                nk = f"{blk}.{'upsamplers' if parts[2] == 'upsample' else 'downsamplers'}.0." + ".".join(parts[3:])
        elif parts[0] == "norm_out":
            nk = "conv_norm_out." + parts[1]
        else:
            nk = k                                                    # conv_in / conv_out
        out[f"{prefix}.{nk.replace('nin_shortcut', 'conv_shortcut')}"] = v
    return out


def _randomise(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in module.parameters():
            if p.dim() == 1:      # norm scales / shifts and biases away from their (1, 0) defaults: order must matter
                p.copy_(torch.randn(p.shape, generator=g) * 0.3 + (1.0 if p.numel() >= 32 else 0.0))
            else:
                fan = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) * fan ** -0.5)


@pytest.mark.parametrize("ch,res", [(32, 32), (64, 16)])
def test_vae_oracle_block_order_matches_the_compvis_autoencoder(ch, res):
    ldm = _load_ldm_autoencoder()
    kw = dict(ch=ch, ch_mult=[1, 2, 4, 4], num_res_blocks=2, in_channels=3, resolution=res, z_channels=4)
    dec = ldm.Decoder(out_ch=3, **kw).eval()
    enc = ldm.Encoder(**kw).eval()
    _randomise(dec, 1), _randomise(enc, 2)
    sd = _ldm_to_diffusers(dec.state_dict(), "decoder", 4)
    sd.update(_ldm_to_diffusers(enc.state_dict(), "encoder", 4))
    eye = torch.eye(8)[:, :, None, None]
    sd.update({"quant_conv.weight": eye, "quant_conv.bias": torch.zeros(8)})       # the Flux variant has no quant convs
    oracle = VAEDecoderOracle(sd)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(2, 4, res // 8, res // 8, generator=g)
    x = torch.rand(2, 3, res, res, generator=g) * 2 - 1
    with torch.no_grad():
        want_d, want_e = dec(z), enc(x)
    got_d, got_e = oracle.decode(z), oracle.encode(x)
    assert got_d.shape == want_d.shape == (2, 3, res, res) and got_e.shape == want_e.shape == (2, 8, res // 8, res // 8)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    assert rel(got_d, want_d) < 2e-5, rel(got_d, want_d)
    assert rel(got_e, want_e) < 2e-5, rel(got_e, want_e)
    # and the order does matter for these weights: swapping two resnets of one level moves the output
    swapped = dict(sd)
    for k in list(sd):
        if k.startswith("decoder.up_blocks.1.resnets.1."):
            other = k.replace("resnets.1.", "resnets.2.")
            swapped[k], swapped[other] = sd[other], sd[k]
    assert rel(VAEDecoderOracle(swapped).decode(z), want_d) > 1e-2
