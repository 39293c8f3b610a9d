"""TEST INFRASTRUCTURE ONLY — oracles for the audio front-end.

* wav2vec2: the reference delegates to the third-party `transformers.Wav2Vec2Model` (Net.py:611-612, 644; unpinned in
  requirements.txt).  transformers IS installed in this image (5.5.0), so the oracle is that class itself, executed on
  the CPU in fp32 with random-init weights of the wav2vec2-base geometry (`facebook/wav2vec2-base-960h` cannot be
  downloaded: no network) together with the processor's normalisation restated from
  `Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm`: (x - mean) / sqrt(var + 1e-7).
* SpeedEncoder: the reference class (Net.py:198-258) is cut out of Net.py with `ast` (Net.py itself cannot be imported: decord
  / mediapipe / diffusers are absent), its `ModelMixin` base swapped for `nn.Module`, and executed unchanged;
  `speed_encoder_restated` is the travelling restatement pinned on it (tests/golden/speed_encoder.pt + live test).
"""
from __future__ import annotations

import ast
from pathlib import Path

import torch
from torch import nn

REFERENCE_NET = Path("/root/reference/Net.py")


def hf_wav2vec2(seed: int = 0, **config_overrides):
    """random-init transformers.Wav2Vec2Model (wav2vec2-base geometry), eval mode, fp32, CPU"""
    import transformers
    cfg = transformers.Wav2Vec2Config(**config_overrides)
    torch.manual_seed(seed)
    m = transformers.Wav2Vec2Model(cfg).eval()
    # the default init leaves most projections at std 0.02 and the conv stack at kaiming scale: keep it (finite, O(1) outputs)
    return m


def normalize_waveform(x: torch.Tensor) -> torch.Tensor:
    """transformers Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm (do_normalize=True of wav2vec2-base-960h)"""
    return (x - x.mean()) / torch.sqrt(x.var(unbiased=False) + 1e-7)


def reference_speed_encoder_class():
    """Net.py:198-258 `SpeedEncoder`, executed from the reference source with nn.Module as its base"""
    tree = ast.parse(REFERENCE_NET.read_text())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SpeedEncoder")
    cls.bases = [ast.Attribute(value=ast.Name(id="nn", ctx=ast.Load()), attr="Module", ctx=ast.Load())]
    ns = {"torch": torch, "nn": nn}
    exec(compile(ast.fix_missing_locations(ast.Module(body=[cls], type_ignores=[])), str(REFERENCE_NET), "exec"), ns)
    return ns["SpeedEncoder"]


def speed_encoder_restated(speeds: torch.Tensor, centers, radii, w1, b1, w2, b2) -> torch.Tensor:
    """Net.py:232-258: v_i = tanh((s - c_i)/r_i * 3); Linear -> ReLU -> Linear"""
    v = torch.stack([torch.tanh((speeds - c) / r * 3) for c, r in zip(centers, radii)], dim=1)
    return torch.relu(v @ w1.t() + b1) @ w2.t() + b2
