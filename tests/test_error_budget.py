"""Error budget of the 16-bit tensor-core path, on the CPU: the fp32 oracle with ONLY the operands of its GEMMs / convs /
attention products rounded to the operand type reproduces the error the CUDA path measures on the B200
(tests/parity_limits.json) — i.e. the kernels add nothing beyond operand rounding — and splits it into its sources.
DESIGN.md §5 quotes this table.  Same network / inputs as __graft_entry__.smoke()."""
import json
from pathlib import Path

import torch
import torch.nn.functional as RF

import oracle.unet3d_port as port
from oracle.unet3d_port import UNet3DOracle
from util_models import TINY_CFG, make_inputs, rel_l2, rerandomise_zero_inits


def _rnd(t, dt):
    return t if dt is None else t.to(dt).to(torch.float32)


class _RoundingFunctional:
    """stands in for torch.nn.functional inside the oracle: linear / conv2d round their operands as configured"""

    def __init__(self, cfg):
        self.cfg = cfg

    def __getattr__(self, k):
        return getattr(RF, k)

    def linear(self, x, w, b=None):
        return RF.linear(_rnd(x, self.cfg.get("lin_a")), _rnd(w, self.cfg.get("lin_w")), b)

    def conv2d(self, x, w, b=None, **kw):
        kind = "conv" if w.shape[-1] == 3 else "lin"           # 1x1 convs are the proj_in / proj_out linears
        return RF.conv2d(_rnd(x, self.cfg.get(kind + "_a")), _rnd(w, self.cfg.get(kind + "_w")), b, **kw)


class _RoundedOracle(UNet3DOracle):
    """cfg keys lin_a / lin_w / conv_a / conv_w / attn -> dtype: which operands are rounded (missing = exact fp32)"""

    def __init__(self, sd, config, cfg):
        super().__init__(sd, config)
        self.cfg = cfg

    def _mha(self, p, x, ctx, heads):
        # the CUDA path stores q / k / v as 16-bit GEMM outputs and rounds P in [0, 1] before the P V product; the row
        # sum comes from the rounded P (P x ones on the tensor core)
        dt = self.cfg.get("attn")
        q, k, v = (_rnd(self._lin(p + n, s), dt) for n, s in ((".to_q", x), (".to_k", ctx), (".to_v", ctx)))
        b, n, c = q.shape
        d = c // heads
        sp = lambda t: t.reshape(t.shape[0], t.shape[1], heads, d).transpose(1, 2)
        s = (sp(q) @ sp(k).transpose(-1, -2)) * d ** -0.5
        pexp = _rnd(torch.exp(s - s.amax(-1, keepdim=True)), dt)
        o = (pexp @ sp(v)) / pexp.sum(-1, keepdim=True)
        return self._lin(p + ".to_out.0", o.transpose(1, 2).reshape(b, n, c))

    def forward(self, *a, **k):
        old, port.F = port.F, _RoundingFunctional(self.cfg)
        try:
            return super().forward(*a, **k)
        finally:
            port.F = old

    __call__ = forward


def test_operand_rounding_alone_explains_the_measured_parity():
    from emote_hack_b200.unet3d import UNet3DConditionModel
    torch.manual_seed(0)
    model = rerandomise_zero_inits(UNet3DConditionModel(**TINY_CFG).eval())
    sd, conf = model.state_dict(), dict(model.config)
    x, ctx = make_inputs(2, 4, 16)
    ref = UNet3DOracle(sd, conf)(x, 481, ctx)
    H, B = torch.float16, torch.bfloat16
    every = ("lin_a", "lin_w", "conv_a", "conv_w", "attn")
    rows = {
        "fp16: every operand": {k: H for k in every},
        "bf16: every operand": {k: B for k in every},
        "fp16: weights only": {"lin_w": H, "conv_w": H},
        "fp16: activations + attention operands only": {"lin_a": H, "conv_a": H, "attn": H},
        "fp16: 3x3 convs only": {"conv_a": H, "conv_w": H},
        "fp16: linears / 1x1 only": {"lin_a": H, "lin_w": H},
        "fp16: attention operands only": {"attn": H},
        "fp16: all but the linear weights": {k: H for k in every if k != "lin_w"},
        "fp16: all but the conv weights": {k: H for k in every if k != "conv_w"},
    }
    err = {name: rel_l2(_RoundedOracle(sd, conf, cfg)(x, 481, ctx), ref) for name, cfg in rows.items()}
    for name, e in err.items():
        print(f"{name:48s} {e:.3e}")
    # the emulation lands where the B200 measurement does (smoke.unet_tiny: same network, inputs and timestep)
    limits = json.loads((Path(__file__).resolve().parent / "parity_limits.json").read_text())
    for op, name in (("fp16", "fp16: every operand"), ("bf16", "bf16: every operand")):
        measured = limits[op]["smoke.unet_tiny"]["measured"]
        assert 0.8 < measured / err[name] < 1.25, (op, measured, err[name])
    # sources add in quadrature: weights and activations contribute about equally, attention operands ~1 %
    w, a = err["fp16: weights only"], err["fp16: activations + attention operands only"]
    assert abs((w * w + a * a) ** 0.5 / err["fp16: every operand"] - 1) < 0.1
    assert 0.6 < w / a < 1.6
    assert err["fp16: attention operands only"] < 0.25 * err["fp16: every operand"]
    # a 16-bit operand path cannot reach 1e-3 on this network without exact weights somewhere
    assert err["fp16: every operand"] > 1e-3 > err["fp16: all but the linear weights"]
