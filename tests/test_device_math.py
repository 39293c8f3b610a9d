"""Accuracy claims of the polynomial device functions, checked on the CPU by re-running the SAME fp32 arithmetic in
numpy with the coefficients parsed out of the CUDA sources (so a typo in a constant fails here, without a GPU):
  common.cuh:gelu_sig          exact GELU in logistic form (GEGLU epilogue),  claimed |abs err| <= 2.6e-5
  attention_tc.cuh:exp2_poly2   2^x on the FMA pipe (softmax of the tcgen05 attention), claimed rel err <= 7.5e-5 (+ fp32)
The reference arithmetic they stand in for: F.gelu (erf form, orig_attention.py:825-827) and softmax's exp
(orig_attention.py:671)."""
import re
from pathlib import Path

import numpy as np
from scipy.special import ndtr

CSRC = Path(__file__).resolve().parent.parent / "emote_hack_b200" / "csrc"
F32 = np.float32


def _floats(src: str, start: str, end: str):
    body = src[src.index(start):]
    body = body[:body.index(end)]
    return [float(x) for x in re.findall(r"(-?\d+\.\d*(?:e-?\d+)?)f", body)]


def test_gelu_sig_polynomial():
    src = (CSRC / "common.cuh").read_text()
    c = _floats(src, "__device__ __forceinline__ float gelu_sig", "return x * r;")
    # order of appearance: clamp lo, clamp hi, c2, c1, c0, -2*log2(e), 1.0
    lo, hi, c2, c1, c0, k, one = c
    assert (lo, hi, one) == (-5.5, 5.5, 1.0) and abs(k + 2 * 1.4426950408889634) < 1e-6
    x = np.linspace(-12, 12, 480001).astype(F32)
    xc = np.clip(x, F32(lo), F32(hi))
    x2 = (xc * xc).astype(F32)
    q = (F32(c2) * x2 + F32(c1)).astype(F32)
    q = (q * x2 + F32(c0)).astype(F32)
    e = np.exp2((q * (xc * F32(k)).astype(F32)).astype(F32).astype(np.float64))
    got = x.astype(np.float64) / (e + 1.0)
    ref = x.astype(np.float64) * ndtr(x.astype(np.float64))
    err = np.abs(got - ref)
    assert err.max() < 2.7e-5, err.max()
    # relative accuracy survives in the negative tail (no 1 + tanh cancellation), monotone saturation outside the clamp
    tail = (x < -3) & (x > -5)
    assert (err[tail] / np.abs(ref[tail])).max() < 5e-2
    assert np.all(got[x > 6] == x[x > 6].astype(np.float64) / (np.exp2((q * (xc * F32(k)))[x > 6].astype(np.float64)) + 1.0))
    assert np.abs(got[x > 8] - x[x > 8]).max() < 1e-6 and np.abs(got[x < -8]).max() < 1e-6


def test_exp2_poly2_cody_waite():
    src = (CSRC / "attention_tc.cuh").read_text()
    c = _floats(src, "__device__ __forceinline__ void exp2_poly2", "r0 = __int_as_float")
    # order: MAGIC, clamp (-126 twice), 1.0, MAGIC ops (1.0, -1.0), c3, c2, c1, c0
    magic = c[0]
    assert magic == 12582912.0 and c.count(-126.0) == 2
    c3, c2, c1, c0 = c[-4:]
    x = np.concatenate([np.linspace(-150, 0, 600001), np.linspace(0, 20, 20001)]).astype(F32)
    xc = np.maximum(x, F32(-126.0))
    xf = (xc + F32(magic)).astype(F32)                 # round-to-nearest integer lands in the low mantissa bits
    n = (xf - F32(magic)).astype(F32)
    f = (xc - n).astype(F32)
    assert np.abs(f).max() <= 0.5
    p = (F32(c3) * f + F32(c2)).astype(F32)
    p = (p * f + F32(c1)).astype(F32)
    p = (p * f + F32(c0)).astype(F32)
    bits = (p.view(np.int32).astype(np.int64) + (xf.view(np.int32).astype(np.int64) << 23)) & 0xFFFFFFFF
    got = bits.astype(np.uint32).view(F32).astype(np.float64)
    ref = np.exp2(xc.astype(np.float64))
    rel = np.abs(got / ref - 1.0)
    ok = x >= -120
    assert rel[ok].max() < 7.6e-5, rel[ok].max()
    # at the clamp (masked keys arrive as -inf) the spliced exponent reaches 0: the value is a harmless ~1e-38, never
    # negative, NaN or large
    assert np.all(np.isfinite(got)) and got.min() >= 0.0 and got[x < -125].max() < 1e-37
