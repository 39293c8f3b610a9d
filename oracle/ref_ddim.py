"""TEST INFRASTRUCTURE ONLY — executes the reference's OWN DDIM algebra without importing its (unimportable) modules.

`EMOAnimationPipeline.py` cannot be imported (it needs diffusers and a missing `animated_diff` module), but the DDIM
update it contains is a self-contained function: `EMOAnimationPipeline.next_step` (:379-400) and its twin
`magicanimate/utils/util.py:next_step` (:64-74).  This module cuts those two `def`s out of the reference sources with
`ast`, compiles them unchanged and runs them against a stub scheduler object, so the restatement in oracle/ddim.py is
pinned on executed reference code (golden vectors: tests/golden/ddim_reference_steps.pt, made by
`python -m oracle.make_golden ddim`; live check in tests/test_oracle.py when /root/reference is present).

The function computes  f(x, eps; a_from, a_to) = sqrt(a_to) (x - sqrt(1-a_from) eps)/sqrt(a_from) + sqrt(1-a_to) eps
with a_from = alphas_cumprod[t - ratio], a_to = alphas_cumprod[t] (inversion direction).  `diffusers.DDIMScheduler.step`
(eta = 0, epsilon prediction; third-party, absent) is the same f with the two alphas exchanged, so running the reference
function against a stub whose table has the two entries swapped (`swapped_alpha_stub`) executes the reference's code as
the forward sampler step x_t -> x_{t-ratio}.
"""
from __future__ import annotations

import ast
import types
from pathlib import Path
from typing import Union  # noqa: F401  (name used by the extracted source)

import numpy as np
import torch

REFERENCE = Path("/root/reference")


def _extract(path: Path, name: str, in_class: str | None = None):
    tree = ast.parse(path.read_text())
    scope = tree.body
    if in_class is not None:
        scope = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == in_class).body
    fn = next(n for n in scope if isinstance(n, ast.FunctionDef) and n.name == name)
    fn.decorator_list = []
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "np": np, "Union": Union}
    exec(compile(ast.fix_missing_locations(mod), str(path), "exec"), ns)
    return ns[name]


def reference_next_step_method():
    """EMOAnimationPipeline.next_step (:379-400) as a plain function(self, model_output, timestep, x, eta, verbose)"""
    return _extract(REFERENCE / "EMOAnimationPipeline.py", "next_step", in_class="EMOAnimationPipeline")


def reference_next_step_util():
    """magicanimate/utils/util.py:64-74 next_step(model_output, timestep, sample, ddim_scheduler)"""
    return _extract(REFERENCE / "magicanimate" / "utils" / "util.py", "next_step")


def scheduler_stub(alphas_cumprod, num_inference_steps: int, num_train_timesteps: int = 1000, final_alpha_cumprod=1.0):
    """the attributes the reference function reads from `self.scheduler` / `ddim_scheduler`"""
    s = types.SimpleNamespace()
    s.config = types.SimpleNamespace(num_train_timesteps=num_train_timesteps)
    s.num_inference_steps = num_inference_steps
    s.alphas_cumprod = torch.as_tensor(np.asarray(alphas_cumprod), dtype=torch.float32)
    s.final_alpha_cumprod = torch.tensor(float(final_alpha_cumprod))
    return s


def swapped_alpha_stub(alphas_cumprod, t: int, num_inference_steps: int, num_train_timesteps: int = 1000,
                       final_alpha_cumprod=1.0):
    """Stub whose table entries for t and t - ratio are exchanged: the reference's next_step(eps, t, x) then evaluates
    the forward DDIM step x_t -> x_{t-ratio} (a_from = abar_t, a_to = abar_{t-ratio}, or 1 past the end)."""
    a = np.array(alphas_cumprod, dtype=np.float32).copy()
    ratio = num_train_timesteps // num_inference_steps
    prev = t - ratio
    a_t = a[t]
    a_prev = a[prev] if prev >= 0 else np.float32(final_alpha_cumprod)
    a[t] = a_prev
    final = final_alpha_cumprod
    if prev >= 0:
        a[prev] = a_t
    else:
        final = a_t          # the function reads final_alpha_cumprod when t - ratio < 0
    return scheduler_stub(a, num_inference_steps, num_train_timesteps, final)
