"""TEST INFRASTRUCTURE ONLY — restatement of the third-party DDIM scheduler the reference calls.

`diffusers.DDIMScheduler` is a dependency absent from /root/reference (requirements.txt:1 `diffusers`, unpinned;
import paths imply 0.26-0.28).  Call sites: EMOAnimationPipeline.py:653-654 (set_timesteps), :367
(init_noise_sigma), :764 (scale_model_input, identity for DDIM), :817 (step, eta=0 from :554).  Constructor
arguments: configs/inference.yaml:23-26 (beta_start 0.00085, beta_end 0.012, beta_schedule "linear") with
steps_offset=1 and clip_sample=False forced by EMOAnimationPipeline.py:105-130.
Algorithm (Song et al. 2021, eq. 12, eta = 0; diffusers `DDIMScheduler.step` with prediction_type "epsilon",
timestep_spacing "leading", set_alpha_to_one=True):
    betas = linspace(beta_start, beta_end, 1000);  abar = cumprod(1 - betas)
    timesteps = (arange(n) * (1000 // n))[::-1] + steps_offset
    x0 = (x_t - sqrt(1 - abar_t) eps) / sqrt(abar_t);   x_prev = sqrt(abar_prev) x0 + sqrt(1 - abar_prev) eps
    with prev = t - 1000 // n and abar_prev = 1 when prev < 0.
PINNED on executed reference code: the reference carries the same algebra in-repo (`next_step`,
EMOAnimationPipeline.py:379-400 and magicanimate/utils/util.py:64-74).  oracle/ref_ddim.py cuts those functions out of
the reference sources with `ast` and runs them unchanged; tests/test_oracle.py holds `ddim_inversion_step` to their
output and `step` to the output of the SAME reference code run with the two alphas exchanged (golden vectors
tests/golden/ddim_reference_steps.pt + a live test).  What remains restated from the published diffusers definition
only: the timestep list ("leading" spacing + steps_offset) — consistent with the reference's own neighbour rule
`t - num_train_timesteps // num_inference_steps` (:391).
"""
from __future__ import annotations

import numpy as np


class DDIMOracle:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "linear", steps_offset: int = 1, set_alpha_to_one: bool = True):
        if beta_schedule == "linear":
            betas = np.linspace(beta_start, beta_end, num_train_timesteps, dtype=np.float32)
        elif beta_schedule == "scaled_linear":
            betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float32) ** 2
        else:
            raise ValueError(beta_schedule)
        self.alphas_cumprod = np.cumprod(1.0 - betas.astype(np.float32), dtype=np.float32)
        self.final_alpha_cumprod = np.float32(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n: int):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = (np.arange(0, n) * ratio).round()[::-1].astype(np.int64) + self.steps_offset
        return self.timesteps

    def scale_model_input(self, x, t=None):
        return x

    def alphas(self, t: int):
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return float(a_t), float(a_prev)

    def step(self, eps, t: int, x):
        a_t, a_prev = self.alphas(int(t))
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps

    def ddim_inversion_step(self, eps, t: int, x):
        """x_{t-ratio} -> x_t (restates EMOAnimationPipeline.next_step, :379-400)."""
        ratio = self.num_train_timesteps // self.num_inference_steps
        cur = min(t - ratio, 999)
        a_cur = float(self.alphas_cumprod[cur]) if cur >= 0 else float(self.final_alpha_cumprod)
        a_next = float(self.alphas_cumprod[t])
        x0 = (x - (1 - a_cur) ** 0.5 * eps) / a_cur ** 0.5
        return a_next ** 0.5 * x0 + (1 - a_next) ** 0.5 * eps


def cfg_combine(noise_pred_sum, counter, guidance_scale: float):
    """Window average + classifier-free guidance (EMOAnimationPipeline.py:812-814)."""
    avg = noise_pred_sum / counter
    half = avg.shape[0] // 2
    uncond, text = avg[:half], avg[half:]
    return uncond + guidance_scale * (text - uncond)


def uniform_windows(step: int, num_steps, num_frames: int, context_size: int, context_stride: int = 3,
                    context_overlap: int = 4, closed_loop: bool = True):
    """Sliding-window schedule (restates magicanimate/pipelines/context.py:12-42; pinned against it in tests)."""
    def ordered_halving(val: int) -> float:
        return int(f"{val:064b}"[::-1], 2) / (1 << 64)

    if num_frames <= context_size:
        return [list(range(num_frames))]
    out = []
    context_stride = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    for context_step in 1 << np.arange(context_stride):
        pad = int(round(num_frames * ordered_halving(step)))
        start = int(ordered_halving(step) * context_step) + pad
        stop = num_frames + pad + (0 if closed_loop else -context_overlap)
        for j in range(start, stop, int(context_size * context_step - context_overlap)):
            out.append([e % num_frames for e in range(j, j + int(context_size * context_step), int(context_step))])
    return out
