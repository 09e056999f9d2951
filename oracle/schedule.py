"""Oracle for the snippet scheduler and the DDIM algebra (TEST INFRASTRUCTURE).

Restates rollingdepth/rollingdepth_pipeline.py:465-515 (window index lists,
dilation cap) and diffusers/src/diffusers/schedulers/scheduling_ddim.py:185-495
(beta table, timestep spacing, `step`, `add_noise`).  Pinned by
tests/golden/snippets.json (reference outputs) and by the reference's own DDIM
known-answer tests (diffusers/tests/schedulers/test_scheduler_ddim.py:114-176).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


# --------------------------------------------------------------------------- snippets
def get_snippet_indice(i_step, total_step, seq_len, snippet_len, dilation_start, dilation_end, stride):
    """rollingdepth_pipeline.py:465-502.  `total_step` = len(timesteps)."""
    gap_start = dilation_start - 1
    gap_end = dilation_end - 1
    assert gap_start >= gap_end and gap_end >= 0
    # float-rounding sensitive: keep the reference's expression shape
    gap_cur = int((1 - (i_step) / total_step) * (gap_start - gap_end) + gap_end)
    win = (snippet_len - 1) * (gap_cur + 1) + 1
    starts = list(range(0, seq_len - win + 1, stride))
    if starts[-1] < seq_len - win:
        starts.append(seq_len - win)
    return [list(range(s0, s0 + win, gap_cur + 1)) for s0 in starts]


def cap_max_dilation(seq_len: int, snippet_len: int, dilation: int) -> int:
    """rollingdepth_pipeline.py:504-515."""
    max_allowed_gap = int(seq_len / snippet_len) - 1
    return min(max_allowed_gap, dilation) if max_allowed_gap < dilation else dilation


# --------------------------------------------------------------------------- DDIM
class DDIMRef:
    """scheduling_ddim.py -- only the members the pipeline touches."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
                 steps_offset=1, prediction_type="v_prediction", timestep_spacing="trailing",
                 clip_sample_range=1.0):
        self.T = num_train_timesteps
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)               # :214-215
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.steps_offset, self.prediction_type, self.spacing = steps_offset, prediction_type, timestep_spacing
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, n: int):
        """:297-340"""
        self.num_inference_steps = n
        if self.spacing == "linspace":
            ts = np.linspace(0, self.T - 1, n).round()[::-1].copy().astype(np.int64)
        elif self.spacing == "leading":
            ratio = self.T // n
            ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.spacing == "trailing":
            ratio = self.T / n
            ts = np.round(np.arange(self.T, 0, -ratio)).astype(np.int64) - 1
        else:
            raise ValueError(self.spacing)
        self.timesteps = torch.from_numpy(ts)
        return self.timesteps

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor) -> torch.Tensor:
        """:342-468 with eta = 0, no thresholding.  Same op order => same roundings."""
        timestep = int(timestep)
        prev_t = timestep - self.T // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        elif self.prediction_type == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            raise ValueError(self.prediction_type)
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        direction = (1 - a_prev - 0.0) ** 0.5 * eps
        return a_prev ** 0.5 * x0 + direction

    def add_noise(self, original: torch.Tensor, noise: torch.Tensor, timestep: int) -> torch.Tensor:
        """:471-495 (alphas cast to the sample dtype first, as the reference does)."""
        ac = self.alphas_cumprod.to(dtype=original.dtype)
        sa = ac[int(timestep)] ** 0.5
        sb = (1 - ac[int(timestep)]) ** 0.5
        return sa * original + sb * noise
