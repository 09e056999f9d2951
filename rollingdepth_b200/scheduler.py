"""Sync-free DDIM schedule for the snippet denoise loops.

Semantics follow the reference's `DDIMScheduler` (diffusers/src/diffusers/schedulers/scheduling_ddim.py:
beta table :185-221, `set_timesteps` :297-340, `step` :342-468 with eta = 0, `add_noise` :471-495).
Differences in mechanism, not in results:
  * timesteps are python ints and the schedule scalars are looked up on the host, so `step` does not
    index a CPU table with a device tensor (the reference's hidden host sync at :405-406);
  * `step` / `add_noise` are one fused element-wise kernel each (C ABI rd_ddim_step / rd_ddim_add_noise)
    that applies the reference's op-by-op rounding in the sample dtype.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import ops


class DDIMSchedule:
    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", clip_sample: bool = False, set_alpha_to_one: bool = False,
                 steps_offset: int = 1, prediction_type: str = "v_prediction", timestep_spacing: str = "trailing",
                 clip_sample_range: float = 1.0, thresholding: bool = False, **_ignored):
        if thresholding:
            raise NotImplementedError("dynamic thresholding is not used by RollingDepth")
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} is not implemented")
        self.num_train_timesteps = num_train_timesteps
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)                      # fp32, host
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.steps_offset = steps_offset
        self.prediction_type = prediction_type
        self.timestep_spacing = timestep_spacing
        self.num_inference_steps: Optional[int] = None
        self.timesteps: List[int] = list(range(num_train_timesteps - 1, -1, -1))

    @classmethod
    def from_scheduler(cls, sch) -> "DDIMSchedule":
        """Build from a diffusers DDIMScheduler (or anything with a `.config` mapping)."""
        if isinstance(sch, cls):
            return sch
        cfg = dict(sch.config)
        keys = ("num_train_timesteps", "beta_start", "beta_end", "beta_schedule", "clip_sample", "set_alpha_to_one",
                "steps_offset", "prediction_type", "timestep_spacing", "clip_sample_range", "thresholding")
        return cls(**{k: cfg[k] for k in keys if k in cfg})

    @property
    def config(self):
        return dict(num_train_timesteps=self.num_train_timesteps, prediction_type=self.prediction_type,
                    timestep_spacing=self.timestep_spacing, steps_offset=self.steps_offset,
                    clip_sample=self.clip_sample)

    def set_timesteps(self, num_inference_steps: int, device=None) -> List[int]:
        T = self.num_train_timesteps
        if num_inference_steps > T:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than {T}")
        self.num_inference_steps = num_inference_steps
        if self.timestep_spacing == "linspace":
            ts = np.linspace(0, T - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif self.timestep_spacing == "leading":
            ratio = T // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ratio = T / num_inference_steps
            ts = np.round(np.arange(T, 0, -ratio)).astype(np.int64) - 1
        else:
            raise ValueError(f"{self.timestep_spacing} is not supported")
        self.timesteps = [int(t) for t in ts]
        return self.timesteps

    def _scalars(self, timestep: int):
        prev_t = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        # 0-dim fp32 tensor arithmetic, exactly as the reference evaluates it (:405-452)
        return float(a_t ** 0.5), float(b_t ** 0.5), float(a_prev ** 0.5), float((1 - a_prev - 0.0) ** 0.5)

    def step(self, model_output: torch.Tensor, timestep: int, sample: torch.Tensor,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' first")
        sa, sb, sap, sbp = self._scalars(int(timestep))
        clip = float(self.clip_sample_range) if self.clip_sample else 0.0
        return ops.ddim_step(model_output, sample, sa, sb, sap, sbp, self.prediction_type, clip, out=out)

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timestep: int) -> torch.Tensor:
        ac = self.alphas_cumprod.to(dtype=original_samples.dtype)          # :480 casts the table first
        sa = float(ac[int(timestep)] ** 0.5)
        sb = float((1 - ac[int(timestep)]) ** 0.5)
        return ops.ddim_add_noise(original_samples, noise.to(original_samples.dtype), sa, sb)
