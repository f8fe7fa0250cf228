"""ORACLE (test infrastructure) -- parity unpinned by the reference.

Restatement of diffusers 0.32.2 `DDIMScheduler` as configured by
cvssp/audioldm-s-full-v2/scheduler/scheduler_config.json (class pinned by
/root/reference/script/train/train_audioldm_lora.py:367; `add_noise` used at :504,
`set_timesteps`/`step` inside AudioLDMPipeline.__call__ -- app.py:14). SURVEY.md App. B.
Written op-by-op the way the scheduler does it (no fused coefficient form) so the
product's precomputed (c1, c2) table is checked against an independent formulation.
"""
from __future__ import annotations

import numpy as np
import torch

Tensor = torch.Tensor


class DDIMRef:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0195,
                 steps_offset=1, set_alpha_to_one=False):
        self.num_train_timesteps = num_train_timesteps
        # beta_schedule="scaled_linear"
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.steps_offset = steps_offset
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n: int):
        # timestep_spacing="leading"
        self.num_inference_steps = n
        step_ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.steps_offset
        self.timesteps = torch.from_numpy(ts)
        return self.timesteps

    def scale_model_input(self, sample: Tensor, t=None) -> Tensor:
        return sample

    def step(self, model_output: Tensor, t: int, sample: Tensor, eta: float = 0.0) -> Tensor:
        assert eta == 0.0, "oracle restates the eta=0 path the reference uses"
        t = int(t)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        beta_t = 1 - a_t
        pred_x0 = (sample - beta_t ** 0.5 * model_output) / a_t ** 0.5     # prediction_type="epsilon"
        pred_eps = model_output                                            # clip_sample=False
        direction = (1 - a_p) ** 0.5 * pred_eps                            # variance = 0 at eta=0
        return a_p ** 0.5 * pred_x0 + direction

    def add_noise(self, x0: Tensor, noise: Tensor, timesteps: Tensor) -> Tensor:
        ac = self.alphas_cumprod.to(x0.device, x0.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < x0.dim():
            sa = sa.unsqueeze(-1); sb = sb.unsqueeze(-1)
        return sa * x0 + sb * noise


class PNDMRef:
    """PLMS (skip_prk_steps=True) epsilon-history form named in BASELINE.json north_star.

    Restates diffusers `PNDMScheduler.step_plms` + `_get_prev_sample` with the same beta
    schedule / steps_offset as above (last bullet).
    """

    def __init__(self, **kw):
        self.base = DDIMRef(**kw)
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def set_timesteps(self, n: int):
        b = self.base
        b.num_inference_steps = n
        ratio = b.num_train_timesteps // n
        _ts = (np.arange(0, n) * ratio).round().astype(np.int64) + b.steps_offset
        # skip_prk_steps: plms_timesteps = [..., t1, t1, t0] reversed with the 2nd-to-last repeated
        plms = np.concatenate([_ts[:-1], _ts[-2:-1], _ts[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms)
        self.ets, self.counter, self.cur_sample = [], 0, None
        return self.timesteps

    def _prev_sample(self, sample, t, prev_t, eps):
        b = self.base
        a_t = b.alphas_cumprod[t]
        a_p = b.alphas_cumprod[prev_t] if prev_t >= 0 else b.final_alpha_cumprod
        beta_t, beta_p = 1 - a_t, 1 - a_p
        coeff = (a_p / a_t) ** 0.5
        denom = a_t * beta_p ** 0.5 + (a_t * beta_t * a_p) ** 0.5
        return coeff * sample - (a_p - a_t) * eps / denom

    def step(self, model_output, t, sample):
        b = self.base
        t = int(t)
        prev_t = t - b.num_train_timesteps // b.num_inference_steps
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_t = t
            t = t + b.num_train_timesteps // b.num_inference_steps
        if len(self.ets) == 1 and self.counter == 0:
            eps = model_output
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            eps = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            eps = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            eps = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            eps = (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4]) / 24
        out = self._prev_sample(sample, t, prev_t, eps)
        self.counter += 1
        return out
