"""Host side of the sampler step: DDIM / PNDM(PLMS) schedules folded into per-step coefficient rows.

`DDIMScheduler` mirrors the diffusers class the reference pins
(/root/reference/script/train/train_audioldm_lora.py:367; used through AudioLDMPipeline.__call__,
/root/reference/app.py:14) with the hub config of cvssp/audioldm-s-full-v2:
scaled_linear betas 0.0015..0.0195, 1000 train steps, leading spacing, steps_offset=1,
set_alpha_to_one=False, epsilon prediction, clip_sample=False.

The device kernel (csrc/sampler.cu) evaluates  x' = a * x_base + b * sum_i w_i e_i  from an 8-float row
per step, so the loop has no host sync and can sit in a CUDA graph.  Coefficients are derived in
float64 from the float32 alphas_cumprod table diffusers builds.
"""
from __future__ import annotations

import math
import struct
from typing import List, Optional

import numpy as np
import torch

Tensor = torch.Tensor


def _f_from_int(i: int) -> float:
    return struct.unpack("<f", struct.pack("<i", i))[0]


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0
    hist_slots = 0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0015, beta_end: float = 0.0195,
                 beta_schedule: str = "scaled_linear", steps_offset: int = 1, set_alpha_to_one: bool = False,
                 timestep_spacing: str = "leading", prediction_type: str = "epsilon", clip_sample: bool = False):
        if beta_schedule != "scaled_linear" or timestep_spacing != "leading" or prediction_type != "epsilon" or clip_sample:
            raise NotImplementedError("only the cvssp/audioldm-s-full-v2 scheduler config is on the reference path")
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps: Optional[int] = None
        self.timesteps: Optional[Tensor] = None

    # -- diffusers API ---------------------------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample: Tensor, timestep=None) -> Tensor:
        return sample

    def add_noise_coefficients(self, timesteps: Tensor):
        ac = self.alphas_cumprod[timesteps.cpu().long()]
        return ac.sqrt(), (1.0 - ac).sqrt()

    # -- device table ----------------------------------------------------------------------------
    def _alpha(self, t: int) -> float:
        return float(self.alphas_cumprod[t]) if t >= 0 else float(self.final_alpha_cumprod)

    def unet_timesteps(self) -> List[int]:
        return [int(t) for t in self.timesteps]

    def step_table(self, eta: float = 0.0) -> Tensor:
        """[num_steps, 8] fp32 rows for b200_sampler_step (DDIM, eta = 0)."""
        if eta != 0.0:
            raise NotImplementedError("eta != 0 (stochastic DDIM) is not on the reference path (pipeline default eta=0.0)")
        ratio = self.num_train_timesteps // self.num_inference_steps
        rows = []
        for t in self.unet_timesteps():
            a_t, a_p = self._alpha(t), self._alpha(t - ratio)
            a = math.sqrt(a_p / a_t)
            b = math.sqrt(1.0 - a_p) - math.sqrt(a_p * (1.0 - a_t) / a_t)
            rows.append([a, b, 1.0, 0.0, 0.0, 0.0, _f_from_int(0), _f_from_int(0)])
        return torch.tensor(rows, dtype=torch.float32)


class PNDMScheduler(DDIMScheduler):
    """PLMS (skip_prk_steps=True) form of diffusers' PNDMScheduler with the same beta schedule;
    named in BASELINE.json's north_star as the alternative latent update."""
    order = 1
    hist_slots = 4

    def set_timesteps(self, num_inference_steps: int, device=None) -> None:
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round().astype(np.int64) + self.steps_offset
        plms = np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms)

    def step_table(self, eta: float = 0.0) -> Tensor:
        ratio = self.num_train_timesteps // self.num_inference_steps
        rows, ets, counter = [], [], 0
        for t in self.unet_timesteps():
            prev_t, flags, push = t - ratio, 0, -1
            if counter != 1:
                ets = ets[-3:]
                push = next(s for s in range(4) if s not in ets)
                ets.append(push)
            else:
                prev_t, t = t, t + ratio
            hs = [0, 0, 0]
            if len(ets) == 1 and counter == 0:
                w = [1.0, 0.0, 0.0, 0.0]
                flags |= 2                      # remember x (cur_sample)
            elif len(ets) == 1 and counter == 1:
                w = [0.5, 0.5, 0.0, 0.0]
                hs[0] = ets[-1]
                flags |= 1                      # update from the remembered x
            elif len(ets) == 2:
                w = [1.5, -0.5, 0.0, 0.0]
                hs[0] = ets[-2]
            elif len(ets) == 3:
                w = [23 / 12, -16 / 12, 5 / 12, 0.0]
                hs[0], hs[1] = ets[-2], ets[-3]
            else:
                w = [55 / 24, -59 / 24, 37 / 24, -9 / 24]
                hs[0], hs[1], hs[2] = ets[-2], ets[-3], ets[-4]
            flags |= (push + 1) << 4
            a_t, a_p = self._alpha(t), self._alpha(prev_t)
            a = math.sqrt(a_p / a_t)
            b = -(a_p - a_t) / (a_t * math.sqrt(1.0 - a_p) + math.sqrt(a_t * (1.0 - a_t) * a_p))
            rows.append([a, b, *w, _f_from_int(flags), _f_from_int(hs[0] | hs[1] << 4 | hs[2] << 8)])
            counter += 1
        return torch.tensor(rows, dtype=torch.float32)
