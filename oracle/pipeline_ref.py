"""ORACLE (test infrastructure) -- parity unpinned by the reference.

Restatement of diffusers 0.32.2 `AudioLDMPipeline.__call__` on the `prompt_embeds=` /
`negative_prompt_embeds=` path (reference callers: /root/reference/app.py:14,
/root/reference/script/inference/generate_audio.py:47-52,
/root/reference/script/train/train_audioldm_lora.py:142,161).  The text encoder is
bypassed with synthetic L2-normalised 512-d embeddings, as BASELINE.json configures.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from .ddim_ref import DDIMRef
from .unet_ref import LoraSet, UNetSpec, unet_forward
from .vae_ref import VAE_SCALING_FACTOR, vae_decode

Tensor = torch.Tensor

VOCODER_UPSAMPLE_FACTOR = (5 * 4 * 2 * 2 * 2) / 16000.0     # prod(upsample_rates)/sampling_rate = 0.01
VAE_SCALE_FACTOR = 4                                        # 2 ** (len(vae.block_out_channels) - 1)
SAMPLING_RATE = 16000


def latent_height(audio_length_in_s: float) -> int:
    height = int(audio_length_in_s / VOCODER_UPSAMPLE_FACTOR)
    if height % VAE_SCALE_FACTOR != 0:
        height = int(np.ceil(height / VAE_SCALE_FACTOR)) * VAE_SCALE_FACTOR
    return height // VAE_SCALE_FACTOR


def denoise_loop(sd: Dict[str, Tensor], spec: UNetSpec, prompt_embeds: Tensor, negative_prompt_embeds: Tensor,
                 latents: Tensor, num_inference_steps: int, guidance_scale: float,
                 lora: Optional[LoraSet] = None, trace: Optional[List[Tensor]] = None,
                 callback: Optional[Callable] = None, eps_trace: Optional[List[Tensor]] = None) -> Tensor:
    """The hot loop of AudioLDMPipeline.__call__. latents [B,8,H,16] fp32."""
    sched = DDIMRef()
    timesteps = sched.set_timesteps(num_inference_steps)
    do_cfg = guidance_scale > 1.0
    embeds = torch.cat([negative_prompt_embeds, prompt_embeds]) if do_cfg else prompt_embeds
    latents = latents * sched.init_noise_sigma
    for i, t in enumerate(timesteps):
        x_in = torch.cat([latents] * 2) if do_cfg else latents
        x_in = sched.scale_model_input(x_in, t)
        eps = unet_forward(sd, spec, x_in, t, embeds, lora=lora)
        if do_cfg:
            e_u, e_t = eps.chunk(2)
            eps = e_u + guidance_scale * (e_t - e_u)
        if eps_trace is not None:
            eps_trace.append(eps.clone())
        latents = sched.step(eps, int(t), latents, eta=0.0)
        if trace is not None:
            trace.append(latents.clone())
        if callback is not None:
            callback(i, t, latents)
    return latents


def decode_tail(vae_sd: Dict[str, Tensor], vocoder, latents: Tensor, audio_length_in_s: float) -> np.ndarray:
    mel = vae_decode(vae_sd, latents / VAE_SCALING_FACTOR)
    if mel.dim() == 4:
        mel = mel.squeeze(1)
    wave = vocoder(mel).cpu().float()
    n = int(audio_length_in_s * SAMPLING_RATE)
    return wave[:, :n].numpy()
