"""Tail of AudioLDMPipeline.__call__: VAE decode + SpeechT5-HiFi-GAN vocoder.

BASELINE.json's north_star keeps these two models "on the reference path" (torch eager: cuDNN /
cuBLAS), timed end-to-end only; they are NOT part of the hand-written hot path (SURVEY.md 8a row
a12, 8f items 1-2).  `AutoencoderKLDecoder` is an nn.Module with diffusers' state-dict key names
(`post_quant_conv`, `decoder.*`) for the cvssp/audioldm-s-full-v2 VAE config (loaded by the
reference at /root/reference/script/train/train_audioldm_lora.py:370); the vocoder is transformers'
own `SpeechT5HifiGan` (train_audioldm_lora.py:371), random-init when no checkpoint is available.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

Tensor = torch.Tensor


class _Res(nn.Module):
    def __init__(self, cin: int, cout: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class _MidAttn(nn.Module):
    def __init__(self, c: int, groups: int, eps: float):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        o = F.scaled_dot_product_attention(self.to_q(t)[:, None], self.to_k(t)[:, None], self.to_v(t)[:, None])[:, 0]
        return self.to_out[0](o).transpose(1, 2).reshape(b, c, h, w) + x


class _Up(nn.Module):
    def __init__(self, cin, cout, layers, groups, eps, upsample):
        super().__init__()
        self.resnets = nn.ModuleList([_Res(cin if j == 0 else cout, cout, groups, eps) for j in range(layers)])
        self.upsamplers = nn.ModuleList([nn.ModuleDict({"conv": nn.Conv2d(cout, cout, 3, padding=1)})]) if upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0]["conv"](F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class _Mid(nn.Module):
    def __init__(self, c, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList([_Res(c, c, groups, eps), _Res(c, c, groups, eps)])
        self.attentions = nn.ModuleList([_MidAttn(c, groups, eps)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Decoder(nn.Module):
    def __init__(self, latent, block_out, layers, groups, eps):
        super().__init__()
        top = block_out[-1]
        self.conv_in = nn.Conv2d(latent, top, 3, padding=1)
        self.mid_block = _Mid(top, groups, eps)
        ups, prev = [], top
        rev = list(reversed(block_out))
        for i, c in enumerate(rev):
            ups.append(_Up(prev, c, layers + 1, groups, eps, upsample=i != len(rev) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(groups, block_out[0], eps=eps)
        self.conv_out = nn.Conv2d(block_out[0], 1, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z))
        for u in self.up_blocks:
            h = u(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKLDecoder(nn.Module):
    """`AutoencoderKL.decode` half of the AudioLDM VAE."""

    class _Cfg:
        scaling_factor = 0.9227914214134216
        block_out_channels = (128, 256, 512)
        latent_channels = 8

    def __init__(self, block_out=(128, 256, 512), latent=8, layers=2, groups=32, eps=1e-6):
        super().__init__()
        self.config = self._Cfg()
        self.config.block_out_channels = tuple(block_out)
        self.post_quant_conv = nn.Conv2d(latent, latent, 1)
        self.decoder = _Decoder(latent, block_out, layers, groups, eps)

    @torch.no_grad()
    def decode(self, z: Tensor) -> Tensor:
        return self.decoder(self.post_quant_conv(z))


def build_vocoder(seed: Optional[int] = 0):
    """Random-init transformers SpeechT5HifiGan with the AudioLDM vocoder config."""
    from transformers import SpeechT5HifiGan, SpeechT5HifiGanConfig
    cfg = SpeechT5HifiGanConfig(model_in_dim=64, sampling_rate=16000, upsample_initial_channel=1024,
                                upsample_rates=[5, 4, 2, 2, 2], upsample_kernel_sizes=[16, 16, 8, 4, 4],
                                resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3,
                                normalize_before=False)
    if seed is not None:
        torch.manual_seed(seed)
    return SpeechT5HifiGan(cfg).eval()


def random_vae_decoder(seed: int = 0, std: float = 0.02) -> AutoencoderKLDecoder:
    from .synthetic import random_state_dict_from_shapes
    vae = AutoencoderKLDecoder().eval()
    shapes = {k: tuple(v.shape) for k, v in vae.state_dict().items()}
    vae.load_state_dict(random_state_dict_from_shapes(shapes, seed, std))
    return vae
