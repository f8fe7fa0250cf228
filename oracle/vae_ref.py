"""ORACLE (test infrastructure) -- parity unpinned by the reference.

Functional torch-eager restatement of diffusers 0.32.2 `AutoencoderKL.decode` for the
cvssp/audioldm-s-full-v2 VAE config (loaded at
/root/reference/script/train/train_audioldm_lora.py:370; run inside
AudioLDMPipeline.__call__ -> decode_latents, app.py:14).  SURVEY.md App. E.
State dict uses diffusers key names (`post_quant_conv`, `decoder.*`).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

VAE_BLOCK_OUT = (128, 256, 512)
VAE_LATENT = 8
VAE_LAYERS = 2
VAE_GROUPS = 32
VAE_EPS = 1e-6
VAE_SCALING_FACTOR = 0.9227914214134216


def vae_decoder_param_shapes() -> Dict[str, Tuple[int, ...]]:
    P: Dict[str, Tuple[int, ...]] = {}

    def conv(n, ci, co, k):
        P[n + ".weight"] = (co, ci, k, k); P[n + ".bias"] = (co,)

    def norm(n, c):
        P[n + ".weight"] = (c,); P[n + ".bias"] = (c,)

    def lin(n, ci, co):
        P[n + ".weight"] = (co, ci); P[n + ".bias"] = (co,)

    def res(n, ci, co):
        norm(n + ".norm1", ci); conv(n + ".conv1", ci, co, 3)
        norm(n + ".norm2", co); conv(n + ".conv2", co, co, 3)
        if ci != co:
            conv(n + ".conv_shortcut", ci, co, 1)

    conv("post_quant_conv", VAE_LATENT, VAE_LATENT, 1)
    top = VAE_BLOCK_OUT[-1]
    conv("decoder.conv_in", VAE_LATENT, top, 3)
    res("decoder.mid_block.resnets.0", top, top)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", top)
    for p in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{p}", top, top)
    res("decoder.mid_block.resnets.1", top, top)
    prev = top
    for i, c in enumerate(reversed(VAE_BLOCK_OUT)):
        for j in range(VAE_LAYERS + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", prev, c)
            prev = c
        if i != len(VAE_BLOCK_OUT) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    norm("decoder.conv_norm_out", VAE_BLOCK_OUT[0])
    conv("decoder.conv_out", VAE_BLOCK_OUT[0], 1, 3)
    return P


def _res(sd, n, x):
    h = F.silu(F.group_norm(x, VAE_GROUPS, sd[n + ".norm1.weight"], sd[n + ".norm1.bias"], VAE_EPS))
    h = F.conv2d(h, sd[n + ".conv1.weight"], sd[n + ".conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, VAE_GROUPS, sd[n + ".norm2.weight"], sd[n + ".norm2.bias"], VAE_EPS))
    h = F.conv2d(h, sd[n + ".conv2.weight"], sd[n + ".conv2.bias"], padding=1)
    if n + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[n + ".conv_shortcut.weight"], sd[n + ".conv_shortcut.bias"])
    return x + h


def _mid_attn(sd, n, x):
    B, C, H, W = x.shape
    res = x
    h = F.group_norm(x, VAE_GROUPS, sd[n + ".group_norm.weight"], sd[n + ".group_norm.bias"], VAE_EPS)
    h = h.view(B, C, H * W).transpose(1, 2)
    q = F.linear(h, sd[n + ".to_q.weight"], sd[n + ".to_q.bias"])
    k = F.linear(h, sd[n + ".to_k.weight"], sd[n + ".to_k.bias"])
    v = F.linear(h, sd[n + ".to_v.weight"], sd[n + ".to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]   # single head, d = C
    o = F.linear(o, sd[n + ".to_out.0.weight"], sd[n + ".to_out.0.bias"])
    return o.transpose(1, 2).reshape(B, C, H, W) + res


def vae_decode(sd: Dict[str, Tensor], z: Tensor) -> Tensor:
    """z [B,8,H,16] (already divided by scaling_factor) -> mel [B,1,4H,64]."""
    h = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(h, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _res(sd, "decoder.mid_block.resnets.0", h)
    h = _mid_attn(sd, "decoder.mid_block.attentions.0", h)
    h = _res(sd, "decoder.mid_block.resnets.1", h)
    for i in range(len(VAE_BLOCK_OUT)):
        for j in range(VAE_LAYERS + 1):
            h = _res(sd, f"decoder.up_blocks.{i}.resnets.{j}", h)
        if i != len(VAE_BLOCK_OUT) - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            n = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            h = F.conv2d(h, sd[n + ".weight"], sd[n + ".bias"], padding=1)
    h = F.silu(F.group_norm(h, VAE_GROUPS, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], VAE_EPS))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


# ------------------------------------------------------------------------------------------ encoder
# `AutoencoderKL.encode` of the same VAE: the training loop's `vae.encode(batch["log_mel_spec"]).latent_dist.sample()`
# (/root/reference/script/train/train_audioldm_lora.py:495-496).  diffusers 0.32.2 Encoder: conv_in (1 -> 128), three
# DownEncoderBlock2D (two ResNets each; Downsample2D(padding=0) = F.pad (0,1,0,1) + conv k3 s2 between them), mid block
# (ResNet, single-head attention, ResNet), GroupNorm + SiLU, conv_out (512 -> 2 * latent), quant_conv 1x1, then
# DiagonalGaussianDistribution(mean, logvar clamped to [-30, 20]).
def vae_encoder_param_shapes() -> Dict[str, Tuple[int, ...]]:
    P: Dict[str, Tuple[int, ...]] = {}

    def conv(n, ci, co, k):
        P[n + ".weight"] = (co, ci, k, k); P[n + ".bias"] = (co,)

    def norm(n, c):
        P[n + ".weight"] = (c,); P[n + ".bias"] = (c,)

    def lin(n, ci, co):
        P[n + ".weight"] = (co, ci); P[n + ".bias"] = (co,)

    def res(n, ci, co):
        norm(n + ".norm1", ci); conv(n + ".conv1", ci, co, 3)
        norm(n + ".norm2", co); conv(n + ".conv2", co, co, 3)
        if ci != co:
            conv(n + ".conv_shortcut", ci, co, 1)

    conv("encoder.conv_in", 1, VAE_BLOCK_OUT[0], 3)
    prev = VAE_BLOCK_OUT[0]
    for i, c in enumerate(VAE_BLOCK_OUT):
        for j in range(VAE_LAYERS):
            res(f"encoder.down_blocks.{i}.resnets.{j}", prev, c)
            prev = c
        if i != len(VAE_BLOCK_OUT) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", c, c, 3)
    top = VAE_BLOCK_OUT[-1]
    res("encoder.mid_block.resnets.0", top, top)
    a = "encoder.mid_block.attentions.0"
    norm(a + ".group_norm", top)
    for p in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(f"{a}.{p}", top, top)
    res("encoder.mid_block.resnets.1", top, top)
    norm("encoder.conv_norm_out", top)
    conv("encoder.conv_out", top, 2 * VAE_LATENT, 3)
    conv("quant_conv", 2 * VAE_LATENT, 2 * VAE_LATENT, 1)
    return P


def vae_encode(sd: Dict[str, Tensor], mel: Tensor) -> Tuple[Tensor, Tensor]:
    """mel [B,1,T,64] -> (mean, logvar) of the latent posterior, each [B,8,T/4,16]; logvar clamped like diffusers."""
    h = F.conv2d(mel, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(len(VAE_BLOCK_OUT)):
        for j in range(VAE_LAYERS):
            h = _res(sd, f"encoder.down_blocks.{i}.resnets.{j}", h)
        if i != len(VAE_BLOCK_OUT) - 1:
            n = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[n + ".weight"], sd[n + ".bias"], stride=2)
    h = _res(sd, "encoder.mid_block.resnets.0", h)
    h = _mid_attn(sd, "encoder.mid_block.attentions.0", h)
    h = _res(sd, "encoder.mid_block.resnets.1", h)
    h = F.silu(F.group_norm(h, VAE_GROUPS, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], VAE_EPS))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    moments = F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])
    mean, logvar = moments.chunk(2, dim=1)
    return mean, logvar.clamp(-30.0, 20.0)
