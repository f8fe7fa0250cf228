"""Deterministic synthetic weights / inputs (there is no network for checkpoints).

Random-init weights of the `cvssp/audioldm-s-full-v2` architecture, seeded per key so that
any subset can be regenerated independently, plus the synthetic CLAP embeddings and
initial latents BASELINE.json's configs name (SURVEY.md 8d "Synthetic inputs").
Weights ~ N(0, 0.02^2); biases ~ N(0, 0.02^2) and norm affine (1 + 0.1 N, 0.1 N) so every
bias / affine path is exercised (the survey's all-zero biases would leave them untested).
LoRA: A ~ N(0, (1/r)^2) (peft "gaussian"), B ~ N(0, 0.02^2) -- NOT zero, otherwise the
LoRA branch is untested.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Iterable, Sequence, Tuple

import torch
import torch.nn.functional as F

from .arch import UNetConfig, attention_paths, unet_param_shapes

Tensor = torch.Tensor


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def _randn(seed: int, key: str, shape: Sequence[int]) -> Tensor:
    g = torch.Generator().manual_seed(_key_seed(seed, key))
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32)


def random_state_dict_from_shapes(shapes: Dict[str, Tuple[int, ...]], seed: int = 0, std: float = 0.02) -> Dict[str, Tensor]:
    sd = {}
    for k, shp in shapes.items():
        is_norm = ".norm" in k or k.startswith("conv_norm_out") or "group_norm" in k or "conv_norm_out" in k
        if is_norm and k.endswith(".weight"):
            sd[k] = 1.0 + 0.1 * _randn(seed, k, shp)
        elif is_norm and k.endswith(".bias"):
            sd[k] = 0.1 * _randn(seed, k, shp)
        else:
            sd[k] = std * _randn(seed, k, shp)
    return sd


def random_unet_state_dict(cfg: UNetConfig, seed: int = 0) -> Dict[str, Tensor]:
    return random_state_dict_from_shapes(unet_param_shapes(cfg), seed)


def random_lora_state_dict(cfg: UNetConfig, r: int, targets: Iterable[str] = ("to_q", "to_k", "to_v", "to_out.0"),
                           seed: int = 100, fmt: str = "peft", b_std: float = 0.02) -> Dict[str, Tensor]:
    """LoRA weights in one of the three key formats of SURVEY.md App. C.

    fmt = "peft"      : base_model.model.<path>.<t>.lora_A.default.weight   (accelerate save_state)
          "peft_sd"   : base_model.model.<path>.<t>.lora_A.weight           (get_peft_model_state_dict)
          "diffusers" : <path>.<t>.lora.down.weight / lora.up.weight         (convert_state_dict_to_diffusers)
    """
    sd = {}
    boc = cfg.block_out_channels
    for path in attention_paths(cfg):
        c = _path_channels(path, cfg)
        for t in targets:
            A = _randn(seed, f"{path}.{t}.A", (r, c)) / r
            B = _randn(seed, f"{path}.{t}.B", (c, r)) * b_std
            if fmt == "peft":
                sd[f"base_model.model.{path}.{t}.lora_A.default.weight"] = A
                sd[f"base_model.model.{path}.{t}.lora_B.default.weight"] = B
            elif fmt == "peft_sd":
                sd[f"base_model.model.{path}.{t}.lora_A.weight"] = A
                sd[f"base_model.model.{path}.{t}.lora_B.weight"] = B
            elif fmt == "diffusers":
                sd[f"{path}.{t}.lora.down.weight"] = A
                sd[f"{path}.{t}.lora.up.weight"] = B
            else:
                raise ValueError(fmt)
    return sd


def _path_channels(path: str, cfg: UNetConfig) -> int:
    p = path.split(".")
    boc = cfg.block_out_channels
    if p[0] == "mid_block":
        return boc[-1]
    i = int(p[1])
    return boc[i] if p[0] == "down_blocks" else boc[len(boc) - 1 - i]


def clap_embeddings(batch: int, seed: int = 1, neg_seed: int = 2) -> Tuple[Tensor, Tensor]:
    """(prompt_embeds [B,512], negative_prompt_embeds [B,512]), L2-normalised like
    /root/reference/script/train/train_audioldm_lora.py:524."""
    g = torch.Generator().manual_seed(seed)
    pos = F.normalize(torch.randn(batch, 512, generator=g), dim=-1)
    g2 = torch.Generator().manual_seed(neg_seed)
    neg = F.normalize(torch.randn(1, 512, generator=g2), dim=-1).expand(batch, 512).contiguous()
    return pos, neg


def initial_latents(batch: int, height: int, seed: int = 3, first_index: int = 0) -> Tensor:
    """randn [B,8,H,16]; one generator per global prompt index so shards agree with the whole."""
    out = []
    for i in range(batch):
        g = torch.Generator().manual_seed(seed + first_index + i)
        out.append(torch.randn(1, 8, height, 16, generator=g))
    return torch.cat(out)
