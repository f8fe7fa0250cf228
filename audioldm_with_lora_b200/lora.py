"""LoRA-loader API of the hot path: peft / diffusers state-dict formats -> engine adapters.

Mirrors what the reference does with third-party loaders:
  * `get_peft_model(unet, LoraConfig(r, lora_alpha, init_lora_weights="gaussian", target_modules))`
    -- /root/reference/script/train/train_audioldm_lora.py:378-385,
       /root/reference/script/inference/generate_audio.py:21-29
  * `load_file(model.safetensors)` + `load_state_dict(strict=False)` -- generate_audio.py:32-33
  * `get_peft_model_state_dict` / `convert_state_dict_to_diffusers` -- train_audioldm_lora.py:578
  * `pipe.unet.load_attn_procs(path)` -- /root/reference/app.py:11
Accepted key formats (all map to the same adapter):
  1. base_model.model.<path>.<t>.lora_A.<adapter>.weight / lora_B.<adapter>.weight   (accelerate save_state)
  2. base_model.model.<path>.<t>.lora_A.weight / lora_B.weight                       (get_peft_model_state_dict)
  3. [unet.]<path>.<t>.lora.down.weight / lora.up.weight                             (diffusers format)
LoRA stays unmerged at run time: y = base(x) + B(A(x)) * (alpha / r) * scale.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .arch import UNetConfig, attention_paths
from .engine import LoraEntry

Tensor = torch.Tensor

_PAT = re.compile(
    r"^(?:base_model\.model\.)?(?:unet\.)?(?P<path>.+?)\."
    r"(?:(?P<peft>lora_[AB])(?:\.(?P<adapter>[^.]+))?|lora\.(?P<dfs>down|up))\.weight$")


@dataclass
class LoraConfig:
    """peft.LoraConfig subset the reference uses (train_audioldm_lora.py:378-383)."""
    r: int = 8
    lora_alpha: float = 8
    init_lora_weights: object = "gaussian"
    target_modules: Sequence[str] = ("to_q", "to_k", "to_v", "to_out.0")
    lora_dropout: float = 0.0

    def __post_init__(self):
        if self.lora_dropout != 0.0:
            raise NotImplementedError("lora_dropout != 0 is not on the reference path (LoraConfig default 0.0)")


def target_linear_paths(cfg: UNetConfig, targets: Iterable[str]) -> List[str]:
    """peft rule: a module is adapted iff its dotted name ends with a target string."""
    out = []
    for p in attention_paths(cfg):
        for lin in ("to_q", "to_k", "to_v", "to_out.0"):
            name = f"{p}.{lin}"
            if any(name == t or name.endswith("." + t) for t in targets):
                out.append(name)
    return out


def init_adapters(cfg: UNetConfig, lcfg: LoraConfig, channels_of, seed: Optional[int] = None) -> Dict[str, LoraEntry]:
    """peft init: A ~ N(0, (1/r)^2) for "gaussian" (kaiming-uniform otherwise), B = 0."""
    g = torch.Generator().manual_seed(seed) if seed is not None else None
    out = {}
    for name in target_linear_paths(cfg, lcfg.target_modules):
        c = channels_of(name)
        if lcfg.init_lora_weights == "gaussian":
            A = torch.randn(lcfg.r, c, generator=g) / lcfg.r
        else:
            A = torch.empty(lcfg.r, c)
            torch.nn.init.kaiming_uniform_(A, a=5 ** 0.5, generator=g)
        out[name] = LoraEntry(A, torch.zeros(c, lcfg.r), float(lcfg.lora_alpha))
    return out


def parse_lora_state_dict(sd: Dict[str, Tensor], alpha: Optional[float] = None, adapter: Optional[str] = None,
                          network_alphas: Optional[Dict[str, float]] = None) -> Dict[str, LoraEntry]:
    """Collect (A, B) pairs from any of the three key formats; non-LoRA keys are ignored
    (`load_state_dict(strict=False)` semantics, generate_audio.py:33).  alpha defaults to r
    (scaling 1), which is what the reference trains with (lora_alpha=2, r=2)."""
    As: Dict[str, Tensor] = {}
    Bs: Dict[str, Tensor] = {}
    for k, v in sd.items():
        m = _PAT.match(k)
        if not m:
            continue
        if m.group("adapter") and adapter and m.group("adapter") != adapter:
            continue
        path = m.group("path")
        if path.endswith(".base_layer"):
            continue
        is_a = (m.group("peft") == "lora_A") or (m.group("dfs") == "down")
        (As if is_a else Bs)[path] = v.detach().float().cpu()
    out = {}
    for path, A in As.items():
        if path not in Bs:
            raise KeyError(f"LoRA state dict has lora_A/down for {path} but no lora_B/up")
        B = Bs[path]
        if A.dim() != 2 or B.dim() != 2 or B.shape[1] != A.shape[0]:
            raise ValueError(f"LoRA shapes for {path}: A {tuple(A.shape)} B {tuple(B.shape)}")
        a = alpha
        if network_alphas and path in network_alphas:
            a = network_alphas[path]
        out[path] = LoraEntry(A, B, float(a if a is not None else A.shape[0]))
    missing = set(Bs) - set(As)
    if missing:
        raise KeyError(f"LoRA state dict has lora_B/up without lora_A/down for {sorted(missing)[:3]}")
    return out


def to_peft_state_dict(adapters: Dict[str, LoraEntry], adapter_name: Optional[str] = None) -> Dict[str, Tensor]:
    """get_peft_model_state_dict layout (adapter_name=None) or the full-state layout (name given)."""
    mid = f".{adapter_name}" if adapter_name else ""
    out = {}
    for p, e in adapters.items():
        out[f"base_model.model.{p}.lora_A{mid}.weight"] = e.A.clone()
        out[f"base_model.model.{p}.lora_B{mid}.weight"] = e.B.clone()
    return out


def convert_state_dict_to_diffusers(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """peft keys -> diffusers `lora.down/up` keys (train_audioldm_lora.py:578)."""
    out = {}
    for k, v in sd.items():
        m = _PAT.match(k)
        if not m:
            out[k] = v
            continue
        is_a = (m.group("peft") == "lora_A") or (m.group("dfs") == "down")
        out[f"{m.group('path')}.lora.{'down' if is_a else 'up'}.weight"] = v
    return out
