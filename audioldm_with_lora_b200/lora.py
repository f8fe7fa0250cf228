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

Checkpoint writers and the opt-in pre-merge:
  * `save_lora_checkpoint(adapters, dir, fmt)` writes what the reference's tooling writes -- accelerate's
    `save_state` file `model.safetensors` with full peft keys (train_audioldm_lora.py:575, read back by
    generate_audio.py:32), the `get_peft_model_state_dict` layout, or diffusers' `pytorch_lora_weights.safetensors`
    (`lora.down/up` keys, train_audioldm_lora.py:577-579 / app.py:11);
  * `merge_lora_into_state_dict(sd, adapters, scale)` returns base weights with W' = W + (alpha / r) * scale * B A
    folded in (peft `merge_and_unload`).  Never applied implicitly: the merged path rounds W' to bf16 once, the
    unmerged path rounds W and the rank-r term separately, so the two are not bit-identical.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from pathlib import Path
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


# ----------------------------------------------------------------------------- checkpoint writers, opt-in merge
_CKPT_FILES = {"peft_full": "model.safetensors", "peft": "adapter_model.safetensors",
               "diffusers": "pytorch_lora_weights.safetensors"}


def lora_state_dict_as(adapters: Dict[str, LoraEntry], fmt: str = "peft", adapter_name: str = "default") -> Dict[str, Tensor]:
    """The adapters under the keys of one of the three formats `parse_lora_state_dict` reads:
    "peft_full" (accelerate save_state: `...lora_A.<adapter>.weight`), "peft" (`get_peft_model_state_dict`),
    "diffusers" (`<path>.lora.down/up.weight`)."""
    if fmt == "peft_full":
        return to_peft_state_dict(adapters, adapter_name)
    if fmt == "peft":
        return to_peft_state_dict(adapters, None)
    if fmt == "diffusers":
        return convert_state_dict_to_diffusers(to_peft_state_dict(adapters, None))
    raise ValueError(f"unknown LoRA checkpoint format {fmt!r} (peft_full | peft | diffusers)")


def save_lora_checkpoint(adapters: Dict[str, LoraEntry], save_directory, fmt: str = "diffusers",
                         adapter_name: str = "default", dtype: torch.dtype = torch.float32) -> Path:
    """Write the adapters as a safetensors file named as the reference's tools name it (see the module docstring);
    alpha / rank go into the file metadata (peft keeps them in adapter_config.json, diffusers in network_alphas).
    Returns the file path; `UNet2DConditionModel.load_attn_procs(save_directory)` reads it back."""
    from safetensors.torch import save_file
    sd = {k: v.detach().to(dtype).contiguous().cpu() for k, v in lora_state_dict_as(adapters, fmt, adapter_name).items()}
    out = Path(save_directory)
    out.mkdir(parents=True, exist_ok=True)
    ranks = sorted({int(e.A.shape[0]) for e in adapters.values()})
    alphas = sorted({float(e.alpha) for e in adapters.values()})
    meta = {"format": fmt, "r": ",".join(map(str, ranks)), "lora_alpha": ",".join(map(str, alphas))}
    path = out / _CKPT_FILES[fmt]
    save_file(sd, str(path), metadata=meta)
    return path


def load_lora_checkpoint(path, alpha: Optional[float] = None) -> Dict[str, LoraEntry]:
    """Read any of the files `save_lora_checkpoint` (or the reference's tooling) writes.  alpha: explicit value, else
    the file metadata when it names a single value, else rank (scaling 1, what the reference trains with)."""
    from safetensors import safe_open
    p = Path(path)
    if p.is_dir():
        cands = [p / f for f in ("pytorch_lora_weights.safetensors", "model.safetensors", "adapter_model.safetensors")]
        p = next((c for c in cands if c.exists()), cands[0])
    sd = {}
    with safe_open(str(p), framework="pt") as f:
        meta = f.metadata() or {}
        for k in f.keys():
            sd[k] = f.get_tensor(k)
    if alpha is None and meta.get("lora_alpha") and "," not in meta["lora_alpha"]:
        alpha = float(meta["lora_alpha"])
    return parse_lora_state_dict(sd, alpha)


def merge_lora_into_state_dict(sd: Dict[str, Tensor], adapters: Dict[str, LoraEntry], scale: float = 1.0,
                               sign: float = 1.0) -> Dict[str, Tensor]:
    """Copy of the base state dict with W' = W + sign * (alpha / r) * scale * B @ A for every adapted projection
    (peft `merge_and_unload`; sign = -1 un-merges).  fp32 arithmetic; keys without an adapter are shared, not copied."""
    out = dict(sd)
    for path, e in adapters.items():
        key = path + ".weight"
        if key not in sd:
            raise KeyError(f"adapter {path} has no base weight {key} in the state dict")
        w = sd[key].detach().float()
        if w.shape != (e.B.shape[0], e.A.shape[1]):
            raise ValueError(f"{key}: base {tuple(w.shape)} vs B A {(e.B.shape[0], e.A.shape[1])}")
        s_eff = sign * scale * e.alpha / e.A.shape[0]
        out[key] = (w + s_eff * (e.B.float() @ e.A.float())).to(sd[key].dtype)
    return out
