"""Drop-in boundary: diffusers-shaped UNet / Attention / attention-processor / LoRA-loader API.

diffusers and peft are not installable here, so this module mirrors the slice of their interface
the reference touches (same names, argument meaning and error behaviour):

  * `UNet2DConditionModel.forward(sample, timestep, encoder_hidden_states=None, class_labels=...,
     cross_attention_kwargs={"scale": s}, return_dict=...)`
        -- /root/reference/script/train/train_audioldm_lora.py:539-546
  * `unet.attn_processors` / `unet.set_attn_processor(proc | {name: proc})`, processor signature
     `__call__(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, **kw)`
        -- diffusers attention-processor API (AttnProcessor2_0 is what the reference runs)
  * `unet.add_adapter(LoraConfig)`, `get_peft_model(unet, LoraConfig)`, `load_state_dict(sd, strict=False)`
     with peft keys, `unet.load_attn_procs(path | dict)`, `get_peft_model_state_dict`
        -- train_audioldm_lora.py:378-387, generate_audio.py:21-39, app.py:11

With the default `B200AttnProcessor` on every Attention module the whole forward runs in the fused
engine (engine.py).  Any other processor installed on a module is honoured: the engine hands that
module's LayerNorm output to `processor(attn, hidden_states)` as a torch tensor and continues with
its result -- so the seam is the real diffusers seam, not a facade.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Union

import torch
import torch.nn as nn

from . import ops, packing
from .arch import CONFIGS, UNetConfig, attention_paths, unet_param_shapes
from .engine import LoraEntry, UNetEngine
from .lora import (LoraConfig, convert_state_dict_to_diffusers, init_adapters, parse_lora_state_dict,
                   to_peft_state_dict)

Tensor = torch.Tensor


@dataclass
class UNet2DConditionOutput:
    sample: Tensor


class LoraLinear(nn.Module):
    """peft.tuners.lora.Linear shaped view of one adapted projection (base_layer + lora_A/B)."""

    def __init__(self, base: nn.Linear):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict()
        self.lora_B = nn.ModuleDict()
        self.scaling: Dict[str, float] = {}
        self.r: Dict[str, int] = {}

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias

    def update_layer(self, name: str, A: Tensor, B: Tensor, alpha: float) -> None:
        r = A.shape[0]
        la, lb = nn.Linear(A.shape[1], r, bias=False), nn.Linear(r, B.shape[0], bias=False)
        la.weight.data.copy_(A); lb.weight.data.copy_(B)
        self.lora_A[name], self.lora_B[name] = la, lb
        self.scaling[name], self.r[name] = alpha / r, r


class Attention(nn.Module):
    """diffusers.models.attention_processor.Attention attribute surface (self-attention use)."""

    def __init__(self, c: int, heads: int, name: str):
        super().__init__()
        self.heads = heads
        self.inner_dim = c
        self.scale = (c // heads) ** -0.5
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.group_norm = self.spatial_norm = self.norm_cross = None
        self.to_q, self.to_k, self.to_v = (nn.Linear(c, c, bias=False) for _ in range(3))
        self.to_out = nn.ModuleList([nn.Linear(c, c, bias=True), nn.Dropout(0.0)])
        self.processor = B200AttnProcessor()
        self.b200_name = name

    def set_processor(self, processor) -> None:
        self.processor = processor
        # a foreign (torch) processor reads this module's weights: keep them where the activations are
        dev = getattr(self, "b200_device", None)
        if dev is not None and dev.type == "cuda" and not isinstance(processor, B200AttnProcessor):
            self.to(dev)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


def _proj_parts(lin):
    """(weight, bias, LoraEntry | None, signature) of an nn.Linear or a LoraLinear (adapter 'default')."""
    if isinstance(lin, LoraLinear) or hasattr(lin, "base_layer"):
        base = lin.base_layer
        ent, sig = None, ()
        if len(lin.lora_A):
            name = "default" if "default" in lin.lora_A else next(iter(lin.lora_A))
            wa, wb = lin.lora_A[name].weight, lin.lora_B[name].weight
            A, B = wa.detach(), wb.detach()
            ent = LoraEntry(A.float().cpu(), B.float().cpu(), lin.scaling[name] * A.shape[0])
            sig = (wa.data_ptr(), wa._version, wb.data_ptr(), wb._version, lin.scaling[name])
        return (base.weight.detach(), None if base.bias is None else base.bias.detach(), ent,
                (base.weight.data_ptr(), base.weight._version) + sig)
    return (lin.weight.detach(), None if lin.bias is None else lin.bias.detach(), None,
            (lin.weight.data_ptr(), lin.weight._version))


class B200AttnProcessor:
    """Attention processor running on the sm_100a kernels: fused QKV GEMM with the LoRA branch in
    the epilogue accumulator, flash-style attention, output projection (+bias, +LoRA).

    Drop-in for diffusers' AttnProcessor2_0 on a self-attention `Attention` module: call signature
    `(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0)`,
    input/output `[B, S, C]`, same dtype out as in.
    """

    def __init__(self):
        self._cache: Dict[int, tuple] = {}

    def _packed(self, attn, m: int, scale: float, device):
        parts = [_proj_parts(getattr(attn, n)) for n in ("to_q", "to_k", "to_v")] + [_proj_parts(attn.to_out[0])]
        key = (m, float(scale), tuple(p[3] for p in parts))
        hit = self._cache.get(id(attn))
        if hit is not None and hit[0] == key:
            return hit[1]
        c = parts[0][0].shape[1]
        mt = -(-m // 128)
        W = {}
        ents = [p[2] for p in parts[:3]]
        wqkv = torch.cat([p[0].float().cpu() for p in parts[:3]])
        if any(p[1] is not None for p in parts[:3]):
            raise NotImplementedError("B200AttnProcessor: to_q/to_k/to_v with bias are not on the reference path")
        if any(e is not None for e in ents):
            a_list = [e.A if e else None for e in ents]
            seg = packing.lora_up_segment([e.B if e else None for e in ents], a_list,
                                          [e.scaling * scale if e else 0.0 for e in ents], c)
            W["down_qkv"] = packing.pack_lora_down(a_list, c, device=device)
            W["qkv"] = packing.pack([wqkv, seg], None, ops.choose_block_n(3 * c, mt), 1, c, seg.shape[1], device=device)
        else:
            W["qkv"] = packing.pack([wqkv], None, ops.choose_block_n(3 * c, mt), 1, c, device=device)
        wo, bo, eo = parts[3][:3]
        if eo is not None:
            seg = packing.lora_up_segment([eo.B], [eo.A], [eo.scaling * scale], c)
            W["down_o"] = packing.pack_lora_down([eo.A], c, device=device)
            W["out"] = packing.pack([wo.float().cpu(), seg], bo, ops.choose_block_n(c, mt), 1, c, seg.shape[1], device=device)
        else:
            W["out"] = packing.pack([wo.float().cpu()], bo, ops.choose_block_n(c, mt), 1, c, device=device)
        self._cache[id(attn)] = (key, W)
        return W

    def __call__(self, attn, hidden_states: Tensor, encoder_hidden_states: Optional[Tensor] = None,
                 attention_mask: Optional[Tensor] = None, temb: Optional[Tensor] = None, scale: float = 1.0, **kwargs):
        if encoder_hidden_states is not None or attention_mask is not None:
            raise NotImplementedError("B200AttnProcessor implements the reference path only: self-attention "
                                      "(encoder_hidden_states=None), no mask")
        if hidden_states.dim() != 3:
            raise ValueError(f"expected [B, S, C] hidden_states, got {tuple(hidden_states.shape)}")
        if not hidden_states.is_cuda:
            raise RuntimeError("B200AttnProcessor needs CUDA tensors (no CPU fallback)")
        b, s, c = hidden_states.shape
        m = b * s
        dev = hidden_states.device
        W = self._packed(attn, m, scale, dev)
        x = hidden_states.to(torch.bfloat16).contiguous()
        bf = dict(dtype=torch.bfloat16, device=dev)
        T = None
        if "down_qkv" in W:
            T = ops.conv_gemm(W["down_qkv"], x, 1, m, 1, torch.empty(m, W["down_qkv"].n_valid, **bf))
        qkv = ops.conv_gemm(W["qkv"], x, 1, m, 1, torch.empty(m, 3 * c, **bf), a1=T)
        ao = ops.attention(qkv, torch.empty(m, c, **bf), b, s, attn.heads, c // attn.heads, scale=attn.scale)
        To = None
        if "down_o" in W:
            To = ops.conv_gemm(W["down_o"], ao, 1, m, 1, torch.empty(m, W["down_o"].n_valid, **bf))
        out = ops.conv_gemm(W["out"], ao, 1, m, 1, torch.empty(m, c, **bf), a1=To)
        return out.view(b, s, c).to(hidden_states.dtype)


class UNet2DConditionModel(nn.Module):
    """AudioLDM UNet (cvssp/audioldm-s-full-v2 layout, diffusers state-dict keys) on the B200 engine."""

    def __init__(self, arch: Union[str, UNetConfig] = "S", state_dict: Optional[Dict[str, Tensor]] = None,
                 device="cuda"):
        super().__init__()
        self.cfg = CONFIGS[arch] if isinstance(arch, str) else arch
        self.b200_device = torch.device(device)
        shapes = unet_param_shapes(self.cfg)
        if state_dict is None:
            from .synthetic import random_unet_state_dict
            state_dict = random_unet_state_dict(self.cfg, seed=0)
        missing = set(shapes) - set(state_dict)
        if missing:
            raise KeyError(f"state dict is missing {len(missing)} UNet keys, e.g. {sorted(missing)[:3]}")
        for k, shp in shapes.items():
            if tuple(state_dict[k].shape) != tuple(shp):
                raise ValueError(f"{k}: expected shape {shp}, got {tuple(state_dict[k].shape)}")
        self._sd = {k: state_dict[k].detach().float().cpu() for k in shapes}
        self.engine = UNetEngine(self.cfg, self._sd, self.b200_device)
        self.config = type("Cfg", (), dict(in_channels=self.cfg.in_channels, sample_size=self.cfg.sample_size,
                                           out_channels=self.cfg.out_channels))()
        # Attention modules (the LoRA targets) exposed under their diffusers names
        self._attn: "OrderedDict[str, Attention]" = OrderedDict()
        mods = nn.ModuleDict()
        for p in attention_paths(self.cfg):
            c = self._sd[p + ".to_q.weight"].shape[0]
            a = Attention(c, self.cfg.heads, p)
            for n in ("to_q", "to_k", "to_v"):
                getattr(a, n).weight.data.copy_(self._sd[f"{p}.{n}.weight"])
            a.to_out[0].weight.data.copy_(self._sd[p + ".to_out.0.weight"])
            a.to_out[0].bias.data.copy_(self._sd[p + ".to_out.0.bias"])
            a.requires_grad_(False)
            a.b200_device = self.b200_device
            self._attn[p] = a
            mods[p.replace(".", "__")] = a
        self.attn_modules = mods
        self._adapters: Dict[str, LoraEntry] = {}
        self.peft_config: Dict[str, LoraConfig] = {}
        self._trainer = None                 # LoraTrainer bound to the adapter parameters (grad-mode forward)
        self._trainer_synced = None
        self.eval()                          # like diffusers' from_pretrained; the reference calls unet.train() (train:479)

    # ------------------------------------------------------------------ attention-processor API
    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{p}.processor": a.processor for p, a in self._attn.items()}

    def set_attn_processor(self, processor) -> None:
        count = len(self._attn)
        if isinstance(processor, dict):
            if len(processor) != count:
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does "
                                 f"not match the number of attention layers: {count}. Please make sure to pass "
                                 f"{count} processor classes.")
            for p, a in self._attn.items():
                a.set_processor(processor[f"{p}.processor"])
        else:
            for a in self._attn.values():
                a.set_processor(processor)

    def set_default_attn_processor(self) -> None:
        self.set_attn_processor(B200AttnProcessor())

    # ------------------------------------------------------------------ LoRA-loader API
    def _channels_of(self, linear_path: str) -> int:
        return self._sd[linear_path.rsplit(".to_", 1)[0] + ".to_q.weight"].shape[0]

    def _install(self, adapters: Dict[str, LoraEntry], adapter_name: str = "default") -> None:
        known = {f"{p}.{n}" for p in self._attn for n in ("to_q", "to_k", "to_v", "to_out.0")}
        for path, e in adapters.items():
            if path not in known:
                raise KeyError(f"LoRA target {path} is not an attention projection of this UNet")
            c = self._channels_of(path)
            if e.A.shape[1] != c or e.B.shape[0] != c:
                raise ValueError(f"LoRA shapes for {path}: A {tuple(e.A.shape)} B {tuple(e.B.shape)} vs C={c}")
            attn_path, lin = path.rsplit(".to_", 1)
            attn = self._attn[attn_path]
            holder, attr = (attn.to_out, 0) if lin == "out.0" else (attn, "to_" + lin)
            cur = holder[attr] if isinstance(attr, int) else getattr(holder, attr)
            if not isinstance(cur, LoraLinear):
                cur = LoraLinear(cur)
                if isinstance(attr, int):
                    holder[attr] = cur
                else:
                    setattr(holder, attr, cur)
            cur.update_layer(adapter_name, e.A, e.B, e.alpha)
            if cur.base_layer.weight.is_cuda:
                cur.to(cur.base_layer.weight.device)
            self._adapters[path] = e
        self.engine.set_lora(self._adapters, self.engine.lora_scale)
        if not getattr(self, "_installing_from_trainer", False):
            self._trainer = None             # adapters changed under the trainer: rebuild it on the next grad-mode call
            self._trainer_synced = None

    def add_adapter(self, adapter_config: LoraConfig, adapter_name: str = "default") -> None:
        if adapter_name in self.peft_config:
            raise ValueError(f"Adapter with name {adapter_name} already exists. Please use a different name.")
        self.peft_config[adapter_name] = adapter_config
        self._install(init_adapters(self.cfg, adapter_config, self._channels_of), adapter_name)

    def load_lora_state_dict(self, sd: Dict[str, Tensor], alpha: Optional[float] = None,
                             adapter_name: str = "default") -> None:
        cfg = self.peft_config.get(adapter_name)
        self._install(parse_lora_state_dict(sd, alpha if alpha is not None else (cfg.lora_alpha if cfg else None)),
                      adapter_name)

    def load_attn_procs(self, pretrained_model_name_or_path_or_dict, **kwargs) -> None:
        """diffusers' `load_attn_procs` (app.py:11): a state dict, a checkpoint file or a directory holding one.
        alpha: `network_alpha=` if given, else what `save_attn_procs` / `save_lora_checkpoint` recorded in the file's
        metadata, else the adapter's rank (scaling 1)."""
        src = pretrained_model_name_or_path_or_dict
        alpha = kwargs.get("network_alpha")
        if isinstance(src, dict):
            self.load_lora_state_dict(src, alpha=alpha)
            return
        p = Path(src)
        if p.is_dir() or p.suffix == ".safetensors":
            from .lora import load_lora_checkpoint
            self._install(load_lora_checkpoint(p, alpha=alpha))
        else:
            self.load_lora_state_dict(torch.load(str(p), map_location="cpu"), alpha=alpha)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """peft-keyed LoRA tensors are routed to the adapters; base tensors must match the engine's."""
        lora = {k: v for k, v in state_dict.items() if ".lora_A" in k or ".lora_B" in k or ".lora." in k}
        unexpected = [k for k in state_dict if k not in lora and
                      k.replace("base_model.model.", "").replace(".base_layer", "") not in self._sd]
        if strict and unexpected:
            raise RuntimeError(f"Unexpected key(s) in state_dict: {unexpected[:5]}")
        if lora:
            self.load_lora_state_dict(lora)
        return type("IncompatibleKeys", (), {"missing_keys": [], "unexpected_keys": unexpected})()

    def lora_state_dict(self, adapter_name: Optional[str] = None) -> Dict[str, Tensor]:
        self._sync_from_trainer()            # the reference checkpoints mid-training (train_audioldm_lora.py:578)
        return to_peft_state_dict(self._adapters, adapter_name)

    def save_attn_procs(self, save_directory, safe_serialization: bool = True, fmt: str = "diffusers", **kwargs):
        """diffusers' `UNet2DConditionLoadersMixin.save_attn_procs`: the adapters as `pytorch_lora_weights.safetensors`
        (`lora.down/up` keys; what train_audioldm_lora.py:577-579 prepares and app.py:11 loads).  fmt="peft_full"
        writes accelerate's `model.safetensors` instead (generate_audio.py:32)."""
        if not safe_serialization:
            raise NotImplementedError("only safetensors checkpoints are written")
        self._sync_from_trainer()
        from .lora import save_lora_checkpoint
        return save_lora_checkpoint(self._adapters, save_directory, fmt)

    def merged_state_dict(self, scale: float = 1.0) -> Dict[str, Tensor]:
        """Base weights with the adapters folded in, W' = W + (alpha / r) * scale * B A (peft `merge_and_unload`):
        `UNet2DConditionModel(arch, unet.merged_state_dict())` is the adapter-free model.  Opt-in only -- the
        run-time path keeps LoRA unmerged like the reference, and the merged model is not bit-identical to it."""
        from .lora import merge_lora_into_state_dict
        self._sync_from_trainer()
        return merge_lora_into_state_dict(self._sd, self._adapters, scale)

    def custom_attn_processors(self) -> Optional[dict]:
        """{attention path: module} for modules whose processor is not the native one."""
        custom = {p: a for p, a in self._attn.items() if not isinstance(a.processor, B200AttnProcessor)}
        return custom or None

    # ------------------------------------------------------------------ training seam (autograd-capable forward)
    def _lora_linear(self, path: str) -> "LoraLinear":
        attn_path, lin = path.rsplit(".to_", 1)
        attn = self._attn[attn_path]
        return attn.to_out[0] if lin == "out.0" else getattr(attn, "to_" + lin)

    def lora_trainer(self, **kw):
        """The LoraTrainer behind grad-mode `forward`: flat fp32 device arenas for the adapter parameters and their
        gradients.  The peft-shaped `lora_A/lora_B[...].weight` Parameters become VIEWS into the parameter arena, so
        `torch.optim.AdamW(filter(requires_grad, unet.parameters()))` (train_audioldm_lora.py:394-403) updates the
        arena in place and `LoraTrainer`'s own fused optimizer sees the same memory."""
        if self._trainer is None:
            from .train import LoraTrainer
            tr = LoraTrainer(self, **kw)
            for path, (A, B) in tr.param_views().items():
                lin = self._lora_linear(path)
                name = "default" if "default" in lin.lora_A else next(iter(lin.lora_A))
                lin.lora_A[name].weight.data = A
                lin.lora_B[name].weight.data = B
                lin.lora_A[name].weight.requires_grad_(True)
                lin.lora_B[name].weight.requires_grad_(True)
            self._trainer = tr
            self._trainer_synced = tr.flat_p.detach().clone()      # built from the adapters just installed: in sync
        return self._trainer

    def _trainer_params(self):
        out = []
        for path in self._trainer.slots:
            lin = self._lora_linear(path)
            name = "default" if "default" in lin.lora_A else next(iter(lin.lora_A))
            out += [lin.lora_A[name].weight, lin.lora_B[name].weight]
        return out

    def _sync_from_trainer(self) -> None:
        """Bring `_adapters` and every packed engine plan (all shapes, LayerNorm-folded forms included) up to date with
        the trainer's flat parameter arena.  The arena is written behind our back -- the reference's loop runs
        `torch.optim.AdamW` over the peft-shaped Parameters, which are views into it with their own version counters --
        so the check compares VALUES against a device snapshot of the last sync: one 3.6 MB compare per eval forward /
        pipeline call / checkpoint, never inside the denoising loop."""
        tr = self._trainer
        if tr is None:
            return
        if self._trainer_synced is not None and self._trainer_synced.shape == tr.flat_p.shape and \
                torch.equal(self._trainer_synced, tr.flat_p):
            return
        self._installing_from_trainer = True
        try:
            from .engine import LoraEntry
            host = tr.flat_p.detach().float().cpu()
            ad = {}
            for p, s in tr.slots.items():
                ad[p] = LoraEntry(host[s.off_a: s.off_a + s.r * s.c].view(s.r, s.c).clone(),
                                  host[s.off_b: s.off_b + s.r * s.c].view(s.c, s.r).clone(), s.scaling * s.r)
            self._adapters.update(ad)
            self.engine.set_lora(self._adapters, self.engine.lora_scale)
        finally:
            self._installing_from_trainer = False
        self._trainer_synced = tr.flat_p.detach().clone()

    def sync_adapters(self) -> None:
        """Public form of the above for callers that drive `unet.engine` directly (AudioLDMPipeline)."""
        self._sync_from_trainer()

    # ------------------------------------------------------------------ forward
    def _sync_engine_lora(self, scale: float) -> None:
        self.engine.set_lora_scale(scale)

    def forward(self, sample: Tensor, timestep, encoder_hidden_states: Optional[Tensor] = None,
                class_labels: Optional[Tensor] = None, timestep_cond=None, attention_mask=None,
                cross_attention_kwargs: Optional[dict] = None, return_dict: bool = True, **kwargs):
        if encoder_hidden_states is not None:
            raise NotImplementedError("the reference path calls the AudioLDM UNet with encoder_hidden_states=None "
                                      "(train_audioldm_lora.py:542)")
        if class_labels is None:
            raise ValueError("class_labels should be provided when num_class_embeds > 0")
        if sample.dim() != 4 or sample.shape[1] != self.cfg.in_channels:
            raise ValueError(f"expected sample [B, {self.cfg.in_channels}, H, W], got {tuple(sample.shape)}")
        scale = float((cross_attention_kwargs or {}).get("scale", 1.0))
        if self.training and torch.is_grad_enabled() and self._adapters and self.custom_attn_processors() is None:
            tr = self.lora_trainer()
            params = self._trainer_params()
            if any(p.requires_grad for p in params):
                tr.lora_scale = scale
                eps = _UNetLoraFunction.apply(self, sample, timestep, class_labels, *params).to(sample.dtype)
                return UNet2DConditionOutput(sample=eps) if return_dict else (eps,)
        self._sync_from_trainer()
        self._sync_engine_lora(scale)
        eps = self.engine.forward(sample, timestep, class_labels, attn_overrides=self.custom_attn_processors(),
                                  lora_scale=scale)
        eps = eps.to(sample.dtype)
        return UNet2DConditionOutput(sample=eps) if return_dict else (eps,)


class _UNetLoraFunction(torch.autograd.Function):
    """Grad-mode UNet forward: the B200 training forward now, the B200 backward walk when autograd asks for the
    adapter gradients (`accelerator.backward(loss)`, train_audioldm_lora.py:557).  Inputs other than the LoRA
    parameters get no gradient (the reference freezes everything else, train:373-376)."""

    @staticmethod
    def forward(ctx, unet, sample, timestep, class_labels, *params):
        from .engine import LATENT_C_PAD
        tr = unet._trainer
        eng, dev = unet.engine, unet.b200_device
        nb, c, h, w = sample.shape
        x = sample.detach().to(dev, torch.float32).contiguous()
        t = torch.as_tensor(timestep, dtype=torch.float32, device=dev).reshape(-1)
        t = t.expand(nb).contiguous() if t.numel() == 1 else t.contiguous()
        tr.refresh(nb, h, w)
        # One activation arena, one outstanding grad-mode forward: a forward whose backward never ran (a validation
        # loss computed without no_grad, say) is abandoned here, and its backward -- should it come later -- raises.
        ctx.b200_token = tr.begin_forward()
        xin = torch.zeros(nb, h * w, LATENT_C_PAD, dtype=torch.bfloat16, device=dev)
        ops.pack_nchw_to_nhwc(x, nb, c, h * w, LATENT_C_PAD, xin)
        silu_emb = torch.empty(nb, eng.cfg.temb_channels, dtype=torch.bfloat16, device=dev)
        eng.embed(t, None, True, class_labels.detach().to(dev, torch.float32).contiguous(), None, silu_emb)
        eps_nhwc = torch.empty(nb, h * w, eng.cfg.out_channels, dtype=torch.float32, device=dev)
        ctx.b200 = (unet, tr.forward_train(xin, silu_emb, nb, h, w, eps_nhwc), (nb, h, w))
        out = torch.empty(nb, eng.cfg.out_channels, h, w, dtype=torch.float32, device=dev)
        ops.unpack_nhwc_to_nchw(eps_nhwc, nb, eng.cfg.out_channels, h * w, out)
        return out

    @staticmethod
    def backward(ctx, d_out):
        from .engine import LATENT_C_PAD
        unet, fctx, (nb, h, w) = ctx.b200
        tr = unet._trainer
        if tr is None or ctx.b200_token != tr.outstanding_forward:
            raise RuntimeError("B200 UNet backward: the activations of this forward are gone -- a later grad-mode forward "
                               "(or a change of adapters) reused the activation arena.  Run each grad-mode forward's "
                               "backward before the next forward, or wrap loss-only forwards in torch.no_grad().")
        deps = tr.arena.alloc((nb * h * w, LATENT_C_PAD), torch.bfloat16)
        deps.zero_()
        ops.pack_nchw_to_nhwc(d_out.detach().to(torch.float32).contiguous(), nb, d_out.shape[1], h * w, LATENT_C_PAD, deps)
        saved = tr.flat_g.clone()            # autograd accumulates into .grad itself: hand it this call's gradients only
        tr.flat_g.zero_()
        tr.backward(fctx, deps)
        grads = []
        for a, b in tr.grad_views().values():
            grads += [a.clone(), b.clone()]
        tr.flat_g.copy_(saved)
        return (None, None, None, None, *grads)


def get_peft_model(unet: UNet2DConditionModel, config: LoraConfig, adapter_name: str = "default") -> UNet2DConditionModel:
    """peft.get_peft_model: mutates `unet` in place and returns it (generate_audio.py:29)."""
    unet.add_adapter(config, adapter_name)
    return unet


def get_peft_model_state_dict(unet: UNet2DConditionModel) -> Dict[str, Tensor]:
    return unet.lora_state_dict(None)
