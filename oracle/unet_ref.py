"""ORACLE (test infrastructure, not product code) -- parity unpinned by the reference.

Torch-eager fp32 restatement of the AudioLDM UNet forward as diffusers 0.32.2
(`UNet2DConditionModel.forward`) executes it for the reference's call sites
(/root/reference/script/train/train_audioldm_lora.py:539-546 and, implicitly,
/root/reference/app.py:14, /root/reference/script/inference/generate_audio.py:47-52):
`encoder_hidden_states=None`, `class_labels=<L2-normalised 512-d CLAP embedding>`.

diffusers / peft are NOT vendored under /root/reference and not installable in this
image (requirements.txt:24,90 pin diffusers 0.32.2 / peft 0.13.2), so this file restates
their published algorithm -- the reference holds no golden vectors
for the path; "parity unpinned".  It issues the same ATen ops diffusers would (NCHW
`F.conv2d`, `F.group_norm`, `F.layer_norm`, `F.linear`, `F.scaled_dot_product_attention`)
with no fusion, and takes a flat state dict with diffusers key names.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# Architecture spec: cvssp/audioldm-s-full-v2/unet/config.json
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class UNetSpec:
    block_out_channels: Tuple[int, ...]
    in_channels: int = 8
    out_channels: int = 8
    layers_per_block: int = 2
    num_heads: int = 8            # diffusers `attention_head_dim=8` is used as the head COUNT
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    class_in_dim: int = 512       # projection_class_embeddings_input_dim
    time_proj_dim: int = 128      # == block_out_channels[0] for S; L uses 256
    # down block i has attention iff attn_levels[i]
    attn_levels: Tuple[bool, ...] = (False, True, True, True)

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def temb_channels(self) -> int:   # class_embeddings_concat=True -> 2x
        return self.time_embed_dim * 2


ARCH_S = UNetSpec(block_out_channels=(128, 256, 384, 640), time_proj_dim=128)
ARCH_L = UNetSpec(block_out_channels=(256, 512, 768, 1280), time_proj_dim=256)
ARCHS = {"S": ARCH_S, "L": ARCH_L}


def param_shapes(spec: UNetSpec) -> Dict[str, Tuple[int, ...]]:
    """Every diffusers state-dict key of the UNet with its shape."""
    P: Dict[str, Tuple[int, ...]] = {}
    boc = spec.block_out_channels
    ted, tch = spec.time_embed_dim, spec.temb_channels

    def lin(name, cin, cout, bias=True):
        P[name + ".weight"] = (cout, cin)
        if bias:
            P[name + ".bias"] = (cout,)

    def conv(name, cin, cout, k):
        P[name + ".weight"] = (cout, cin, k, k)
        P[name + ".bias"] = (cout,)

    def norm(name, c):
        P[name + ".weight"] = (c,)
        P[name + ".bias"] = (c,)

    def resnet(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cin, cout, 3)
        lin(name + ".time_emb_proj", tch, cout)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".conv_shortcut", cin, cout, 1)

    def attn(name, c):
        for p in ("to_q", "to_k", "to_v"):
            lin(f"{name}.{p}", c, c, bias=False)
        lin(f"{name}.to_out.0", c, c, bias=True)

    def tfm(name, c):
        norm(name + ".norm", c)
        conv(name + ".proj_in", c, c, 1)
        b = name + ".transformer_blocks.0"
        norm(b + ".norm1", c); attn(b + ".attn1", c)
        norm(b + ".norm2", c); attn(b + ".attn2", c)
        norm(b + ".norm3", c)
        lin(b + ".ff.net.0.proj", c, 8 * c)
        lin(b + ".ff.net.2", 4 * c, c)
        conv(name + ".proj_out", c, c, 1)

    lin("time_embedding.linear_1", spec.time_proj_dim, ted)
    lin("time_embedding.linear_2", ted, ted)
    lin("class_embedding", spec.class_in_dim, ted)
    conv("conv_in", spec.in_channels, boc[0], 3)

    # down
    skip_ch: List[int] = [boc[0]]
    cout = boc[0]
    for i, c in enumerate(boc):
        cin, cout = cout, c
        for j in range(spec.layers_per_block):
            resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
            if spec.attn_levels[i]:
                tfm(f"down_blocks.{i}.attentions.{j}", cout)
            skip_ch.append(cout)
        if i != len(boc) - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
            skip_ch.append(cout)
    # mid
    c = boc[-1]
    resnet("mid_block.resnets.0", c, c)
    tfm("mid_block.attentions.0", c)
    resnet("mid_block.resnets.1", c, c)
    # up
    rev = list(reversed(boc))
    rev_attn = list(reversed(spec.attn_levels))
    prev = rev[0]
    for i, c in enumerate(rev):
        for j in range(spec.layers_per_block + 1):
            sk = skip_ch.pop()
            resnet(f"up_blocks.{i}.resnets.{j}", prev + sk, c)
            prev = c
            if rev_attn[i]:
                tfm(f"up_blocks.{i}.attentions.{j}", c)
        if i != len(rev) - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", c, c, 3)
    assert not skip_ch
    norm("conv_norm_out", boc[0])
    conv("conv_out", boc[0], spec.out_channels, 3)
    return P


def count_params(spec: UNetSpec) -> int:
    return sum(math.prod(s) for s in param_shapes(spec).values())


def attention_module_names(spec: UNetSpec) -> List[str]:
    """Dotted names of the 32 Attention modules."""
    names = sorted({k.rsplit(".to_q.weight", 1)[0] for k in param_shapes(spec) if k.endswith(".to_q.weight")})
    return names


# --------------------------------------------------------------------------------------
# LoRA (peft 0.13.2 `lora.Linear.forward`)
# --------------------------------------------------------------------------------------
class LoraSet:
    """Unmerged LoRA adapters: {module path (e.g. '...attn1.to_q'): (A [r,in], B [out,r], alpha)}.

    peft: y = base(x) + lora_B(lora_A(x)) * (alpha / r) * runtime_scale
    (reference config: train_audioldm_lora.py:378-385, generate_audio.py:21-29).
    """

    def __init__(self, adapters: Optional[Dict[str, Tuple[Tensor, Tensor, float]]] = None, scale: float = 1.0):
        self.adapters = adapters or {}
        self.scale = scale     # cross_attention_kwargs={"scale": s} (train:544)

    def linear(self, path: str, x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
        y = F.linear(x, w, b)
        ad = self.adapters.get(path)
        if ad is not None:
            A, B, alpha = ad
            r = A.shape[0]
            y = y + F.linear(F.linear(x.to(A.dtype), A), B) * (alpha / r * self.scale)
        return y


# --------------------------------------------------------------------------------------
# Forward
# --------------------------------------------------------------------------------------
def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0) -> [cos | sin]."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    arg = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)


# Parity instrument (tests/test_gpu_train.py): with STORAGE_ROUND = torch.bfloat16 every tensor the B200 path keeps in HBM
# between two kernels (conv / linear / norm / attention outputs) is rounded to that type here as well -- forward values
# with a straight-through estimator, gradients through a hook -- while all arithmetic stays fp32.  The difference between
# this and the plain fp32 run is what bf16 STORAGE costs; what remains against the kernels is arithmetic / bugs.
STORAGE_ROUND = None


def _st(x: Tensor) -> Tensor:
    if STORAGE_ROUND is None:
        return x
    dt = STORAGE_ROUND
    y = x + (x.to(dt).to(x.dtype) - x).detach()
    if y.requires_grad:
        y.register_hook(lambda g: g.to(dt).to(g.dtype))
    return y


def _gn(sd, name, x, groups, eps):
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps)


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride, padding=padding)


def _resnet(sd, name, x, emb, spec: UNetSpec):
    h = _st(F.silu(_gn(sd, name + ".norm1", x, spec.norm_num_groups, spec.norm_eps)))
    h = _conv(sd, name + ".conv1", h)
    t = F.linear(F.silu(emb), sd[name + ".time_emb_proj.weight"], sd[name + ".time_emb_proj.bias"])
    h = _st(h + t[:, :, None, None])
    h = _st(F.silu(_gn(sd, name + ".norm2", h, spec.norm_num_groups, spec.norm_eps)))
    h = _conv(sd, name + ".conv2", h)          # dropout p=0
    if name + ".conv_shortcut.weight" in sd:
        x = _conv(sd, name + ".conv_shortcut", x, padding=0)
    return _st((x + h) / 1.0)                   # output_scale_factor = 1


def _attention(sd, name, x, heads: int, lora: LoraSet):
    """AttnProcessor2_0 with encoder_hidden_states=None (self-attention)."""
    B, S, C = x.shape
    q = _st(lora.linear(name + ".to_q", x, sd[name + ".to_q.weight"], None))
    k = _st(lora.linear(name + ".to_k", x, sd[name + ".to_k.weight"], None))
    v = _st(lora.linear(name + ".to_v", x, sd[name + ".to_v.weight"], None))
    d = C // heads
    q = q.view(B, S, heads, d).transpose(1, 2)
    k = k.view(B, S, heads, d).transpose(1, 2)
    v = v.view(B, S, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = _st(o.transpose(1, 2).reshape(B, S, C))
    return lora.linear(name + ".to_out.0", o, sd[name + ".to_out.0.weight"], sd[name + ".to_out.0.bias"])


def _basic_transformer_block(sd, name, x, heads, lora):
    C = x.shape[-1]
    ln = lambda n, y: _st(F.layer_norm(y, (C,), sd[f"{name}.{n}.weight"], sd[f"{name}.{n}.bias"], 1e-5))
    x = _st(x + _attention(sd, name + ".attn1", ln("norm1", x), heads, lora))
    x = _st(x + _attention(sd, name + ".attn2", ln("norm2", x), heads, lora))   # encoder_hidden_states=None
    h = _st(F.linear(ln("norm3", x), sd[name + ".ff.net.0.proj.weight"], sd[name + ".ff.net.0.proj.bias"]))
    val, gate = h.chunk(2, dim=-1)
    h = _st(val * F.gelu(gate))                  # GEGLU, exact erf gelu
    h = F.linear(h, sd[name + ".ff.net.2.weight"], sd[name + ".ff.net.2.bias"])
    return _st(x + h)


def _transformer2d(sd, name, x, spec: UNetSpec, lora):
    B, C, H, W = x.shape
    res = x
    h = _st(_gn(sd, name + ".norm", x, spec.norm_num_groups, 1e-6))
    h = _st(_conv(sd, name + ".proj_in", h, padding=0))
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = _basic_transformer_block(sd, name + ".transformer_blocks.0", h, spec.num_heads, lora)
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    h = _conv(sd, name + ".proj_out", h, padding=0)
    return _st(h + res)


def compute_emb(sd, spec: UNetSpec, timestep, class_labels: Tensor) -> Tensor:
    B = class_labels.shape[0]
    t = torch.as_tensor(timestep, device=class_labels.device)
    if t.dim() == 0:
        t = t[None]
    t = t.expand(B)
    t_emb = timestep_embedding(t, spec.time_proj_dim).to(class_labels.dtype)
    e = F.linear(t_emb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    e = F.linear(F.silu(e), sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    c = F.linear(class_labels, sd["class_embedding.weight"], sd["class_embedding.bias"])
    return torch.cat([e, c], dim=-1)


def unet_forward(sd: Dict[str, Tensor], spec: UNetSpec, sample: Tensor, timestep, class_labels: Tensor,
                 lora: Optional[LoraSet] = None, taps: Optional[dict] = None) -> Tensor:
    """sample [B,8,H,16] NCHW, timestep scalar or [B], class_labels [B,512] -> eps [B,8,H,16].

    `taps`, if given, is filled with intermediate activations for per-layer parity tests.
    """
    lora = lora or LoraSet()
    boc = spec.block_out_channels
    emb = compute_emb(sd, spec, timestep, class_labels)
    h = _st(_conv(sd, "conv_in", sample))
    if taps is not None:
        taps["emb"] = emb; taps["conv_in"] = h
    skips = [h]
    for i in range(len(boc)):
        for j in range(spec.layers_per_block):
            h = _resnet(sd, f"down_blocks.{i}.resnets.{j}", h, emb, spec)
            if taps is not None:
                taps[f"down_blocks.{i}.resnets.{j}"] = h
            if spec.attn_levels[i]:
                h = _transformer2d(sd, f"down_blocks.{i}.attentions.{j}", h, spec, lora)
                if taps is not None:
                    taps[f"down_blocks.{i}.attentions.{j}"] = h
            skips.append(h)
        if i != len(boc) - 1:
            h = _st(_conv(sd, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=1))
            skips.append(h)
    h = _resnet(sd, "mid_block.resnets.0", h, emb, spec)
    h = _transformer2d(sd, "mid_block.attentions.0", h, spec, lora)
    h = _resnet(sd, "mid_block.resnets.1", h, emb, spec)
    if taps is not None:
        taps["mid_block"] = h
    rev_attn = list(reversed(spec.attn_levels))
    for i in range(len(boc)):
        for j in range(spec.layers_per_block + 1):
            sk = skips.pop()
            h = torch.cat([h, sk], dim=1)
            h = _resnet(sd, f"up_blocks.{i}.resnets.{j}", h, emb, spec)
            if rev_attn[i]:
                h = _transformer2d(sd, f"up_blocks.{i}.attentions.{j}", h, spec, lora)
            if taps is not None:
                taps[f"up_blocks.{i}.{j}"] = h
        if i != len(boc) - 1:
            size = skips[-1].shape[-2:]            # forward_upsample_size: explicit target size
            h = F.interpolate(h, size=tuple(size), mode="nearest")
            h = _st(_conv(sd, f"up_blocks.{i}.upsamplers.0.conv", h))
    h = _st(F.silu(_gn(sd, "conv_norm_out", h, spec.norm_num_groups, spec.norm_eps)))
    return _conv(sd, "conv_out", h)
