"""UNet architecture tables for the AudioLDM hot path (product side).

Describes `cvssp/audioldm-s-full-v2/unet/config.json` (and the L variant) the way the
reference loads it (`UNet2DConditionModel.from_pretrained(base_model_id, subfolder="unet")`,
/root/reference/script/train/train_audioldm_lora.py:364,
/root/reference/script/inference/generate_audio.py:18).  This is an enumeration written
independently of oracle/unet_ref.py; tests/test_oracle.py cross-checks the two.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple


@dataclass(frozen=True)
class UNetConfig:
    name: str
    block_out_channels: Tuple[int, ...]
    in_channels: int = 8
    out_channels: int = 8
    layers_per_block: int = 2
    heads: int = 8
    groups: int = 32
    class_in_dim: int = 512
    sample_size: int = 128

    @property
    def time_proj_dim(self) -> int:
        return self.block_out_channels[0]

    @property
    def time_embed_dim(self) -> int:
        return 4 * self.block_out_channels[0]

    @property
    def temb_channels(self) -> int:
        return 2 * self.time_embed_dim       # class_embeddings_concat=True


AUDIOLDM_S = UNetConfig("audioldm-s-full-v2", (128, 256, 384, 640))
AUDIOLDM_L = UNetConfig("audioldm-l-full", (256, 512, 768, 1280))
CONFIGS = {"S": AUDIOLDM_S, "L": AUDIOLDM_L}


@dataclass
class ResnetDesc:
    name: str
    cin: int
    cout: int
    skip_c: int = 0            # channels that come from the skip stack (up blocks), part of cin

    @property
    def has_shortcut(self) -> bool:
        return self.cin != self.cout


@dataclass
class TfmDesc:
    name: str
    c: int


@dataclass
class StageDesc:
    """One resolution-preserving unit: resnet (+ transformer)."""
    resnet: ResnetDesc
    tfm: Optional[TfmDesc] = None


@dataclass
class UNetGraph:
    cfg: UNetConfig
    down: List[List[StageDesc]] = field(default_factory=list)       # per level
    downsamplers: List[Optional[str]] = field(default_factory=list)
    mid: List[object] = field(default_factory=list)                  # [ResnetDesc, TfmDesc, ResnetDesc]
    up: List[List[StageDesc]] = field(default_factory=list)
    upsamplers: List[Optional[str]] = field(default_factory=list)

    def resnets(self) -> Iterator[ResnetDesc]:
        for lvl in self.down:
            for s in lvl:
                yield s.resnet
        yield self.mid[0]
        yield self.mid[2]
        for lvl in self.up:
            for s in lvl:
                yield s.resnet

    def transformers(self) -> Iterator[TfmDesc]:
        for lvl in self.down:
            for s in lvl:
                if s.tfm:
                    yield s.tfm
        yield self.mid[1]
        for lvl in self.up:
            for s in lvl:
                if s.tfm:
                    yield s.tfm


def build_graph(cfg: UNetConfig) -> UNetGraph:
    g = UNetGraph(cfg)
    boc = cfg.block_out_channels
    nlev = len(boc)
    skips = [boc[0]]
    prev = boc[0]
    for i, c in enumerate(boc):
        stages = []
        for j in range(cfg.layers_per_block):
            r = ResnetDesc(f"down_blocks.{i}.resnets.{j}", prev, c)
            t = TfmDesc(f"down_blocks.{i}.attentions.{j}", c) if i > 0 else None
            stages.append(StageDesc(r, t))
            prev = c
            skips.append(c)
        g.down.append(stages)
        if i < nlev - 1:
            g.downsamplers.append(f"down_blocks.{i}.downsamplers.0.conv")
            skips.append(c)
        else:
            g.downsamplers.append(None)
    c = boc[-1]
    g.mid = [ResnetDesc("mid_block.resnets.0", c, c), TfmDesc("mid_block.attentions.0", c),
             ResnetDesc("mid_block.resnets.1", c, c)]
    for i, c in enumerate(reversed(boc)):
        stages = []
        for j in range(cfg.layers_per_block + 1):
            sk = skips.pop()
            r = ResnetDesc(f"up_blocks.{i}.resnets.{j}", prev + sk, c, skip_c=sk)
            t = TfmDesc(f"up_blocks.{i}.attentions.{j}", c) if i < nlev - 1 else None
            stages.append(StageDesc(r, t))
            prev = c
        g.up.append(stages)
        g.upsamplers.append(f"up_blocks.{i}.upsamplers.0.conv" if i < nlev - 1 else None)
    assert not skips
    return g


def unet_param_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """diffusers state-dict keys -> shapes."""
    g = build_graph(cfg)
    P: Dict[str, Tuple[int, ...]] = {}
    ted, tch = cfg.time_embed_dim, cfg.temb_channels

    def wb(n, *shape):
        P[n + ".weight"] = tuple(shape)
        P[n + ".bias"] = (shape[0],)

    wb("time_embedding.linear_1", ted, cfg.time_proj_dim)
    wb("time_embedding.linear_2", ted, ted)
    wb("class_embedding", ted, cfg.class_in_dim)
    wb("conv_in", cfg.block_out_channels[0], cfg.in_channels, 3, 3)
    for r in g.resnets():
        wb(r.name + ".norm1", r.cin)
        wb(r.name + ".conv1", r.cout, r.cin, 3, 3)
        wb(r.name + ".time_emb_proj", r.cout, tch)
        wb(r.name + ".norm2", r.cout)
        wb(r.name + ".conv2", r.cout, r.cout, 3, 3)
        if r.has_shortcut:
            wb(r.name + ".conv_shortcut", r.cout, r.cin, 1, 1)
    for t in g.transformers():
        c = t.c
        wb(t.name + ".norm", c)
        wb(t.name + ".proj_in", c, c, 1, 1)
        b = t.name + ".transformer_blocks.0"
        for k in (1, 2, 3):
            wb(f"{b}.norm{k}", c)
        for a in ("attn1", "attn2"):
            for p in ("to_q", "to_k", "to_v"):
                P[f"{b}.{a}.{p}.weight"] = (c, c)
            wb(f"{b}.{a}.to_out.0", c, c)
        wb(f"{b}.ff.net.0.proj", 8 * c, c)
        wb(f"{b}.ff.net.2", c, 4 * c)
        wb(t.name + ".proj_out", c, c, 1, 1)
    for n in g.downsamplers + g.upsamplers:
        if n:
            c = _conv_ch(n, cfg)
            wb(n, c, c, 3, 3)
    wb("conv_norm_out", cfg.block_out_channels[0])
    wb("conv_out", cfg.out_channels, cfg.block_out_channels[0], 3, 3)
    return P


def _conv_ch(name: str, cfg: UNetConfig) -> int:
    parts = name.split(".")
    i = int(parts[1])
    boc = cfg.block_out_channels
    return boc[i] if parts[0] == "down_blocks" else boc[len(boc) - 1 - i]


def attention_paths(cfg: UNetConfig) -> List[str]:
    """The 32 `Attention` module paths LoRA targets (e.g. '...transformer_blocks.0.attn1')."""
    out = []
    for t in build_graph(cfg).transformers():
        for a in ("attn1", "attn2"):
            out.append(f"{t.name}.transformer_blocks.0.{a}")
    return out


def level_sizes(h: int, w: int = 16, levels: int = 4) -> List[Tuple[int, int]]:
    """Spatial size per UNet level: stride-2 pad-1 k3 conv -> floor((x-1)/2)+1."""
    out = [(h, w)]
    for _ in range(levels - 1):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        out.append((h, w))
    return out
