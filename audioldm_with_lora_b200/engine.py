"""UNet forward engine: schedules the hand-written kernels over an HBM arena (NHWC bf16).

This is the B200-native body of `UNet2DConditionModel.forward` as the reference drives it
(/root/reference/script/train/train_audioldm_lora.py:539-546: encoder_hidden_states=None,
class_labels = CLAP embedding; implicit in AudioLDMPipeline.__call__, /root/reference/app.py:14).
Graph semantics follow SURVEY.md App. A; the fp32 oracle in oracle/unet_ref.py is the parity check.

Data layout in HBM
  * activations: NHWC bf16, `[nb, H, W, C]` == token matrix `[nb*H*W, C]`; residual stream bf16,
    every GEMM accumulates in fp32 (TMEM) and the epilogue adds bias / embedding / residual in fp32.
  * weights: bf16 `[n_pad, K]` K-major, tap-major K for 3x3 convs (packing.py); biases / norm affine fp32.
  * `torch.cat([h, skip])` is never materialised: GroupNorm reads two sources; the 1x1 conv_shortcut
    over the concatenation is two extra K segments of conv2's GEMM.
  * LoRA stays UNMERGED (peft semantics): T = x.A^T (one small GEMM for q,k,v together), then
    [x | T] . [W | s.B]^T in the base GEMM's accumulator.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import ops, packing
from .arch import ResnetDesc, TfmDesc, UNetConfig, build_graph, level_sizes
from .ops import PackedWeight

Tensor = torch.Tensor
LATENT_C_PAD = 64          # conv_in input channels 8 -> one 64-channel K block
# ff.net.2 + proj_out as one GEMM in the sampling engine (B200_FFPROJ=0: two launches, as the fine-tuning walk keeps them)
FFPROJ_FUSED = os.environ.get("B200_FFPROJ", "1") != "0"


class Arena:
    """First-fit allocator over one device buffer; deterministic for a fixed call sequence so that
    pointers captured in a CUDA graph stay valid."""

    ALIGN = 1024

    def __init__(self, nbytes: int, device):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.free: List[Tuple[int, int]] = [(0, nbytes)]
        self.live: Dict[int, Tuple[int, int]] = {}
        self.attached: Dict[int, Tensor] = {}      # side buffers that live and die with a tensor (GroupNorm partial statistics)
        self.peak = 0

    def alloc(self, shape, dtype) -> Tensor:
        n = int(math.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        n = (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        for i, (off, size) in enumerate(self.free):
            if size >= n:
                if size == n:
                    self.free.pop(i)
                else:
                    self.free[i] = (off + n, size - n)
                t = self.buf[off: off + n].view(dtype)[: int(math.prod(shape))].view(*shape)
                self.live[t.data_ptr()] = (off, n)
                self.peak = max(self.peak, off + n)
                return t
        raise MemoryError(f"b200 arena exhausted allocating {n} bytes")

    def attach(self, t: Tensor, extra: Tensor) -> None:
        self.attached[t.data_ptr()] = extra

    def release(self, t: Optional[Tensor]) -> None:
        if t is None:
            return
        extra = self.attached.pop(t.data_ptr(), None)
        if extra is not None:
            self.release(extra)
        off, n = self.live.pop(t.data_ptr())
        self.free.append((off, n))
        self.free.sort()
        merged: List[Tuple[int, int]] = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        self.free = merged


@dataclass
class LoraEntry:
    A: Tensor          # [r, in]
    B: Tensor          # [out, r]
    alpha: float

    @property
    def scaling(self) -> float:
        return self.alpha / self.A.shape[0]


_PW_SCALARS = ("n_valid", "block_n", "ntaps", "c0", "c1", "c2", "geglu", "ksplit")


def _same_layout(a: PackedWeight, b: PackedWeight) -> bool:
    if a.w.shape != b.w.shape or a.w.device != b.w.device or any(getattr(a, f) != getattr(b, f) for f in _PW_SCALARS):
        return False
    for f in ("bias", "ln_g"):
        ta, tb = getattr(a, f, None), getattr(b, f, None)
        if (ta is None) != (tb is None) or (ta is not None and ta.shape != tb.shape):
            return False
    return True


def _overwrite(dst: PackedWeight, src: PackedWeight) -> None:
    """dst <- src without moving dst's device tensors (same layout, see _same_layout)."""
    dst.w.copy_(src.w)
    for f in ("bias", "ln_g"):
        if getattr(src, f, None) is not None:
            getattr(dst, f).copy_(getattr(src, f))
    for f in ("alg_macs_per_row", "lora_fused", "lora_rows"):
        if hasattr(src, f):
            setattr(dst, f, getattr(src, f))


class UNetEngine:
    def __init__(self, cfg: UNetConfig, state_dict: Dict[str, Tensor], device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.graph = build_graph(cfg)
        # fp32 master copy (CPU) in diffusers key names; packed per (nb, H, W) plan on demand
        self.sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        self.lora: Dict[str, LoraEntry] = {}
        self.lora_scale = 1.0
        self._plans: Dict[Tuple[int, int, int], dict] = {}
        self._small: Dict[str, Tensor] = {}
        self.attn_variant = 0
        self.weights_version = 0      # bumped whenever packed device weights are (re)built
        self.arena: Optional[Arena] = None
        # concurrent sub-batch branches (forward_branched): per-branch arenas and side streams
        self._branch_arenas: Dict[int, List[Arena]] = {}
        self._branch_streams: List["torch.cuda.Stream"] = []
        self._init_small()

    # ------------------------------------------------------------------ weights
    def _dev(self, t: Tensor, dtype=torch.float32) -> Tensor:
        return t.to(self.device, dtype).contiguous()

    def _init_small(self) -> None:
        sd, S = self.sd, self._small
        for k in ("time_embedding.linear_1", "time_embedding.linear_2", "class_embedding"):
            S[k + ".weight"] = self._dev(sd[k + ".weight"].t())          # [in, out]: coalesced in the embed kernel
            S[k + ".bias"] = self._dev(sd[k + ".bias"])
        for k, v in sd.items():
            if (".norm" in k or k.startswith("conv_norm_out")) and v.dim() == 1:
                S[k] = self._dev(v)

    def set_lora(self, adapters: Optional[Dict[str, LoraEntry]], scale: float = 1.0) -> None:
        """adapters: {'<attention path>.to_q' | '.to_k' | '.to_v' | '.to_out.0': LoraEntry}."""
        self.lora = dict(adapters or {})
        self.lora_scale = float(scale)
        for plan in self._plans.values():
            self._pack_attention(plan)

    def set_lora_scale(self, scale: float) -> None:
        if float(scale) != self.lora_scale:
            self.set_lora(self.lora, scale)

    def _plan(self, nb: int, h: int, w: int, sms: int = ops.NUM_SMS) -> dict:
        """Packed weights + tilings for a (sub-)batch shape, tiled for `sms` SMs (a concurrent chain's share)."""
        key = (nb, h, w) if sms == ops.NUM_SMS else (nb, h, w, sms)
        if key not in self._plans:
            self._tiling_sms = sms
            try:
                self._plans[key] = self._build_plan(nb, h, w)
            finally:
                self._tiling_sms = ops.NUM_SMS
            self._plans[key]["sms"] = sms
        return self._plans[key]

    def _pw(self, segs, bias, m_tiles, ntaps, c0, c1=0, c2=0, geglu=False, block_n=None, split=True,
            max_bn: int = 256) -> PackedWeight:
        n = segs[0].shape[0]
        num_kb = (ntaps * c0 + c1 + c2) // 64
        pair = None
        if block_n:
            bn, ks = block_n, 1
        else:
            bn, ks, pair = ops.choose_tiling_ex(n, m_tiles, num_kb, geglu, allow_split=split,
                                                num_sms=getattr(self, "_tiling_sms", ops.NUM_SMS), max_bn=max_bn)
        pw = packing.pack(segs, bias, bn, ntaps, c0, c1, c2, geglu, device=self.device, ksplit=ks)
        pw.pair = pair
        return pw

    def _ln_pack(self, w: Tensor, bias: Optional[Tensor], gamma: Tensor, beta: Tensor, m_tiles: int, c: int, *,
                 geglu: bool = False, lora_seg: Optional[Tensor] = None, max_bn: int = 256) -> PackedWeight:
        """Weights of a linear layer that consumes LayerNorm(x), with the LayerNorm folded in (b200_linear_ln):
        rows gamma o W (+ the s.B K segment), bias' = W beta + bias, ln_g = column sums of the bf16-rounded rows."""
        w = w.float()
        wg = w * gamma.float()[None, :]
        b2 = w @ beta.float() + (bias.float() if bias is not None else 0.0)
        segs = [wg] + ([lora_seg] if lora_seg is not None else [])
        pw = self._pw(segs, b2, m_tiles, 1, c, lora_seg.shape[1] if lora_seg is not None else 0, geglu=geglu,
                      split=False, max_bn=max_bn)
        pw.ln_g = pw.w[:, :c].float().sum(1).contiguous()
        pw.lora_fused = lora_seg is not None
        return pw

    def _build_plan(self, nb: int, h: int, w: int) -> dict:
        cfg, sd, g = self.cfg, self.sd, self.graph
        sizes = level_sizes(h, w, len(cfg.block_out_channels))
        plan: dict = {"sizes": sizes, "W": {}, "nb": nb}
        W = plan["W"]

        def conv_tiles(lvl):
            hh, ww = sizes[lvl]
            return ops.num_m_tiles(nb, hh, ww)

        def lin_tiles(lvl):
            hh, ww = sizes[lvl]
            return math.ceil(nb * hh * ww / 128)

        def conv3(name, lvl, c_pad=None, tiles_lvl=None):
            wk = packing.conv3x3_to_k(sd[name + ".weight"], c_pad)
            W[name] = self._pw([wk], sd[name + ".bias"], conv_tiles(lvl if tiles_lvl is None else tiles_lvl), 9, wk.shape[1] // 9)

        def resnet(r: ResnetDesc, lvl):
            conv3(r.name + ".conv1", lvl)
            w2 = packing.conv3x3_to_k(sd[r.name + ".conv2.weight"])
            if r.has_shortcut:
                ws = sd[r.name + ".conv_shortcut.weight"][:, :, 0, 0]
                c_h = r.cin - r.skip_c
                segs = [w2, ws[:, :c_h]] + ([ws[:, c_h:]] if r.skip_c else [])
                bias = sd[r.name + ".conv2.bias"] + sd[r.name + ".conv_shortcut.bias"]
                W[r.name + ".conv2"] = self._pw(segs, bias, conv_tiles(lvl), 9, r.cout, c_h, r.skip_c)
            else:
                W[r.name + ".conv2"] = self._pw([w2], sd[r.name + ".conv2.bias"], conv_tiles(lvl), 9, r.cout)

        def tfm(t: TfmDesc, lvl):
            c, mt = t.c, lin_tiles(lvl)
            b = t.name + ".transformer_blocks.0"
            W[t.name + ".proj_in"] = self._pw([sd[t.name + ".proj_in.weight"][:, :, 0, 0]], sd[t.name + ".proj_in.bias"], mt, 1, c)
            W[t.name + ".proj_out"] = self._pw([sd[t.name + ".proj_out.weight"][:, :, 0, 0]], sd[t.name + ".proj_out.bias"], mt, 1, c)
            W[b + ".ff.net.0.proj"] = self._pw([sd[b + ".ff.net.0.proj.weight"]], sd[b + ".ff.net.0.proj.bias"], mt, 1, c, geglu=True)
            if ops.ln_fusion_wanted("ff", mt):          # norm3 folded into ff.net.0.proj (engine._Runner.tfm)
                W[b + ".ff.net.0.proj.ln"] = self._ln_pack(sd[b + ".ff.net.0.proj.weight"], sd[b + ".ff.net.0.proj.bias"],
                                                           sd[b + ".norm3.weight"], sd[b + ".norm3.bias"], mt, c, geglu=True)
            W[b + ".ff.net.2"] = self._pw([sd[b + ".ff.net.2.weight"]], sd[b + ".ff.net.2.bias"], mt, 1, 4 * c)
            if FFPROJ_FUSED:
                # ff.net.2 (+ the block's residual) and Transformer2DModel.proj_out compose exactly -- nothing non-linear sits
                # between them:  proj_out(x + W2 g + b2) = (Wp W2) g + Wp x + (Wp b2 + bp).  One GEMM over the K segments
                # [g | x] replaces two launches and the token tensor between them (same FLOPs: C x 5C per row either way).
                wp, bp = sd[t.name + ".proj_out.weight"][:, :, 0, 0].float(), sd[t.name + ".proj_out.bias"].float()
                w2, b2 = sd[b + ".ff.net.2.weight"].float(), sd[b + ".ff.net.2.bias"].float()
                W[t.name + ".ffout_proj"] = self._pw([wp @ w2, wp], wp @ b2 + bp, mt, 1, 4 * c, c)

        lvl_of: Dict[str, int] = {}
        conv3("conv_in", 0, LATENT_C_PAD)
        W["conv_in"].alg_macs_per_row = float(cfg.block_out_channels[0] * 9 * cfg.in_channels)
        for i, stages in enumerate(g.down):
            for s in stages:
                resnet(s.resnet, i); lvl_of[s.resnet.name] = i
                if s.tfm:
                    tfm(s.tfm, i); lvl_of[s.tfm.name] = i
            if g.downsamplers[i]:
                conv3(g.downsamplers[i], i, tiles_lvl=i + 1)    # stride 2: its tiles are output pixels (strided TMA boxes)
        top = len(g.down) - 1
        resnet(g.mid[0], top); tfm(g.mid[1], top); resnet(g.mid[2], top)
        lvl_of[g.mid[0].name] = lvl_of[g.mid[1].name] = lvl_of[g.mid[2].name] = top
        for i, stages in enumerate(g.up):
            lvl = top - i
            for s in stages:
                resnet(s.resnet, lvl); lvl_of[s.resnet.name] = lvl
                if s.tfm:
                    tfm(s.tfm, lvl); lvl_of[s.tfm.name] = lvl
            if g.upsamplers[i]:
                conv3(g.upsamplers[i], lvl - 1)
        # conv_out: 8 output channels -> one 32-wide tile
        wk = packing.conv3x3_to_k(sd["conv_out.weight"])
        W["conv_out"] = self._pw([wk], sd["conv_out.bias"], conv_tiles(0), 9, wk.shape[1] // 9, block_n=32)
        # all 22 time_emb_proj layers as ONE GEMM over silu(emb): column offsets per resnet
        offs, ws, bs, off = {}, [], [], 0
        for r in g.resnets():
            offs[r.name] = off
            ws.append(sd[r.name + ".time_emb_proj.weight"]); bs.append(sd[r.name + ".time_emb_proj.bias"])
            off += r.cout
        plan["temb_off"], plan["temb_total"] = offs, off
        W["temb"] = self._pw([torch.cat(ws)], torch.cat(bs), 1, 1, cfg.temb_channels, block_n=128)
        plan["lvl_of"] = lvl_of
        self._pack_attention(plan)
        return plan

    def _pack_attention(self, plan: dict) -> None:
        """(Re)pack the attention projections of one plan.  Packed tensors whose shape and tiling are unchanged are
        overwritten IN PLACE, so device pointers captured in CUDA graphs (the denoising step, the fine-tuning step) and
        in the trainer's refresh table stay valid; `weights_version` moves only when a pointer did."""
        prev, self._tiling_sms = getattr(self, "_tiling_sms", ops.NUM_SMS), plan.get("sms", getattr(self, "_tiling_sms", ops.NUM_SMS))
        old = {k: v for k, v in plan["W"].items() if ".attn1." in k or ".attn2." in k}
        try:
            self._pack_attention_impl(plan)
        finally:
            self._tiling_sms = prev
        W, moved = plan["W"], False
        for k in [k for k in W if ".attn1." in k or ".attn2." in k]:
            new, was = W[k], old.pop(k, None)
            if was is not None and was is not new and _same_layout(was, new):
                _overwrite(was, new)
                W[k] = was
            else:
                moved = True
        if moved or old:                  # a tensor was created, resized or dropped
            self.weights_version += 1

    def _pack_attention_impl(self, plan: dict) -> None:
        sd, W, sizes, nb = self.sd, plan["W"], plan["sizes"], plan["nb"]
        for t in self.graph.transformers():
            lvl = plan["lvl_of"][t.name]
            mt = math.ceil(nb * sizes[lvl][0] * sizes[lvl][1] / 128)
            c = t.c
            for a in ("attn1", "attn2"):
                p = f"{t.name}.transformer_blocks.0.{a}"
                ents = [self.lora.get(f"{p}.{n}") for n in ("to_q", "to_k", "to_v")]
                a_list = [e.A if e else None for e in ents]
                wqkv = torch.cat([sd[f"{p}.{n}.weight"] for n in ("to_q", "to_k", "to_v")])
                if any(e is not None for e in ents):
                    kp = packing.lora_pad(sum(x.shape[0] for x in a_list if x is not None))
                    seg = packing.lora_up_segment([e.B if e else None for e in ents], a_list,
                                                  [e.scaling * self.lora_scale if e else 0.0 for e in ents], c)
                    r_tot = sum(x.shape[0] for x in a_list if x is not None)
                    W[p + ".lora_down_qkv"] = packing.pack_lora_down(a_list, c, device=self.device)
                    W[p + ".lora_down_qkv"].alg_macs_per_row = float(r_tot * c)
                    W[p + ".lora_down_qkv"].lora_rows = r_tot
                    fuse = ops.lora_fusion_pays(3 * c, mt, kp, getattr(self, "_tiling_sms", ops.NUM_SMS))
                    W[p + ".qkv"] = self._pw([wqkv, seg], None, mt, 1, c, kp, split=not fuse,
                                             max_bn=ops.LORA_FUSED_MAX_BN if fuse else 256)
                    W[p + ".qkv"].lora_fused = fuse
                    W[p + ".qkv"].alg_macs_per_row = float(3 * c * c + r_tot * c)
                else:
                    W.pop(p + ".lora_down_qkv", None)
                    W[p + ".qkv"] = self._pw([wqkv], None, mt, 1, c)
                # the same projection with norm1 / norm2 folded in (sampling path)
                W.pop(p + ".qkv.ln", None); W.pop(p + ".lora_down_qkv.ln", None)
                if ops.ln_fusion_wanted("qkv", mt):
                    ln = "norm1" if a == "attn1" else "norm2"
                    b_ = f"{t.name}.transformer_blocks.0"
                    gamma, beta = sd[f"{b_}.{ln}.weight"].float(), sd[f"{b_}.{ln}.bias"].float()
                    if any(e is not None for e in ents):
                        if ops.LORA_FUSED and kp == 64:
                            W[p + ".qkv.ln"] = self._ln_pack(wqkv, None, gamma, beta, mt, c, lora_seg=seg,
                                                             max_bn=ops.LORA_FUSED_MAX_BN)
                            W[p + ".qkv.ln"].alg_macs_per_row = float(3 * c * c + 2 * r_tot * c)
                            down = packing.pack_lora_down([x * gamma[None, :] if x is not None else None for x in a_list], c,
                                                          device=self.device)
                            down.lora_rows = r_tot
                            down.ln_g = down.w.float().sum(1).contiguous()
                            ba = torch.zeros(64)
                            ba[:r_tot] = torch.cat([x.float() for x in a_list if x is not None]) @ beta
                            down.bias = ba.to(self.device)
                            W[p + ".lora_down_qkv.ln"] = down
                    else:
                        W[p + ".qkv.ln"] = self._ln_pack(wqkv, None, gamma, beta, mt, c)
                eo = self.lora.get(f"{p}.to_out.0")
                wo, bo = sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"]
                if eo is not None:
                    kp = packing.lora_pad(eo.A.shape[0])
                    seg = packing.lora_up_segment([eo.B], [eo.A], [eo.scaling * self.lora_scale], c)
                    W[p + ".lora_down_o"] = packing.pack_lora_down([eo.A], c, device=self.device)
                    W[p + ".lora_down_o"].alg_macs_per_row = float(eo.A.shape[0] * c)
                    W[p + ".lora_down_o"].lora_rows = eo.A.shape[0]
                    fuse = ops.lora_fusion_pays(c, mt, kp, getattr(self, "_tiling_sms", ops.NUM_SMS))
                    W[p + ".to_out"] = self._pw([wo, seg], bo, mt, 1, c, kp, split=not fuse,
                                                max_bn=ops.LORA_FUSED_MAX_BN if fuse else 256)
                    W[p + ".to_out"].lora_fused = fuse
                    W[p + ".to_out"].alg_macs_per_row = float(c * c + eo.A.shape[0] * c)
                else:
                    W.pop(p + ".lora_down_o", None)
                    W[p + ".to_out"] = self._pw([wo], bo, mt, 1, c)

    # ------------------------------------------------------------------ arena
    def _arena_bytes(self, nb: int, h: int, w: int) -> int:
        c0 = self.cfg.block_out_channels[0]
        return nb * h * w * c0 * 2 * 40 + (64 << 20)      # ~40 level-0-sized bf16 tensors: generous

    def _ensure_arena(self, nb: int, h: int, w: int) -> Arena:
        need = self._arena_bytes(nb, h, w)
        if self.arena is None or self.arena.buf.numel() < need:
            self.arena = Arena(need, self.device)
        return self.arena

    def _ensure_branches(self, branches: int, nb_sub: int, h: int, w: int) -> List[Arena]:
        need = self._arena_bytes(nb_sub, h, w)
        ars = self._branch_arenas.get(branches)
        if ars is None or ars[0].buf.numel() < need:
            ars = self._branch_arenas[branches] = [Arena(need, self.device) for _ in range(branches)]
        while self.device.type == "cuda" and len(self._branch_streams) < branches - 1:
            self._branch_streams.append(torch.cuda.Stream(device=self.device))
        return ars

    def arena_token(self, nb: int, h: int, w: int, branches: int = 1) -> tuple:
        """Identity of the device buffers a captured CUDA graph of this shape points into."""
        branches = self.effective_branches(nb, branches)
        tok = (id(self._ensure_arena(nb, h, w)),)
        if branches > 1:
            tok += tuple(id(a) for a in self._ensure_branches(branches, nb // branches, h, w))
        return tok

    @staticmethod
    def effective_branches(nb: int, branches: int) -> int:
        branches = max(1, int(branches))
        while branches > 1 and nb % branches:
            branches -= 1
        return branches

    # ------------------------------------------------------------------ forward
    def embed(self, t_steps: Tensor, step_ptr: Optional[Tensor], per_sample: bool, labels: Tensor,
              emb_out: Optional[Tensor], silu_out: Tensor) -> None:
        cfg, S = self.cfg, self._small
        ops.time_class_embed(t_steps, step_ptr, per_sample, labels, labels.shape[0], cfg.time_proj_dim,
                             cfg.time_embed_dim, cfg.class_in_dim, S["time_embedding.linear_1.weight"],
                             S["time_embedding.linear_1.bias"], S["time_embedding.linear_2.weight"],
                             S["time_embedding.linear_2.bias"], S["class_embedding.weight"], S["class_embedding.bias"],
                             emb_out, silu_out)

    # ------------------------------------------------------------------ the forward as a linear program
    def program(self) -> List[tuple]:
        """The UNet forward as a flat list of steps over (hcur, skip stack):
        ("conv_in",) ("res", ResnetDesc, lvl, pops_skip) ("tfm", TfmDesc, lvl) ("push", channels) ("down", name, lvl)
        ("up", name, lvl_from, channels) ("out",).  A contiguous slice of it can run on a sub-batch (forward_nhwc)."""
        if getattr(self, "_program", None) is None:
            cfg, g = self.cfg, self.graph
            prog: List[tuple] = [("conv_in",), ("push", cfg.block_out_channels[0])]
            for i, stages in enumerate(g.down):
                for s in stages:
                    prog.append(("res", s.resnet, i, False))
                    if s.tfm:
                        prog.append(("tfm", s.tfm, i))
                    prog.append(("push", s.resnet.cout))
                if g.downsamplers[i]:
                    prog.append(("down", g.downsamplers[i], i))
                    prog.append(("push", cfg.block_out_channels[i]))
            top = len(g.down) - 1
            prog += [("res", g.mid[0], top, False), ("tfm", g.mid[1], top), ("res", g.mid[2], top, False),
                     ("tap", "mid_block", top, cfg.block_out_channels[-1])]
            for i, stages in enumerate(g.up):
                lvl = top - i
                for j, s in enumerate(stages):
                    prog.append(("res", s.resnet, lvl, True))
                    if s.tfm:
                        prog.append(("tfm", s.tfm, lvl))
                    prog.append(("tap", f"up_blocks.{i}.{j}", lvl, s.resnet.cout))
                if g.upsamplers[i]:
                    prog.append(("up", g.upsamplers[i], lvl, stages[-1].resnet.cout))
            prog.append(("out",))
            self._program = prog
        return self._program

    def middle_region(self, min_level: int = 2) -> Optional[Tuple[int, int, int]]:
        """(i0, i1, outer_skips): the contiguous run of program steps at levels >= min_level (deep down blocks, mid
        block, the matching up blocks and the up-sampler that leaves the region) and how many skip tensors produced
        BEFORE the region it pops.  None if the architecture has no such level."""
        prog = self.program()

        def lvl_of(st):
            if st[0] in ("res", "tfm", "tap"):
                return st[2]
            if st[0] == "down":
                return st[2] + 1            # belongs to the level it produces
            if st[0] == "up":
                return st[2]                # consumes lvl, produces lvl - 1: last step of the region
            return None

        idx = [k for k, st in enumerate(prog) if (lvl_of(st) or 0) >= min_level]
        if not idx or prog[idx[0]][0] != "down" or prog[idx[-1]][0] != "up":
            return None
        i0, i1 = idx[0], idx[-1] + 1
        depth, need = 0, 0
        for st in prog[i0:i1]:
            if st[0] == "push":
                depth += 1
            elif st[0] == "res" and st[3]:
                if depth == 0:
                    need += 1
                else:
                    depth -= 1
        return i0, i1, need

    def forward_nhwc(self, xin: Tensor, silu_emb: Tensor, nb: int, h: int, w: int, eps_out: Tensor,
                     taps: Optional[dict] = None, attn_overrides: Optional[dict] = None,
                     mid_branches: int = 1, rowvec: Optional[Tensor] = None, rowvec_ready=None) -> Tensor:
        """xin bf16 [nb, h, w, 64] (channels >= 8 zero), silu_emb bf16 [nb, temb_channels]
        -> eps_out fp32 [nb, h*w, 8] (NHWC).

        mid_branches > 1: the deep levels (63x4 and 32x2 latents at 10 s: ~55 % of the step's launches, each with
        only 8..96 tiles for 148 SMs and therefore bound by per-launch latency) run as `mid_branches` independent
        sub-batch chains on parallel streams -- captured into the denoising-step CUDA graph as parallel branches.
        Measured on B200 (tools/concurrency_probe.py): chains whose kernels together stay under ~148 CTAs overlap
        perfectly (7.0 us per link for 1, 2 or 3 chains); the shallow levels already fill the GPU and stay one chain."""
        prog = self.program()
        # rowvec (+ rowvec_ready, a CUDA event): the batched time_emb_proj GEMM already computed -- or still being computed
        # on a side stream (AudioLDMPipeline overlaps the embedding kernels with conv_in); the first ResNet waits for it.
        main = _Runner(self, nb, h, w, self._ensure_arena(nb, h, w), silu_emb, taps, attn_overrides, rowvec=rowvec)
        main.rowvec_ready = rowvec_ready
        region = self.middle_region() if (mid_branches > 1 and taps is None and not attn_overrides) else None
        branches = self.effective_branches(nb, mid_branches) if region else 1
        if branches <= 1:
            hcur, skips = main.run(prog, None, [], xin=xin, eps_out=eps_out)
            main.finish()
            return eps_out
        i0, i1, need = region
        hcur, skips = main.run(prog[:i0], None, [], xin=xin)
        assert prog[i0][0] == "down" and len(skips) >= need
        lvl_in = prog[i0][2]                # the region starts with the down-sampler into its first level
        up_step = prog[i1 - 1]
        assert up_step[0] == "up"
        lvl_out, c_out = up_step[2] - 1, up_step[3]
        region_out = main.ar.alloc((main.M(lvl_out), c_out), torch.bfloat16)
        outer = [skips.pop() for _ in range(need)][::-1]          # bottom .. top of what the region pops
        sub = nb // branches
        arenas = self._ensure_branches(branches, sub, h, w)
        rows_in = sub * main.sizes[lvl_in][0] * main.sizes[lvl_in][1]
        rows_out = sub * main.sizes[lvl_out][0] * main.sizes[lvl_out][1]

        share = max(1, ops.NUM_SMS // branches)

        def run_branch(b):
            r = _Runner(self, sub, h, w, arenas[b], None, None, None, rowvec=main.rowvec[b * sub:(b + 1) * sub],
                        sms=share)
            sk = [(t[b * rows_in:(b + 1) * rows_in], c) for t, c in outer]
            ops.set_sm_budget(share)
            try:
                r.run(prog[i0:i1], hcur[b * rows_in:(b + 1) * rows_in], sk,
                      final_out=region_out[b * rows_out:(b + 1) * rows_out])
            finally:
                ops.set_sm_budget(0)
            r.finish(keep_rowvec=True)

        if self.device.type != "cuda":          # host-logic tests (tests/fake_ops.py): same split, no streams
            for b in range(branches):
                run_branch(b)
        else:
            cur = torch.cuda.current_stream()
            streams = [cur] + self._branch_streams[: branches - 1]
            for st in streams[1:]:
                st.wait_stream(cur)             # fork point: BEFORE anything of branch 0 is enqueued on `cur`
            for b, st in enumerate(streams):
                with torch.cuda.stream(st):
                    run_branch(b)
            for st in streams[1:]:
                cur.wait_stream(st)
        # outer tensors whose last consumer was inside the region (the region's input may also be a live skip)
        keep = {t.data_ptr() for t, _ in skips}
        for t, _ in outer + [(hcur, 0)]:
            if t.data_ptr() not in keep:
                keep.add(t.data_ptr())
                main.ar.release(t)
        main.run(prog[i1:], region_out, skips, eps_out=eps_out)
        main.finish()
        return eps_out

    def forward(self, sample: Tensor, timestep, class_labels: Tensor, taps: Optional[dict] = None,
                attn_overrides: Optional[dict] = None, lora_scale: Optional[float] = None) -> Tensor:
        """NCHW fp32/bf16 sample [B,8,H,W], scalar or [B] timestep, class_labels [B,512] -> eps NCHW fp32."""
        nb, c, h, w = sample.shape
        dev = self.device
        if lora_scale is not None:
            self.set_lora_scale(lora_scale)
        x = sample.to(dev, torch.float32).contiguous()
        labels = class_labels.to(dev, torch.float32).contiguous()
        t = torch.as_tensor(timestep, dtype=torch.float32, device=dev).reshape(-1)
        t = t.expand(nb).contiguous() if t.numel() == 1 else t.contiguous()
        xin = torch.zeros(nb, h * w, LATENT_C_PAD, dtype=torch.bfloat16, device=dev)
        ops.pack_nchw_to_nhwc(x, nb, c, h * w, LATENT_C_PAD, xin)
        silu_emb = torch.empty(nb, self.cfg.temb_channels, dtype=torch.bfloat16, device=dev)
        emb = torch.empty(nb, self.cfg.temb_channels, dtype=torch.float32, device=dev) if taps is not None else None
        self.embed(t, None, True, labels, emb, silu_emb)
        if taps is not None:
            taps["emb"] = emb
        eps_nhwc = torch.empty(nb, h * w, self.cfg.out_channels, dtype=torch.float32, device=dev)
        self.forward_nhwc(xin, silu_emb, nb, h, w, eps_nhwc, taps, attn_overrides)
        out = torch.empty(nb, self.cfg.out_channels, h, w, dtype=torch.float32, device=dev)
        ops.unpack_nhwc_to_nchw(eps_nhwc, nb, self.cfg.out_channels, h * w, out)
        return out


class _Runner:
    """Executes program steps for one (sub-)batch on the current stream, allocating from one arena."""

    def __init__(self, eng: UNetEngine, nb: int, h: int, w: int, arena: Arena, silu_emb: Optional[Tensor],
                 taps: Optional[dict], attn_overrides: Optional[dict], rowvec: Optional[Tensor] = None,
                 sms: int = ops.NUM_SMS):
        self.eng, self.nb, self.ar = eng, nb, arena
        self.cfg = eng.cfg
        self.plan = eng._plan(nb, h, w, sms)
        self.max_ctas = 0 if sms == ops.NUM_SMS else sms
        self.W, self.sizes, self.S = self.plan["W"], self.plan["sizes"], eng._small
        self.taps, self.attn_overrides = taps, attn_overrides
        self.own_rowvec = rowvec is None
        if rowvec is None:
            rowvec = arena.alloc((nb, self.plan["temb_total"]), torch.float32)
            ops.conv_gemm(self.W["temb"], silu_emb, 1, nb, 1, rowvec, out_ld=self.plan["temb_total"])
        self.rowvec = rowvec
        self.rowvec_ready = None

    def M(self, lvl: int) -> int:
        return self.nb * self.sizes[lvl][0] * self.sizes[lvl][1]

    def rel(self, t: Optional[Tensor]) -> None:
        """Release `t` if this runner's arena owns it (slices of another arena's tensors are borrowed)."""
        if t is not None and t.data_ptr() in self.ar.live:
            self.ar.release(t)

    def finish(self, keep_rowvec: bool = False) -> None:
        if self.own_rowvec and not keep_rowvec:
            self.ar.release(self.rowvec)
        assert not self.ar.live, f"arena leak: {len(self.ar.live)} buffers"

    # ---- kernels
    def tap(self, name, buf, lvl, c):
        if self.taps is not None:
            hh, ww = self.sizes[lvl]
            self.taps[name] = buf.view(self.nb, hh, ww, c).permute(0, 3, 1, 2).float().clone()

    def gn(self, x0, c0, x1, c1, lvl, name, eps, silu):
        hh, ww = self.sizes[lvl]
        y = self.ar.alloc((self.M(lvl), c0 + c1), torch.bfloat16)
        st0 = self.ar.attached.get(x0.data_ptr())
        st1 = self.ar.attached.get(x1.data_ptr()) if x1 is not None else None
        if st0 is not None and (x1 is None or st1 is not None) and ((c0 + c1) // self.cfg.groups) % 4 == 0:
            # one pass: the producers' epilogues left the partial statistics (b200_conv_gemm_gnstat)
            return ops.groupnorm_apply(x0, c0, st0, x1, c1, st1, self.nb, hh * ww, self.S[name + ".weight"],
                                       self.S[name + ".bias"], eps, silu, y, self.cfg.groups)
        return ops.groupnorm_silu(x0, c0, x1, c1, self.nb, hh * ww, self.S[name + ".weight"], self.S[name + ".bias"], eps,
                                  silu, y, self.cfg.groups)

    def _ws(self, pw, lvl, stride=1, fp32=False):
        if pw.ksplit > 1 and stride == 1 and not fp32:
            return self.ar.alloc((pw.ksplit * self.M(lvl) * pw.n_pad,), torch.float32)
        return None

    def _gn_stat(self, pw, out, lvl, stride=1):
        """Partial-statistics buffer for an output that a GroupNorm will read (attached to `out`: released with it), or
        None where the producing launch cannot leave them (split-K, stride 2, fp32, borrowed output, tiny images)."""
        if not ops.GN_ONEPASS:
            return None
        hh, ww = self.sizes[lvl]
        slabs = ops.gn_stat_slabs(self.nb, hh, ww)
        if (pw.ksplit > 1 or stride != 1 or out.dtype != torch.bfloat16 or slabs <= 0 or pw.n_valid % 4 or pw.block_n % 64
                or pw.geglu or out.data_ptr() not in self.ar.live):
            return None
        st = self.ar.alloc((self.nb, slabs, pw.n_valid // 4, 2), torch.float32)
        self.ar.attach(out, st)
        return st

    def conv(self, name, a0, lvl, *, a1=None, a2=None, rowvec_off=None, residual=None, stride=1, out=None, out_lvl=None,
             gn_next=False):
        pw = self.W[name]
        hh, ww = self.sizes[lvl]
        if out is None:
            out = self.ar.alloc((self.M(out_lvl if out_lvl is not None else lvl), pw.n_valid), torch.bfloat16)
        rv = self.rowvec[:, rowvec_off:] if rowvec_off is not None else None
        ws = self._ws(pw, lvl, stride, out.dtype == torch.float32)
        st = self._gn_stat(pw, out, lvl, stride) if gn_next else None
        ops.conv_gemm(pw, a0, self.nb, hh, ww, out, a1=a1, a2=a2, stride=stride, rowvec=rv,
                      rowvec_ld=self.plan["temb_total"], residual=residual, workspace=ws, max_ctas=self.max_ctas, gn_stat=st)
        self.ar.release(ws)
        return out

    def linear(self, name, a0, lvl, *, a1=None, residual=None, gn_next=False):
        """gn_next: the output is a feature map a GroupNorm reads next (Transformer2DModel.proj_out): run it with the
        level's image geometry so the epilogue can leave per-image statistics (a 1x1 conv is the same GEMM)."""
        pw = self.W[name]
        out = self.ar.alloc((self.M(lvl), pw.n_valid), torch.bfloat16)
        ws = self._ws(pw, lvl)
        st = self._gn_stat(pw, out, lvl) if gn_next else None
        if st is not None:
            hh, ww = self.sizes[lvl]
            ops.conv_gemm(pw, a0, self.nb, hh, ww, out, a1=a1, residual=residual, max_ctas=self.max_ctas, gn_stat=st)
        else:
            ops.conv_gemm(pw, a0, 1, self.M(lvl), 1, out, a1=a1, residual=residual, workspace=ws, max_ctas=self.max_ctas)
        self.ar.release(ws)
        return out

    def linear_ln(self, name, a0, stats, lvl, *, down=None, eps=1e-5):
        pw = self.W[name]
        out = self.ar.alloc((self.M(lvl), pw.n_valid), torch.bfloat16)
        return ops.linear_ln(pw, a0, self.M(lvl), out, stats, eps=eps, down=down, max_ctas=self.max_ctas)

    def linear_with_stats(self, name, a0, lvl, *, down=None, residual=None):
        """A linear layer whose output feeds a LayerNorm-folded GEMM: (output, its per-row chunk statistics)."""
        pw = self.W[name]
        m = self.M(lvl)
        out = self.ar.alloc((m, pw.n_valid), torch.bfloat16)
        stats = self.ar.alloc((m, pw.n_valid // 64, 2), torch.float32)
        T = None
        use_down = down if (down is not None and ops.linear_lora_ok(pw, down)) else None
        if down is not None and use_down is None:
            T = self.linear(name.replace(".to_out", ".lora_down_o"), a0, lvl)
        ops.linear_stats(pw, a0, m, out, stats, a1=T, down=use_down, residual=residual, max_ctas=self.max_ctas)
        self.ar.release(T)
        return out, stats

    def linear_lora(self, name, down, a0, lvl, *, residual=None):
        pw = self.W[name]
        out = self.ar.alloc((self.M(lvl), pw.n_valid), torch.bfloat16)
        return ops.linear_lora(pw, down, a0, self.M(lvl), out, residual=residual, max_ctas=self.max_ctas)

    def resnet(self, r: ResnetDesc, x0, x1, lvl):
        c_h = r.cin - r.skip_c
        n1 = self.gn(x0, c_h, x1, r.skip_c, lvl, r.name + ".norm1", 1e-5, True)
        h1 = self.conv(r.name + ".conv1", n1, lvl, rowvec_off=self.plan["temb_off"][r.name], gn_next=True)
        self.ar.release(n1)
        n2 = self.gn(h1, r.cout, None, 0, lvl, r.name + ".norm2", 1e-5, True)
        self.ar.release(h1)
        if r.has_shortcut:
            out = self.conv(r.name + ".conv2", n2, lvl, a1=x0, a2=x1, gn_next=True)
        else:
            out = self.conv(r.name + ".conv2", n2, lvl, residual=x0, gn_next=True)
        self.ar.release(n2)
        return out

    def attention(self, p, x, lvl, c):
        """x: LayerNorm output [M, c] -> attention output (before to_out)."""
        hh, ww = self.sizes[lvl]
        down = self.W.get(p + ".lora_down_qkv")
        if ops.linear_lora_ok(self.W[p + ".qkv"], down):
            qkv = self.linear_lora(p + ".qkv", down, x, lvl)       # T = x . A^T never leaves the SM
        else:
            T = self.linear(p + ".lora_down_qkv", x, lvl) if down is not None else None
            qkv = self.linear(p + ".qkv", x, lvl, a1=T)
            self.ar.release(T)
        ao = self.ar.alloc((self.M(lvl), c), torch.bfloat16)
        ops.attention(qkv, ao, self.nb, hh * ww, self.cfg.heads, c // self.cfg.heads, variant=self.eng.attn_variant)
        self.ar.release(qkv)
        return ao

    def tfm(self, t: TfmDesc, x, lvl):
        ar, S, W, M = self.ar, self.S, self.W, self.M(lvl)
        c = t.c
        hh, ww = self.sizes[lvl]
        b = t.name + ".transformer_blocks.0"

        def foreign(p_):
            return bool(self.attn_overrides and p_ in self.attn_overrides)

        def fused(p_):          # norm folded into this attention's QKV GEMM?  (its input's producer must be one of ours)
            prev_ok = not foreign(b + ".attn1") if p_.endswith("attn2") else W[t.name + ".proj_in"].ksplit == 1
            return p_ + ".qkv.ln" in W and not foreign(p_) and prev_ok and W[p_ + ".to_out"].ksplit == 1 and c % 64 == 0

        ff_fused = b + ".ff.net.0.proj.ln" in W and not foreign(b + ".attn2") and W[b + ".attn2.to_out"].ksplit == 1 \
            and c % 64 == 0
        n0 = self.gn(x, c, None, 0, lvl, t.name + ".norm", 1e-6, False)
        # every producer of the token stream leaves the row statistics its consumer's folded LayerNorm needs
        stats = None
        if fused(b + ".attn1"):
            tok, stats = self.linear_with_stats(t.name + ".proj_in", n0, lvl)
        else:
            tok = self.linear(t.name + ".proj_in", n0, lvl)
        ar.release(n0)
        order = (("attn1", "norm1"), ("attn2", "norm2"))
        for i, (a, ln_name) in enumerate(order):
            p = f"{b}.{a}"
            next_fused = fused(f"{b}.attn2") if i == 0 else ff_fused
            ln = None
            if stats is not None:
                # LayerNorm folded into the QKV GEMM (and the q/k/v LoRA down-projection computed inside it)
                qkv = self.linear_ln(p + ".qkv.ln", tok, stats, lvl, down=W.get(p + ".lora_down_qkv.ln"))
                ar.release(stats)
                stats = None
                ao = ar.alloc((M, c), torch.bfloat16)
                ops.attention(qkv, ao, self.nb, hh * ww, self.cfg.heads, c // self.cfg.heads, variant=self.eng.attn_variant)
                ar.release(qkv)
            else:
                ln = ar.alloc((M, c), torch.bfloat16)
                ops.layernorm(tok, M, c, S[f"{b}.{ln_name}.weight"], S[f"{b}.{ln_name}.bias"], 1e-5, ln)
                if self.attn_overrides and p in self.attn_overrides:
                    # foreign attention processor installed through the diffusers seam: hand it the
                    # LayerNorm output as a torch tensor, add its result to the residual stream.
                    # (torch glue on purpose: this is not the B200 path.)
                    res = self.attn_overrides[p](ln.view(self.nb, hh * ww, c))
                    new_tok = ar.alloc((M, c), torch.bfloat16)
                    new_tok.copy_((tok.float() + res.reshape(M, c).float()).to(torch.bfloat16))
                    ar.release(ln); ar.release(tok)
                    tok = new_tok
                    continue
                ao = self.attention(p, ln, lvl, c)
                ar.release(ln)
            down_o = W.get(p + ".lora_down_o")
            if next_fused:
                new_tok, stats = self.linear_with_stats(p + ".to_out", ao, lvl, down=down_o, residual=tok)
            elif ops.linear_lora_ok(W[p + ".to_out"], down_o):
                new_tok = self.linear_lora(p + ".to_out", down_o, ao, lvl, residual=tok)
            else:
                To = self.linear(p + ".lora_down_o", ao, lvl) if down_o is not None else None
                new_tok = self.linear(p + ".to_out", ao, lvl, a1=To, residual=tok)
                ar.release(To)
            ar.release(ao); ar.release(tok)
            tok = new_tok
        if stats is not None:
            ffh = self.linear_ln(b + ".ff.net.0.proj.ln", tok, stats, lvl)      # norm3 folded in, GEGLU in the same epilogue
            ar.release(stats)
        else:
            ln = ar.alloc((M, c), torch.bfloat16)
            ops.layernorm(tok, M, c, S[b + ".norm3.weight"], S[b + ".norm3.bias"], 1e-5, ln)
            ffh = self.linear(b + ".ff.net.0.proj", ln, lvl)
            ar.release(ln)
        if t.name + ".ffout_proj" in W:
            out = self.linear(t.name + ".ffout_proj", ffh, lvl, a1=tok, residual=x, gn_next=True)
            ar.release(ffh); ar.release(tok)
            return out
        new_tok = self.linear(b + ".ff.net.2", ffh, lvl, residual=tok)
        ar.release(ffh); ar.release(tok)
        out = self.linear(t.name + ".proj_out", new_tok, lvl, residual=x, gn_next=True)
        ar.release(new_tok)
        return out

    # ---- the walk
    def run(self, steps: List[tuple], hcur: Optional[Tensor], skips: List[Tuple[Tensor, int]], *, xin=None, eps_out=None,
            final_out=None):
        """Execute `steps` from state (hcur, skips); returns the new state.  A tensor is released when its last
        consumer has been enqueued; tensors borrowed from another arena (sub-batch slices) are left alone."""
        cfg = self.cfg
        on_stack = {t.data_ptr() for t, _ in skips}
        last = len(steps) - 1
        for k, st in enumerate(steps):
            kind = st[0]
            if kind == "conv_in":
                hcur = self.conv("conv_in", xin, 0, gn_next=True)
                self.tap("conv_in", hcur, 0, cfg.block_out_channels[0])
            elif kind == "push":
                skips.append((hcur, st[1]))
                on_stack.add(hcur.data_ptr())
            elif kind == "res":
                _, r, lvl, pops = st
                if self.rowvec_ready is not None:          # join the side stream that computes the embedding projections
                    torch.cuda.current_stream().wait_event(self.rowvec_ready)
                    self.rowvec_ready = None
                sk = None
                if pops:
                    sk, _ = skips.pop()
                    on_stack.discard(sk.data_ptr())
                hnew = self.resnet(r, hcur, sk, lvl)
                if hcur.data_ptr() not in on_stack:
                    self.rel(hcur)
                if sk is not None and sk.data_ptr() != hcur.data_ptr():
                    self.rel(sk)
                hcur = hnew
                if not pops:
                    self.tap(r.name, hcur, lvl, r.cout)            # oracle taps: down_blocks.i.resnets.j
            elif kind == "tfm":
                _, t, lvl = st
                hnew = self.tfm(t, hcur, lvl)
                if hcur.data_ptr() not in on_stack:
                    self.rel(hcur)
                hcur = hnew
                self.tap(t.name, hcur, lvl, t.c)                   # oracle taps: down_blocks.i.attentions.j
            elif kind == "tap":
                self.tap(st[1], hcur, st[2], st[3])
            elif kind == "down":
                _, name, lvl = st
                hnew = self.conv(name, hcur, lvl, stride=2, out_lvl=lvl + 1)
                if hcur.data_ptr() not in on_stack:
                    self.rel(hcur)
                hcur = hnew
            elif kind == "up":
                _, name, lvl, c = st
                (hs, ws_), (ho, wo) = self.sizes[lvl], self.sizes[lvl - 1]
                up = self.ar.alloc((self.M(lvl - 1), c), torch.bfloat16)
                ops.upsample_nearest(hcur, self.nb, hs, ws_, c, ho, wo, up)
                self.rel(hcur)
                hcur = self.conv(name, up, lvl - 1, out=final_out if k == last else None, gn_next=True)
                self.ar.release(up)
            elif kind == "out":
                n = self.gn(hcur, cfg.block_out_channels[0], None, 0, 0, "conv_norm_out", 1e-5, True)
                self.rel(hcur)
                self.conv("conv_out", n, 0, out=eps_out)
                self.ar.release(n)
                hcur = None
        return hcur, skips
