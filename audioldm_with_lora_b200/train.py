"""LoRA fine-tuning step of the AudioLDM UNet on the B200 kernels.

Restates the reference's training step body -- /root/reference/script/train/train_audioldm_lora.py:499-565:
    noisy = noise_scheduler.add_noise(latents, noise, t)                    :504
    pred  = unet(noisy, t, encoder_hidden_states=None, class_labels=embeds,
                 cross_attention_kwargs={"scale": 1.0})                     :539-546   (frozen base, unmerged LoRA)
    loss  = F.mse_loss(pred.float(), noise.float(), reduction="mean")       :549
    accelerator.backward(loss)        # autograd + DDP all-reduce (C1)      :557
    optimizer.step()                  # torch.optim.AdamW over LoRA params  :394-403, :563
    lr_scheduler.step()               # get_scheduler("polynomial")         :438-443, :564
-- with hand-written kernels for the forward, the backward (activation gradients everywhere downstream of the first
adapted attention layer; weight gradients only for the rank-r LoRA matrices) and the optimizer:

  * forward: the inference kernels, in forms that keep what the backward needs (attention LSE, GroupNorm statistics,
    the GEGLU pre-activation); activations live in one arena and are released as the backward consumes them;
  * dgrad of every conv / linear layer = b200_conv_gemm with transposed, tap-flipped packed weights; the LoRA branch of
    the backward is again a K segment:  d_x = [dY | dY.(sB)] . [W | A]  ;
  * LoRA weight gradients go straight into ONE flat fp32 arena (layout: per adapter A then B), which is exactly the
    buffer `torch.distributed.all_reduce` (NCCL over NVLink) sums across data-parallel ranks -- no bucket copies;
  * fused multi-tensor AdamW over the flat arena, then one kernel re-materialises the bf16 packed LoRA operands.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops, packing
from .arch import ResnetDesc, TfmDesc
from .engine import LATENT_C_PAD, Arena, UNetEngine
from .ops import PackedWeight, WgradDesc
from .scheduler import DDIMScheduler

Tensor = torch.Tensor
PROJ = ("to_q", "to_k", "to_v", "to_out.0")


@dataclass
class Slot:
    """One adapter inside the flat arenas: A [r, c] at off_a, B [c, r] at off_b."""
    off_a: int
    off_b: int
    r: int
    c: int
    scaling: float        # alpha / r


def polynomial_lr(step: int, lr_init: float, num_training_steps: int, lr_end: float = 1e-7, power: float = 1.0) -> float:
    """diffusers get_scheduler("polynomial", num_warmup_steps=0) multiplier applied to lr_init (train:438-443)."""
    if step > num_training_steps:
        return lr_end
    decay = (1 - step / num_training_steps) ** power
    return (lr_init - lr_end) * decay + lr_end


class LoraTrainer:
    def __init__(self, unet, lr: float = 1.0e-5, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-5, num_training_steps: Optional[int] = None, lora_scale: float = 1.0,
                 process_group=None, scheduler: Optional[DDIMScheduler] = None):
        self.unet = unet
        self.eng: UNetEngine = unet.engine
        self.device = self.eng.device
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.num_training_steps = num_training_steps
        self.lora_scale = float(lora_scale)
        self.pg = process_group
        self.sched = scheduler or DDIMScheduler()
        self.step_count = 0
        if not self.eng.lora:
            raise ValueError("LoraTrainer: the UNet has no LoRA adapters (call add_adapter / get_peft_model first)")
        # ---- flat arenas
        self.slots: Dict[str, Slot] = {}
        off = 0
        for t in self.eng.graph.transformers():
            for a in ("attn1", "attn2"):
                for pr in PROJ:
                    path = f"{t.name}.transformer_blocks.0.{a}.{pr}"
                    e = self.eng.lora.get(path)
                    if e is None:
                        continue
                    r, c = e.A.shape
                    self.slots[path] = Slot(off, off + r * c, r, c, e.alpha / r)
                    off += 2 * r * c
        self.numel = off
        host = torch.zeros(off, dtype=torch.float32)
        for path, s in self.slots.items():
            e = self.eng.lora[path]
            host[s.off_a: s.off_a + s.r * s.c] = e.A.float().reshape(-1)
            host[s.off_b: s.off_b + s.r * s.c] = e.B.float().reshape(-1)
        # (a CPU device only occurs in the host-logic tests, where tests/fake_ops.py stands in for the kernels)
        self.flat_p = host.to(self.device)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=self.device)
        ac = self.sched.alphas_cumprod.to(self.device, torch.float32)
        self.sqrt_ac, self.sqrt_1mac = ac.sqrt().contiguous(), (1 - ac).sqrt().contiguous()
        self._tplans: Dict[Tuple[int, int, int], dict] = {}
        self.arena: Optional[Arena] = None
        self.first_tfm = self._first_adapted_tfm()
        self._fwd_tokens = 0
        self.outstanding_forward: Optional[int] = None      # token of the grad-mode forward whose activations are live

    def begin_forward(self) -> int:
        """Autograd seam: claim the activation arena for a new grad-mode forward.  Activations of a forward whose
        backward never ran are dropped (the arena is reset); that forward's backward then fails loudly."""
        if self.outstanding_forward is not None and self.arena is not None:
            self.arena.live.clear()
            self.arena.free = [(0, self.arena.buf.numel())]
        self._fwd_tokens += 1
        self.outstanding_forward = self._fwd_tokens
        return self._fwd_tokens

    # ------------------------------------------------------------------ bookkeeping
    def _first_adapted_tfm(self) -> str:
        for t in self.eng.graph.transformers():
            if any(p.startswith(t.name + ".") for p in self.slots):
                return t.name
        raise ValueError("no adapted transformer block")

    def param_views(self) -> Dict[str, Tuple[Tensor, Tensor]]:
        """{path: (A [r, c], B [c, r])} views into the flat parameter arena."""
        return {p: (self.flat_p[s.off_a: s.off_a + s.r * s.c].view(s.r, s.c),
                    self.flat_p[s.off_b: s.off_b + s.r * s.c].view(s.c, s.r)) for p, s in self.slots.items()}

    def grad_views(self) -> Dict[str, Tuple[Tensor, Tensor]]:
        return {p: (self.flat_g[s.off_a: s.off_a + s.r * s.c].view(s.r, s.c),
                    self.flat_g[s.off_b: s.off_b + s.r * s.c].view(s.c, s.r)) for p, s in self.slots.items()}

    def sync_to_model(self) -> None:
        """Write the trained adapters back into the model (peft-shaped modules + inference engine)."""
        from .engine import LoraEntry
        ad = {}
        for p, (A, B) in self.param_views().items():
            s = self.slots[p]
            ad[p] = LoraEntry(A.detach().float().cpu().clone(), B.detach().float().cpu().clone(), s.scaling * s.r)
        self.unet._install(ad)

    # ------------------------------------------------------------------ plans (packed backward weights, refresh table)
    def _tplan(self, nb: int, h: int, w: int) -> dict:
        key = (nb, h, w)
        if key not in self._tplans:
            self._tplans[key] = self._build_tplan(nb, h, w)
        return self._tplans[key]

    def _build_tplan(self, nb: int, h: int, w: int) -> dict:
        eng = self.eng
        plan = eng._plan(nb, h, w)
        sd, g, sizes = eng.sd, eng.graph, plan["sizes"]
        W = plan["W"]
        Wb: Dict[str, PackedWeight] = {}

        def conv_tiles(lvl):
            return ops.num_m_tiles(nb, *sizes[lvl])

        def lin_tiles(lvl):
            return math.ceil(nb * sizes[lvl][0] * sizes[lvl][1] / 128)

        def conv3_bwd(name, lvl, c_pad=None):
            # dX = conv(dY, W'), W'[ci, (dh, dw), co] = W[co, ci, 2 - (dh+1), 2 - (dw+1)]: swap in/out, flip both taps
            wt = sd[name + ".weight"].permute(1, 0, 2, 3).flip(2, 3).contiguous()
            wk = packing.conv3x3_to_k(wt, c_pad)
            Wb[name] = eng._pw([wk], None, conv_tiles(lvl), 9, wk.shape[1] // 9)

        def lin_bwd(name, w2d, mt, key=None):
            wt = w2d.t().contiguous()               # [in, out]: N = in, K = out
            Wb[key or name] = eng._pw([wt], None, mt, 1, wt.shape[1])

        lvl_of = plan["lvl_of"]
        for r in g.resnets():
            lvl = lvl_of[r.name]
            conv3_bwd(r.name + ".conv1", lvl)
            conv3_bwd(r.name + ".conv2", lvl)
            if r.has_shortcut:
                lin_bwd(r.name + ".conv_shortcut", sd[r.name + ".conv_shortcut.weight"][:, :, 0, 0], lin_tiles(lvl))
        for i, name in enumerate(g.downsamplers):
            if name:
                conv3_bwd(name, i)
        top = len(g.down) - 1
        for i, name in enumerate(g.upsamplers):
            if name:
                conv3_bwd(name, top - i - 1)
        conv3_bwd("conv_out", 0, LATENT_C_PAD)
        refresh: List[tuple] = []
        for t in g.transformers():
            lvl = lvl_of[t.name]
            mt, c = lin_tiles(lvl), t.c
            b = t.name + ".transformer_blocks.0"
            lin_bwd(t.name + ".proj_in", sd[t.name + ".proj_in.weight"][:, :, 0, 0], mt)
            lin_bwd(t.name + ".proj_out", sd[t.name + ".proj_out.weight"][:, :, 0, 0], mt)
            lin_bwd(b + ".ff.net.0.proj", sd[b + ".ff.net.0.proj.weight"], mt)
            lin_bwd(b + ".ff.net.2", sd[b + ".ff.net.2.weight"], mt)
            # training forward keeps the GEGLU pre-activation: plain (un-fused) ff.net.0.proj
            Wb[b + ".ff.net.0.proj.fwd"] = eng._pw([sd[b + ".ff.net.0.proj.weight"]], sd[b + ".ff.net.0.proj.bias"], mt, 1, c)
            for a in ("attn1", "attn2"):
                p = f"{b}.{a}"
                slots = [self.slots.get(f"{p}.{n}") for n in ("to_q", "to_k", "to_v")]
                wqkv_t = torch.cat([sd[f"{p}.{n}.weight"] for n in ("to_q", "to_k", "to_v")]).t().contiguous()   # [C, 3C]
                if any(s is not None for s in slots):
                    kp = packing.lora_pad(sum(s.r for s in slots if s is not None))
                    fuse = ops.lora_fusion_pays(c, mt, kp)
                    Wb[p + ".bwd_down_qkv"] = packing.pack([torch.zeros(kp, 3 * c)], None, min(64, kp), 1, 3 * c, device=self.device)
                    Wb[p + ".bwd_down_qkv"].lora_rows = sum(s.r for s in slots if s is not None)
                    Wb[p + ".bwd_qkv"] = eng._pw([wqkv_t, torch.zeros(c, kp)], None, mt, 1, 3 * c, kp, split=not fuse,
                                                 max_bn=ops.LORA_FUSED_MAX_BN if fuse else 256)
                    Wb[p + ".bwd_qkv"].lora_fused = fuse
                    off = 0
                    for i, s in enumerate(slots):
                        if s is None:
                            continue
                        sc = s.scaling * self.lora_scale
                        # forward operands (engine plan): A rows of the down-projection, s.B columns of the K segment
                        fw_down, fw_main = W[p + ".lora_down_qkv"].w, W[p + ".qkv"].w
                        refresh.append((fw_down, off, 0, s.off_a, s.r, s.c, 0, 1.0))
                        refresh.append((fw_main, i * c, c + off, s.off_b, s.c, s.r, 0, sc))
                        # backward operands: (s.B)^T rows of the down-projection, A^T columns of the K segment
                        refresh.append((Wb[p + ".bwd_down_qkv"].w, off, i * c, s.off_b, s.c, s.r, 1, sc))
                        refresh.append((Wb[p + ".bwd_qkv"].w, 0, 3 * c + off, s.off_a, s.r, s.c, 1, 1.0))
                        off += s.r
                else:
                    Wb[p + ".bwd_qkv"] = eng._pw([wqkv_t], None, mt, 1, 3 * c)
                so = self.slots.get(f"{p}.to_out.0")
                wo_t = sd[f"{p}.to_out.0.weight"].t().contiguous()
                if so is not None:
                    kp = packing.lora_pad(so.r)
                    sc = so.scaling * self.lora_scale
                    fuse = ops.lora_fusion_pays(c, mt, kp)
                    Wb[p + ".bwd_down_o"] = packing.pack([torch.zeros(kp, c)], None, min(64, kp), 1, c, device=self.device)
                    Wb[p + ".bwd_down_o"].lora_rows = so.r
                    Wb[p + ".bwd_to_out"] = eng._pw([wo_t, torch.zeros(c, kp)], None, mt, 1, c, kp, split=not fuse,
                                                    max_bn=ops.LORA_FUSED_MAX_BN if fuse else 256)
                    Wb[p + ".bwd_to_out"].lora_fused = fuse
                    refresh.append((W[p + ".lora_down_o"].w, 0, 0, so.off_a, so.r, so.c, 0, 1.0))
                    refresh.append((W[p + ".to_out"].w, 0, c, so.off_b, so.c, so.r, 0, sc))
                    refresh.append((Wb[p + ".bwd_down_o"].w, 0, 0, so.off_b, so.c, so.r, 1, sc))
                    refresh.append((Wb[p + ".bwd_to_out"].w, 0, c, so.off_a, so.r, so.c, 1, 1.0))
                else:
                    Wb[p + ".bwd_to_out"] = eng._pw([wo_t], None, mt, 1, c)
        descs = np.zeros(len(refresh), dtype=ops.REFRESH_DTYPE)
        for i, (dst, row, col, src_off, sr, scn, tr, scale) in enumerate(refresh):
            ld = dst.shape[1]
            descs[i] = (dst.data_ptr() + (row * ld + col) * 2, src_off, ld, sr, scn, tr, scale, 0)
        return {"Wb": Wb, "refresh_host": refresh, "refresh_n": len(refresh), "weights_version": eng.weights_version,
                "refresh_dev": torch.from_numpy(descs.view(np.uint8).copy()).to(self.device)}

    def refresh(self, nb: int, h: int, w: int) -> None:
        """bf16 packed LoRA operands (forward + backward) <- flat fp32 parameters: one launch."""
        tp = self._tplan(nb, h, w)
        if tp["weights_version"] != self.eng.weights_version:      # the engine re-packed its weights: pointers moved
            self._tplans.pop((nb, h, w))
            tp = self._tplan(nb, h, w)
        ops.lora_refresh(tp["refresh_dev"], tp["refresh_n"], self.flat_p)

    def _ensure_arena(self, nb: int, h: int, w: int) -> Arena:
        c0 = self.eng.cfg.block_out_channels[0]
        need = nb * h * w * c0 * 2 * 330 + (256 << 20)
        if self.arena is None or self.arena.buf.numel() < need:
            self.arena = Arena(need, self.device)
        return self.arena

    # ------------------------------------------------------------------ forward (keeps the backward's inputs)
    def forward_train(self, xin: Tensor, silu_emb: Tensor, nb: int, h: int, w: int, eps_out: Tensor) -> dict:
        eng, cfg, g = self.eng, self.eng.cfg, self.eng.graph
        plan, tp = eng._plan(nb, h, w), self._tplan(nb, h, w)
        W, Wb, sizes, S = plan["W"], tp["Wb"], plan["sizes"], eng._small
        ar = self._ensure_arena(nb, h, w)
        bf16, f32 = torch.bfloat16, torch.float32
        ctx: dict = {"nb": nb, "h": h, "w": w, "blocks": {}, "skip_idx": {}}

        def M(lvl):
            return nb * sizes[lvl][0] * sizes[lvl][1]

        rowvec = ar.alloc((nb, plan["temb_total"]), f32)
        ops.conv_gemm(W["temb"], silu_emb, 1, nb, 1, rowvec, out_ld=plan["temb_total"])

        def gemm(pw, a0, lvl, *, conv=False, a1=None, a2=None, rowvec_off=None, residual=None, stride=1, out=None,
                 out_lvl=None):
            hh, ww = sizes[lvl]
            m_out = M(out_lvl if out_lvl is not None else lvl)
            if out is None:
                out = ar.alloc((m_out, pw.n_valid), bf16)
            ws = None
            if pw.ksplit > 1 and stride == 1 and out.dtype != f32:
                ws = ar.alloc((pw.ksplit * M(lvl) * pw.n_pad,), f32)
            rv = rowvec[:, rowvec_off:] if rowvec_off is not None else None
            if conv:
                ops.conv_gemm(pw, a0, nb, hh, ww, out, a1=a1, a2=a2, stride=stride, rowvec=rv,
                              rowvec_ld=plan["temb_total"], residual=residual, workspace=ws)
            else:
                ops.conv_gemm(pw, a0, 1, M(lvl), 1, out, a1=a1, residual=residual, workspace=ws)
            ar.release(ws)
            return out

        def gn(x0, c0, x1, c1, lvl, name, eps, silu):
            hh, ww = sizes[lvl]
            y = ar.alloc((M(lvl), c0 + c1), bf16)
            st = ar.alloc((nb, cfg.groups, 2), f32)
            ops.groupnorm_silu_stats(x0, c0, x1, c1, nb, hh * ww, S[name + ".weight"], S[name + ".bias"], eps, silu, y, st,
                                     cfg.groups)
            return y, st

        def resnet(r: ResnetDesc, x0, x1, lvl):
            c_h = r.cin - r.skip_c
            n1, st1 = gn(x0, c_h, x1, r.skip_c, lvl, r.name + ".norm1", 1e-5, True)
            h1 = gemm(W[r.name + ".conv1"], n1, lvl, conv=True, rowvec_off=plan["temb_off"][r.name])
            ar.release(n1)
            n2, st2 = gn(h1, r.cout, None, 0, lvl, r.name + ".norm2", 1e-5, True)
            if r.has_shortcut:
                out = gemm(W[r.name + ".conv2"], n2, lvl, conv=True, a1=x0, a2=x1)
            else:
                out = gemm(W[r.name + ".conv2"], n2, lvl, conv=True, residual=x0)
            ar.release(n2)
            ctx["blocks"][r.name] = {"x0": x0, "x1": x1, "h1": h1, "st1": st1, "st2": st2, "lvl": lvl}
            return out

        def tfm(t: TfmDesc, x, lvl):
            c, m = t.c, M(lvl)
            hh, ww = sizes[lvl]
            b = t.name + ".transformer_blocks.0"
            n0, st0 = gn(x, c, None, 0, lvl, t.name + ".norm", 1e-6, False)
            tok = gemm(W[t.name + ".proj_in"], n0, lvl)
            ar.release(n0)
            bc = {"x": x, "st0": st0, "lvl": lvl, "attn": {}}
            for a, ln_name in (("attn1", "norm1"), ("attn2", "norm2")):
                p = f"{b}.{a}"
                ln = ar.alloc((m, c), bf16)
                ops.layernorm(tok, m, c, S[f"{b}.{ln_name}.weight"], S[f"{b}.{ln_name}.bias"], 1e-5, ln)
                down = W.get(p + ".lora_down_qkv")
                if ops.linear_lora_ok(W[p + ".qkv"], down):      # T = ln . A^T computed inside the QKV GEMM, kept for dB
                    T = ar.alloc((m, 64), bf16)
                    qkv = ar.alloc((m, 3 * c), bf16)
                    ops.linear_lora(W[p + ".qkv"], down, ln, m, qkv, t_out=T)
                else:
                    T = gemm(down, ln, lvl) if down is not None else None
                    qkv = gemm(W[p + ".qkv"], ln, lvl, a1=T)
                ao = ar.alloc((m, c), bf16)
                lse = ar.alloc((nb, cfg.heads, hh * ww), f32)
                ops.attention_lse(qkv, ao, lse, nb, hh * ww, cfg.heads, c // cfg.heads)
                down_o = W.get(p + ".lora_down_o")
                if ops.linear_lora_ok(W[p + ".to_out"], down_o):
                    To = ar.alloc((m, 64), bf16)
                    new_tok = ar.alloc((m, c), bf16)
                    ops.linear_lora(W[p + ".to_out"], down_o, ao, m, new_tok, residual=tok, t_out=To)
                else:
                    To = gemm(down_o, ao, lvl) if down_o is not None else None
                    new_tok = gemm(W[p + ".to_out"], ao, lvl, a1=To, residual=tok)
                bc["attn"][a] = {"tok_in": tok, "ln": ln, "T": T, "qkv": qkv, "ao": ao, "lse": lse, "To": To}
                tok = new_tok
            ln3 = ar.alloc((m, c), bf16)
            ops.layernorm(tok, m, c, S[b + ".norm3.weight"], S[b + ".norm3.bias"], 1e-5, ln3)
            hpre = gemm(Wb[b + ".ff.net.0.proj.fwd"], ln3, lvl)           # [m, 8c] = [values | gates]
            ar.release(ln3)
            ffh = ar.alloc((m, 4 * c), bf16)
            ops.geglu_fwd(hpre, m, 4 * c, ffh)
            new_tok = gemm(W[b + ".ff.net.2"], ffh, lvl, residual=tok)
            ar.release(ffh)
            out = gemm(W[t.name + ".proj_out"], new_tok, lvl, residual=x)
            ar.release(new_tok)
            bc["tok2"], bc["hpre"] = tok, hpre
            ctx["blocks"][t.name] = bc
            return out

        # ---- down path.  Block outputs are kept (they are inputs of the next block's backward and/or skips).
        hcur = gemm(W["conv_in"], xin, 0, conv=True)
        skips: List[Tuple[Tensor, int]] = [(hcur, cfg.block_out_channels[0])]
        order: List[tuple] = []                         # forward order of blocks, for the backward walk
        for i, stages in enumerate(g.down):
            for j, s in enumerate(stages):
                hnew = resnet(s.resnet, hcur, None, i)
                order.append(("res", s.resnet, i))
                if s.tfm:
                    hnew = tfm(s.tfm, hnew, i)
                    order.append(("tfm", s.tfm, i))
                hcur = hnew
                skips.append((hcur, s.resnet.cout))
                order.append(("skip_push", len(skips) - 1, i))
            if g.downsamplers[i]:
                c = cfg.block_out_channels[i]
                hcur = gemm(W[g.downsamplers[i]], hcur, i, conv=True, stride=2, out_lvl=i + 1)
                order.append(("down", g.downsamplers[i], i, c))
                skips.append((hcur, c))
                order.append(("skip_push", len(skips) - 1, i + 1))
        top = len(g.down) - 1
        hcur = resnet(g.mid[0], hcur, None, top); order.append(("res", g.mid[0], top))
        hcur = tfm(g.mid[1], hcur, top); order.append(("tfm", g.mid[1], top))
        hcur = resnet(g.mid[2], hcur, None, top); order.append(("res", g.mid[2], top))
        for i, stages in enumerate(g.up):
            lvl = top - i
            for j, s in enumerate(stages):
                sk, sk_c = skips.pop()
                ctx["skip_idx"][s.resnet.name] = len(skips)        # index of the skip this resnet consumed
                hcur = resnet(s.resnet, hcur, sk, lvl)
                order.append(("res", s.resnet, lvl))
                if s.tfm:
                    hcur = tfm(s.tfm, hcur, lvl)
                    order.append(("tfm", s.tfm, lvl))
            if g.upsamplers[i]:
                c = s.resnet.cout
                (hs, ws_), (ho, wo) = sizes[lvl], sizes[lvl - 1]
                up = ar.alloc((M(lvl - 1), c), bf16)
                ops.upsample_nearest(hcur, nb, hs, ws_, c, ho, wo, up)
                hcur = gemm(W[g.upsamplers[i]], up, lvl - 1, conv=True)
                ar.release(up)
                order.append(("up", g.upsamplers[i], lvl, c))
        n, st = gn(hcur, cfg.block_out_channels[0], None, 0, 0, "conv_norm_out", 1e-5, True)
        gemm(W["conv_out"], n, 0, conv=True, out=eps_out)
        ar.release(n)
        ctx.update(order=order, final_x=hcur, final_st=st, rowvec=rowvec)
        return ctx

    # ------------------------------------------------------------------ backward
    def backward(self, ctx: dict, deps: Tensor) -> None:
        """deps: bf16 [nb*h*w, 64] = d loss / d eps (columns >= 8 zero).  Accumulates LoRA gradients into flat_g and
        releases every activation kept by forward_train."""
        eng, cfg, g = self.eng, self.eng.cfg, self.eng.graph
        nb, h, w = ctx["nb"], ctx["h"], ctx["w"]
        plan, tp = eng._plan(nb, h, w), self._tplan(nb, h, w)
        Wb, sizes, S = tp["Wb"], plan["sizes"], eng._small
        ar = self.arena
        bf16, f32 = torch.bfloat16, torch.float32
        blocks = ctx["blocks"]
        flat_g = self.flat_g

        def M(lvl):
            return nb * sizes[lvl][0] * sizes[lvl][1]

        def gemm(pw, a0, lvl, *, conv=False, a1=None, residual=None):
            hh, ww = sizes[lvl]
            out = ar.alloc((M(lvl), pw.n_valid), bf16)
            ws = ar.alloc((pw.ksplit * M(lvl) * pw.n_pad,), f32) if pw.ksplit > 1 else None
            if conv:
                ops.conv_gemm(pw, a0, nb, hh, ww, out, a1=a1, residual=residual, workspace=ws)
            else:
                ops.conv_gemm(pw, a0, 1, M(lvl), 1, out, a1=a1, residual=residual, workspace=ws)
            ar.release(ws)
            return out

        def gn_bwd(x0, c0, x1, c1, lvl, name, st, silu, dy, dres, res_ld, need_dx1=True):
            hh, ww = sizes[lvl]
            dx0 = ar.alloc((M(lvl), c0), bf16)
            dx1 = ar.alloc((M(lvl), c1), bf16) if (c1 and need_dx1) else None
            ops.groupnorm_silu_bwd(x0, c0, x1, c1, nb, hh * ww, S[name + ".weight"], S[name + ".bias"], st, silu, dy, dres,
                                   res_ld, dx0, dx1, cfg.groups)
            return dx0, dx1

        def res_bwd(r: ResnetDesc, d_out, lvl, need_dx=True, need_dskip=True):
            bc = blocks.pop(r.name)
            c_h = r.cin - r.skip_c
            if not need_dx:
                for k in ("h1", "st1", "st2"):
                    ar.release(bc[k])
                ar.release(d_out)
                return None, None
            d_n2 = gemm(Wb[r.name + ".conv2"], d_out, lvl, conv=True)
            d_h1, _ = gn_bwd(bc["h1"], r.cout, None, 0, lvl, r.name + ".norm2", bc["st2"], True, d_n2, None, 0)
            ar.release(d_n2); ar.release(bc["h1"]); ar.release(bc["st2"])
            d_n1 = gemm(Wb[r.name + ".conv1"], d_h1, lvl, conv=True)
            ar.release(d_h1)
            if r.has_shortcut:
                dres = gemm(Wb[r.name + ".conv_shortcut"], d_out, lvl)
                ar.release(d_out)
            else:
                dres = d_out
            d_x0, d_x1 = gn_bwd(bc["x0"], c_h, bc["x1"], r.skip_c, lvl, r.name + ".norm1", bc["st1"], True, d_n1, dres, r.cin,
                                need_dx1=need_dskip)
            ar.release(d_n1); ar.release(dres); ar.release(bc["st1"])
            return d_x0, d_x1

        def tfm_bwd(t: TfmDesc, d_out, lvl, need_dx=True):
            bc = blocks.pop(t.name)
            c, m = t.c, M(lvl)
            hh, ww = sizes[lvl]
            b = t.name + ".transformer_blocks.0"
            d_newtok = gemm(Wb[t.name + ".proj_out"], d_out, lvl)
            d_ffh = gemm(Wb[b + ".ff.net.2"], d_newtok, lvl)
            d_h = ar.alloc((m, 8 * c), bf16)
            ops.geglu_bwd(bc["hpre"], d_ffh, m, 4 * c, d_h)
            ar.release(d_ffh); ar.release(bc["hpre"])
            d_ln3 = gemm(Wb[b + ".ff.net.0.proj"], d_h, lvl)
            ar.release(d_h)
            d_tok = ar.alloc((m, c), bf16)
            ops.layernorm_bwd(bc["tok2"], d_ln3, m, c, S[b + ".norm3.weight"], 1e-5, d_newtok, d_tok)
            ar.release(d_ln3); ar.release(d_newtok); ar.release(bc["tok2"])
            for a, ln_name in (("attn2", "norm2"), ("attn1", "norm1")):
                p = f"{b}.{a}"
                ac = bc["attn"][a]
                so = self.slots.get(p + ".to_out.0")
                sl = [self.slots.get(f"{p}.{n}") for n in ("to_q", "to_k", "to_v")]
                descs: List[WgradDesc] = []
                if so is not None and ops.linear_lora_ok(Wb[p + ".bwd_to_out"], Wb[p + ".bwd_down_o"]):
                    dTo = ar.alloc((m, 64), bf16)                # dT = d_tok . (s B) computed inside the dgrad GEMM
                    d_ao = ar.alloc((m, c), bf16)
                    ops.linear_lora(Wb[p + ".bwd_to_out"], Wb[p + ".bwd_down_o"], d_tok, m, d_ao, t_out=dTo)
                else:
                    dTo = gemm(Wb[p + ".bwd_down_o"], d_tok, lvl) if so is not None else None
                    d_ao = gemm(Wb[p + ".bwd_to_out"], d_tok, lvl, a1=dTo)
                if so is not None:
                    sc = so.scaling * self.lora_scale
                    descs.append(WgradDesc(d_tok.data_ptr(), ac["To"].data_ptr(), flat_g.data_ptr() + 4 * so.off_b,
                                           c, ac["To"].shape[1], c, so.r, so.r, 1, sc))
                    descs.append(WgradDesc(ac["ao"].data_ptr(), dTo.data_ptr(), flat_g.data_ptr() + 4 * so.off_a,
                                           c, dTo.shape[1], c, so.r, 1, c, 1.0))
                d_qkv = ar.alloc((m, 3 * c), bf16)
                delta = ar.alloc((nb, cfg.heads, hh * ww), f32)
                ops.attention_bwd(ac["qkv"], ac["ao"], d_ao, ac["lse"], delta, d_qkv, nb, hh * ww, cfg.heads, c // cfg.heads)
                ar.release(delta); ar.release(d_ao)
                has_qkv = any(s is not None for s in sl)
                stop = (not need_dx) and a == "attn1"
                d_ln = None
                if has_qkv and not stop and ops.linear_lora_ok(Wb[p + ".bwd_qkv"], Wb[p + ".bwd_down_qkv"]):
                    dT = ar.alloc((m, 64), bf16)
                    d_ln = ar.alloc((m, c), bf16)
                    ops.linear_lora(Wb[p + ".bwd_qkv"], Wb[p + ".bwd_down_qkv"], d_qkv, m, d_ln, t_out=dT)
                else:
                    dT = gemm(Wb[p + ".bwd_down_qkv"], d_qkv, lvl) if has_qkv else None
                off = 0
                for i, s in enumerate(sl):
                    if s is None:
                        continue
                    sc = s.scaling * self.lora_scale
                    ldt = ac["T"].shape[1]
                    descs.append(WgradDesc(d_qkv.data_ptr() + 2 * i * c, ac["T"].data_ptr() + 2 * off,
                                           flat_g.data_ptr() + 4 * s.off_b, 3 * c, ldt, c, s.r, s.r, 1, sc))
                    descs.append(WgradDesc(ac["ln"].data_ptr(), dT.data_ptr() + 2 * off, flat_g.data_ptr() + 4 * s.off_a,
                                           c, dT.shape[1], c, s.r, 1, c, 1.0))
                    off += s.r
                if descs:
                    ops.lora_wgrad(descs, m)
                if not stop:
                    if d_ln is None:
                        d_ln = gemm(Wb[p + ".bwd_qkv"], d_qkv, lvl, a1=dT)
                    new_d_tok = ar.alloc((m, c), bf16)
                    ops.layernorm_bwd(ac["tok_in"], d_ln, m, c, S[f"{b}.{ln_name}.weight"], 1e-5, d_tok, new_d_tok)
                    ar.release(d_ln)
                else:
                    new_d_tok = None
                for k in ("ln", "T", "qkv", "ao", "lse", "To", "tok_in"):
                    ar.release(ac[k])
                ar.release(dTo); ar.release(dT); ar.release(d_qkv); ar.release(d_tok)
                d_tok = new_d_tok
            if not need_dx:
                ar.release(bc["st0"]); ar.release(d_out)
                return None
            d_n0 = gemm(Wb[t.name + ".proj_in"], d_tok, lvl)
            ar.release(d_tok)
            d_x, _ = gn_bwd(bc["x"], c, None, 0, lvl, t.name + ".norm", bc["st0"], False, d_n0, d_out, c)
            ar.release(d_n0); ar.release(d_out); ar.release(bc["st0"])
            return d_x

        # ---- output head
        d_n = gemm(Wb["conv_out"], deps, 0, conv=True)
        d, _ = gn_bwd(ctx["final_x"], cfg.block_out_channels[0], None, 0, 0, "conv_norm_out", ctx["final_st"], True, d_n, None, 0)
        ar.release(d_n); ar.release(ctx["final_st"])
        # ---- reverse walk.  `d` is the gradient w.r.t. the current block's output.
        order = ctx["order"]
        first = next(i for i, o in enumerate(order) if o[0] == "tfm" and o[1].name == self.first_tfm)
        first_skip = self._first_skip_needing_grad(order, first)
        skip_grads: Dict[int, Optional[Tensor]] = {}
        for idx in range(len(order) - 1, first - 1, -1):
            o = order[idx]
            kind = o[0]
            if kind == "up":
                _, name, lvl, c = o
                (hs, ws_), (ho, wo) = sizes[lvl], sizes[lvl - 1]
                d_up = gemm(Wb[name], d, lvl - 1, conv=True)
                ar.release(d)
                d = ar.alloc((M(lvl), c), bf16)
                ops.upsample_nearest_bwd(d_up, nb, hs, ws_, c, ho, wo, d)
                ar.release(d_up)
            elif kind == "tfm":
                _, t, lvl = o
                d = tfm_bwd(t, d, lvl, need_dx=idx > first)
            elif kind == "res":
                _, r, lvl = o
                if r.skip_c:
                    k = ctx["skip_idx"][r.name]
                    d, dsk = res_bwd(r, d, lvl, need_dskip=k >= first_skip)
                    skip_grads[k] = dsk
                else:
                    d, _ = res_bwd(r, d, lvl)
            elif kind == "skip_push":
                _, k, lvl = o
                dsk = skip_grads.pop(k, None)
                if dsk is not None:
                    ops.add_bf16(d, dsk)
                    ar.release(dsk)
            elif kind == "down":
                _, name, i, c = o
                hh, ww = sizes[i]
                z = ar.alloc((M(i), c), bf16)
                ops.zero_insert(d, nb, hh, ww, c, z)
                ar.release(d)
                d = gemm(Wb[name], z, i, conv=True)
                ar.release(z)
        # Everything upstream of the first adapted block has no trainable parameter: nothing to compute.  Block
        # inputs / outputs (owned by the walk, not by a block context) are reclaimed wholesale: the arena is reset.
        blocks.clear()
        self.arena_peak = ar.peak
        ar.live.clear()
        ar.free = [(0, ar.buf.numel())]
        self.outstanding_forward = None

    @staticmethod
    def _first_skip_needing_grad(order, first: int) -> int:
        """Index of the first skip-stack entry produced at or after the first adapted block."""
        for o in order[first:]:
            if o[0] == "skip_push":
                return o[1]
        return 1 << 30

    # ------------------------------------------------------------------ the step
    def forward_backward(self, noisy: Tensor, timesteps: Tensor, prompt_embeds: Tensor, noise: Tensor) -> Tensor:
        """noisy / noise NCHW fp32 [B, 8, H, W], timesteps [B], prompt_embeds [B, 512].  Adds the LoRA gradients of
        mean((unet(noisy) - noise)^2) to flat_g and returns the loss (0-dim device tensor, no host sync)."""
        eng = self.eng
        nb, c, h, w = noisy.shape
        dev = self.device
        self.refresh(nb, h, w)
        xin = torch.zeros(nb, h * w, LATENT_C_PAD, dtype=torch.bfloat16, device=dev)
        ops.pack_nchw_to_nhwc(noisy.contiguous(), nb, c, h * w, LATENT_C_PAD, xin)
        silu_emb = torch.empty(nb, eng.cfg.temb_channels, dtype=torch.bfloat16, device=dev)
        eng.embed(timesteps.to(dev, torch.float32).contiguous(), None, True, prompt_embeds.to(dev, torch.float32).contiguous(),
                  None, silu_emb)
        eps = torch.empty(nb, h * w, eng.cfg.out_channels, dtype=torch.float32, device=dev)
        ctx = self.forward_train(xin, silu_emb, nb, h, w, eps)
        count = nb * eng.cfg.out_channels * h * w
        deps = self.arena.alloc((nb * h * w, LATENT_C_PAD), torch.bfloat16)
        self.loss_sum.zero_()
        ops.mse_grad(eps, noise.contiguous(), nb, h * w, LATENT_C_PAD, 1.0 / count, self.loss_sum, deps)
        self.last_pred_nhwc = eps
        self.backward(ctx, deps)
        return (self.loss_sum / count).squeeze(0)

    def current_lr(self) -> float:
        if self.num_training_steps is None:
            return self.lr
        return polynomial_lr(self.step_count, self.lr, self.num_training_steps)

    def allreduce_grads(self) -> float:
        """DDP semantics (C1): sum the flat LoRA-gradient arena over the data-parallel group; the 1/world
        average is folded into the optimizer kernel's grad_scale."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.pg)
            return 1.0 / dist.get_world_size(self.pg)
        return 1.0

    def optimizer_step(self, grad_scale: float = 1.0) -> None:
        lr = self.current_lr()                 # LambdaLR: optimizer step k+1 runs with lambda(k)
        self.step_count += 1
        ops.adamw_flat(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr, self.betas[0], self.betas[1], self.eps,
                       self.weight_decay, self.step_count, grad_scale)

    # ------------------------------------------------------------------ the step as a replayed CUDA graph
    def capture(self, nb: int, h: int, w: int = 16) -> None:
        """Capture {LoRA refresh, add_noise, forward, loss, backward} for a fixed batch shape into one CUDA graph (about
        1000 kernel launches).  Inputs are staged into static device buffers.  The gradient all-reduce (NCCL) and the
        fused AdamW launch stay outside the graph: two host-side launches per step, and the collective never runs under
        stream capture."""
        dev, f32 = self.device, torch.float32
        c = self.eng.cfg.in_channels
        self._g_shape = (nb, h, w)
        self._g_lat = torch.zeros(nb, c, h, w, dtype=f32, device=dev)
        self._g_noise = torch.zeros_like(self._g_lat)
        self._g_t = torch.zeros(nb, dtype=torch.long, device=dev)
        self._g_emb = torch.zeros(nb, self.eng.cfg.class_in_dim, dtype=f32, device=dev)
        self._g_loss = torch.zeros((), dtype=f32, device=dev)

        def body():
            noisy = torch.empty_like(self._g_lat)
            ops.add_noise(self._g_lat, self._g_noise, self.sqrt_ac[self._g_t].contiguous(),
                          self.sqrt_1mac[self._g_t].contiguous(), noisy)
            self.flat_g.zero_()
            self._g_loss.copy_(self.forward_backward(noisy, self._g_t, self._g_emb, self._g_noise))

        # warm-up outside capture: packs weights, sizes the arena, sets function attributes (no optimizer step)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            body()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        self._graph = g
        # the graph holds raw pointers into the engine's packed weights and this trainer's refresh table
        self._g_version = self.eng.weights_version

    def train_step_graphed(self, latents: Tensor, noise: Tensor, timesteps: Tensor, prompt_embeds: Tensor) -> Tensor:
        """train_step with the forward/backward replayed from a CUDA graph.  Returns the loss buffer (a static device
        tensor that the next step overwrites)."""
        nb, _, h, w = latents.shape
        if getattr(self, "_graph", None) is None or self._g_shape != (nb, h, w) or \
                self._g_version != self.eng.weights_version:
            self._graph = None
            self.capture(nb, h, w)
        self._g_lat.copy_(latents, non_blocking=True)
        self._g_noise.copy_(noise, non_blocking=True)
        self._g_t.copy_(timesteps, non_blocking=True)
        self._g_emb.copy_(prompt_embeds, non_blocking=True)
        self._graph.replay()
        self.optimizer_step(self.allreduce_grads())
        return self._g_loss

    def train_step(self, latents: Tensor, noise: Tensor, timesteps: Tensor, prompt_embeds: Tensor) -> Tensor:
        """One optimizer step on this rank's batch (train_audioldm_lora.py:499-565 with synthetic latents / embeddings)."""
        dev = self.device
        t = timesteps.to(dev).long()
        x0, nz = latents.to(dev, torch.float32).contiguous(), noise.to(dev, torch.float32).contiguous()
        noisy = torch.empty_like(x0)
        ops.add_noise(x0, nz, self.sqrt_ac[t].contiguous(), self.sqrt_1mac[t].contiguous(), noisy)
        self.flat_g.zero_()
        loss = self.forward_backward(noisy, t, prompt_embeds, nz)
        scale = self.allreduce_grads()
        self.optimizer_step(scale)
        return loss
