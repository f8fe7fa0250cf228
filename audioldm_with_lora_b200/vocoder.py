"""HiFi-GAN vocoder (transformers `SpeechT5HifiGan`) on the sm_100a kernels.

The reference loads the vocoder at /root/reference/script/train/train_audioldm_lora.py:371 and runs it as the last stage of
`AudioLDMPipeline.__call__` (`mel_spectrogram_to_waveform`, /root/reference/app.py:14, generate_audio.py:47-52): log-mel
[B, T, 64] -> waveform [B, 160 T].  Architecture (cvssp/audioldm-s-full-v2 vocoder/config.json): conv_pre (64 -> 1024, k 7),
five stages of {LeakyReLU(0.1), ConvTranspose1d (rates 5, 4, 2, 2, 2; kernels 16, 16, 8, 4, 4; channels halve), mean of three
residual blocks (kernels 3 / 7 / 11, dilations 1 / 3 / 5: x += conv2(lrelu(conv1(lrelu(x)))), three times)}, LeakyReLU(0.01),
conv_post (32 -> 1, k 7), tanh.

Every convolution is one `ops.conv1d` launch (the tcgen05 implicit-GEMM kernel with a 1-D tap walk over time-major
[B, L, C] bf16 activations; TMA zero fill = the convolution's zero padding); a transposed convolution is one launch per
output phase (stride s -> s launches, each a k/s-tap convolution writing every s-th output row).  Layout conventions:
  * activations are stored POST-LeakyReLU: y = lrelu(x, 0.1) is what the next convolution reads; where the pre-activation
    x is also needed (the residual add) the kernel recovers it on the fly, x = min(y, 10 y) -- exact up to bf16 rounding,
    which both forms incur once;
  * channel counts below 64 (the last stage's 32) are stored in 64 channels with zero weights / zero activations, since a
    K block of the kernel is 64 channels;
  * the mean over the three residual blocks and the activation in front of the next layer are one elementwise kernel
    (`ops.lrelu_mean3`).

`B200HifiGan` has the `__call__(mel)` / `.config` surface `AudioLDMPipeline` uses, so it is a drop-in for the torch module
(which stays the reference-path implementation and the parity partner: it IS the reference's own vocoder code).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from . import ops, packing

Tensor = torch.Tensor


def _cpad(c: int) -> int:
    return (c + 63) // 64 * 64


def _conv1d_to_k(w: Tensor, co_pad: int) -> Tensor:
    """nn.Conv1d weight [Co, Ci, k] -> [co_pad, k * Ci_pad], K index = tap * Ci_pad + ci."""
    co, ci, k = w.shape
    out = w.new_zeros(co_pad, k, _cpad(ci))
    out[:co, :, :ci] = w.permute(0, 2, 1)
    return out.reshape(co_pad, k * _cpad(ci))


def _convT_phase_to_k(w: Tensor, stride: int, a: int, ntaps: int, co_pad: int) -> Tensor:
    """nn.ConvTranspose1d weight [Ci, Co, k], output phase with kernel offset a: taps W[:, :, stride * j + a]^T."""
    ci, co, k = w.shape
    out = w.new_zeros(co_pad, ntaps, _cpad(ci))
    for j in range(ntaps):
        kk = stride * j + a
        if kk < k:
            out[:co, j, :ci] = w[:, :, kk].t()
    return out.reshape(co_pad, ntaps * _cpad(ci))


class B200HifiGan:
    def __init__(self, vocoder, device="cuda"):
        cfg = vocoder.config
        if getattr(cfg, "normalize_before", False):
            raise NotImplementedError("normalize_before=True (AudioLDM's vocoder config sets it False)")
        self.config = cfg
        self.device = torch.device(device)
        self.slope = float(cfg.leaky_relu_slope)
        self.rates = list(cfg.upsample_rates)
        self.up_kernels = list(cfg.upsample_kernel_sizes)
        self.rb_kernels = list(cfg.resblock_kernel_sizes)
        self.rb_dilations = [list(d) for d in cfg.resblock_dilation_sizes]
        self.c0 = int(cfg.upsample_initial_channel)
        self.in_dim = int(cfg.model_in_dim)
        sd = {k: v.detach().float().cpu() for k, v in vocoder.state_dict().items()}
        if any(k.endswith("weight_g") or "parametrizations" in k for k in sd):
            raise ValueError("remove weight norm first (SpeechT5HifiGan.remove_weight_norm())")
        self.sd = sd
        self._plans: Dict[Tuple[int, int], dict] = {}
        self._bufs: Dict[tuple, Tensor] = {}

    # torch-module surface used by AudioLDMPipeline
    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    # ------------------------------------------------------------------ weights
    def _pw(self, wk: Tensor, bias: Tensor, co_pad: int, m_tiles: int, ntaps: int, cin_pad: int, block_n=None):
        b = torch.zeros(co_pad)
        b[: bias.numel()] = bias
        bn = block_n or ops.choose_tiling(co_pad, m_tiles, ntaps * cin_pad // 64, allow_split=False)[0]
        return packing.pack([wk], b, bn, ntaps, cin_pad, device=self.device)

    def _plan(self, nb: int, t: int) -> dict:
        key = (nb, t)
        if key in self._plans:
            return self._plans[key]
        sd, W = self.sd, {}

        def tiles(length):
            return nb * math.ceil(length / 128)

        cp = _cpad(self.c0)
        W["conv_pre"] = self._pw(_conv1d_to_k(sd["conv_pre.weight"], cp), sd["conv_pre.bias"], cp, tiles(t), 7, _cpad(self.in_dim))
        length, cin = t, self.c0
        stages: List[dict] = []
        for i, (s, k) in enumerate(zip(self.rates, self.up_kernels)):
            cout = self.c0 // (2 ** (i + 1))
            cop = _cpad(cout)
            pad = (k - s) // 2
            l_out = (length - 1) * s - 2 * pad + k
            w_up, b_up = sd[f"upsampler.{i}.weight"], sd[f"upsampler.{i}.bias"]
            phases = []
            for phi in range(s):
                a, b = (phi + pad) % s, (phi + pad) // s
                ntaps = (k - a + s - 1) // s
                rows = (l_out - phi + s - 1) // s
                phases.append({"pw": self._pw(_convT_phase_to_k(w_up, s, a, ntaps, cop), b_up, cop, tiles(rows), ntaps, _cpad(cin)),
                               "dh0": b, "rows": rows, "phi": phi})
            blocks = []
            for j, (rk, dils) in enumerate(zip(self.rb_kernels, self.rb_dilations)):
                rb = f"resblocks.{i * len(self.rb_kernels) + j}"
                pairs = []
                for q, d in enumerate(dils):
                    w1, b1 = sd[f"{rb}.convs1.{q}.weight"], sd[f"{rb}.convs1.{q}.bias"]
                    w2, b2 = sd[f"{rb}.convs2.{q}.weight"], sd[f"{rb}.convs2.{q}.bias"]
                    pairs.append({"c1": self._pw(_conv1d_to_k(w1, cop), b1, cop, tiles(l_out), rk, cop), "d": d,
                                  "c2": self._pw(_conv1d_to_k(w2, cop), b2, cop, tiles(l_out), rk, cop), "k": rk})
                blocks.append(pairs)
            stages.append({"s": s, "cin": _cpad(cin), "c": cop, "len_in": length, "len": l_out, "phases": phases, "blocks": blocks})
            length, cin = l_out, cout
        # conv_post: one output channel -> 8 stored fp32 columns (column 0 is the waveform)
        W["conv_post"] = self._pw(_conv1d_to_k(sd["conv_post.weight"], 8), sd["conv_post.bias"], 8, tiles(length), 7, _cpad(cin),
                                  block_n=32)
        plan = {"W": W, "stages": stages, "len_out": length}
        self._plans[key] = plan
        return plan

    def _buf(self, name: str, shape, dtype=torch.bfloat16) -> Tensor:
        key = (name, tuple(shape), dtype)
        if key not in self._bufs:
            self._bufs[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return self._bufs[key]

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def __call__(self, spectrogram: Tensor) -> Tensor:
        """log-mel [B, T, model_in_dim] (or [T, model_in_dim]) -> waveform [B, T * prod(rates)] fp32 (or 1-D)."""
        batched = spectrogram.dim() == 3
        mel = spectrogram if batched else spectrogram.unsqueeze(0)
        # (no CPU fallback: ops -> _lib.ptr() refuses CPU tensors)
        nb, t, f = mel.shape
        if f != self.in_dim:
            raise ValueError(f"expected {self.in_dim} mel bins, got {f}")
        plan = self._plan(nb, t)
        W, sl = plan["W"], self.slope
        fp = _cpad(f)
        if fp == f and mel.dtype == torch.bfloat16 and mel.is_contiguous():
            x = mel
        else:
            x = self._buf("mel", (nb, t, fp))
            if fp == f and mel.dtype == torch.float32 and mel.is_contiguous():
                ops.f32_to_bf16(mel, x)
            else:
                x.zero_()
                x[:, :, :f] = mel
        # conv_pre, stored post-LeakyReLU(slope) for the first upsampler
        h = self._buf("pre", (nb, t, _cpad(self.c0)))
        ops.conv1d(W["conv_pre"], x, nb, t, h, dh0=-3, dh_step=1, act_slope=sl)
        nstage = len(plan["stages"])
        for i, st in enumerate(plan["stages"]):
            c, length, s = st["c"], st["len"], st["s"]
            u = self._buf(f"u{i}", (nb, length, c))            # lrelu(upsampled, slope): input AND residual of the three blocks
            for ph in st["phases"]:
                ops.conv1d(ph["pw"], h, nb, st["len_in"], u.view(-1)[ph["phi"] * c:], dh0=ph["dh0"], dh_step=-1,
                           m_rows=ph["rows"], act_slope=sl, out_ld=s * c, out_batch_stride=length * c)
            tmp = self._buf(f"t{i}", (nb, length, c))
            ping = [self._buf(f"p{i}.{z}", (nb, length, c)) for z in range(2)]
            finals = [self._buf(f"r{i}.{j}", (nb, length, c)) for j in range(len(st["blocks"]))]
            for j, pairs in enumerate(st["blocks"]):
                cur = u
                for q, pr in enumerate(pairs):
                    k, d = pr["k"], pr["d"]
                    ops.conv1d(pr["c1"], cur, nb, length, tmp, dh0=-d * (k - 1) // 2, dh_step=d, act_slope=sl)
                    dst = finals[j] if q == len(pairs) - 1 else ping[q & 1]
                    ops.conv1d(pr["c2"], tmp, nb, length, dst, dh0=-(k - 1) // 2, dh_step=1, residual=cur, res_slope=sl,
                               act_slope=sl)
                    cur = dst
            # mean of the blocks, then the activation in front of the next layer (the last one: F.leaky_relu's default 0.01)
            h = self._buf(f"h{i}", (nb, length, c))
            assert len(finals) == 3, "the mean kernel takes the three residual blocks of the AudioLDM vocoder"
            ops.lrelu_mean3(finals[0], finals[1], finals[2], sl, sl if i != nstage - 1 else 0.01, h)
        length = plan["len_out"]
        out8 = self._buf("out8", (nb * length, 8), torch.float32)
        ops.conv1d(W["conv_post"], h, nb, length, out8, dh0=-3, dh_step=1, act_tanh=True, out_ld=8)
        wave = out8.view(nb, length, 8)[..., 0].clone()
        return wave if batched else wave.reshape(-1)


def from_torch_vocoder(vocoder, device="cuda") -> B200HifiGan:
    """Build the B200 vocoder from transformers' `SpeechT5HifiGan` (weight norm removed, as the hub checkpoint is)."""
    return B200HifiGan(vocoder, device=device)
