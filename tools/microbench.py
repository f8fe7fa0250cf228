#!/usr/bin/env python
"""Kernel-only timings (CUDA-graph-captured back-to-back launches) of single ops.
    python tools/microbench.py gn|ln|attn|gemm [...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops, packing  # noqa: E402
from tools.layer_profile import time_fn  # noqa: E402

g = torch.Generator().manual_seed(0)
bf16 = torch.bfloat16


def gn():
    for (nb, hw, c) in [(16, 4000, 128), (16, 4000, 256), (16, 1000, 256), (16, 252, 384), (16, 64, 640), (16, 64, 1280)]:
        x = torch.randn(nb, hw, c, generator=g).to("cuda", bf16)
        y = torch.empty_like(x)
        gm, bt = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        ms = time_fn(lambda: ops.groupnorm_silu(x, c, None, 0, nb, hw, gm, bt, 1e-5, True, y))
        print(f"groupnorm nb{nb} hw{hw} c{c}: {ms * 1e3:.2f} us  {nb * hw * c * 4 / ms / 1e6:.0f} GB/s (r+w)")


def ln():
    for (m, c) in [(16000, 256), (4032, 384), (1024, 640)]:
        x = torch.randn(m, c, generator=g).to("cuda", bf16)
        y = torch.empty_like(x)
        gm, bt = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        ms = time_fn(lambda: ops.layernorm(x, m, c, gm, bt, 1e-5, y))
        print(f"layernorm m{m} c{c}: {ms * 1e3:.2f} us")


def attn():
    for (b, s, d) in [(16, 1000, 32), (16, 252, 48), (16, 64, 80), (32, 3000, 64)]:
        qkv = torch.randn(b, s, 3 * 8 * d, generator=g).to("cuda", bf16)
        o = torch.empty(b, s, 8 * d, dtype=bf16, device="cuda")
        ms = time_fn(lambda: ops.attention(qkv, o, b, s, 8, d), reps=10)
        print(f"attention b{b} s{s} d{d}: {ms * 1e3:.2f} us  {4.0 * b * 8 * s * s * d / ms / 1e9:.0f} TFLOP/s")


def gemm():
    shapes = [("conv L0 128->128", 16, 250, 16, 128, 128, 9), ("conv L1 256->256", 16, 125, 8, 256, 256, 9),
              ("conv L2 384->384", 16, 63, 4, 384, 384, 9), ("conv L3 640->640", 16, 32, 2, 640, 640, 9),
              ("conv L3 1280->640", 16, 32, 2, 1280, 640, 9), ("lin L1 qkv 256->768", 1, 16000, 1, 256, 768, 1),
              ("lin L1 ff2 1024->256", 1, 16000, 1, 1024, 256, 1), ("lin L3 640->640", 1, 1024, 1, 640, 640, 1),
              ("lin L3 lora 640->64", 1, 1024, 1, 640, 64, 1)]
    import os
    flt = os.environ.get("MB_FILTER", "")
    for label, nb, hh, ww, ci, co, taps in shapes:
        if flt and flt not in label:
            continue
        x = torch.randn(nb * hh * ww, ci, generator=g).to("cuda", bf16)
        wt = torch.randn(co, taps * ci, generator=g) * (taps * ci) ** -0.5
        out = torch.empty(nb * hh * ww, co, dtype=bf16, device="cuda")
        flops = 2.0 * nb * hh * ww * co * taps * ci
        import math
        bn = ops.choose_block_n(co, ops.num_m_tiles(nb, hh, ww))
        pw = packing.pack([wt], torch.zeros(co), bn, taps, ci, device="cuda")
        ms = time_fn(lambda: ops.conv_gemm(pw, x, nb, hh, ww, out))
        print(f"{label} bn={bn}: {ms * 1e3:.2f} us  {flops / ms / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    for name in sys.argv[1:]:
        print("==", name)
        globals()[name]()
