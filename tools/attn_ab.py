#!/usr/bin/env python
"""A/B of the attention kernel's forms on the step's shapes (and config 5's): variant 0 = the default
choice, 8 = the single-buffered form, 16 = the pipelined form (two S / P buffers), +4 flips the packed / fp32 exp2 choice,
+32 / +64 = P through shared memory / through TMEM (head_dim 32 / 64).  Usage: attn_ab.py [variant ...]"""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops  # noqa: E402

SHAPES = [(16, 1000, 8, 32), (16, 252, 8, 48), (16, 64, 8, 80), (32, 3000, 8, 64), (32, 752, 8, 96), (32, 188, 8, 160),
          (2, 1, 8, 32), (3, 129, 8, 48), (2, 333, 8, 64)]
variants = [int(a) for a in sys.argv[1:]] or [8 + 32, 8 + 64, 16 + 32, 16 + 64]
for b, s, h, d in SHAPES:
    torch.manual_seed(0)
    qkv = (torch.randn(b, s, 3 * h * d, device="cuda") * 1.5).to(torch.bfloat16)
    out = torch.empty(b, s, h * d, dtype=torch.bfloat16, device="cuda")
    q, k, v = [t.view(b, s, h, d).transpose(1, 2).float() for t in qkv.chunk(3, -1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s, h * d)
    for variant in variants:
        out.fill_(float("nan"))
        for _ in range(5):
            ops.attention(qkv, out, b, s, h, d, variant=variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention(qkv, out, b, s, h, d, variant=variant)
        e1.record()
        torch.cuda.synchronize()
        err = ((out.float() - ref).norm() / ref.norm()).item()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"attention b{b} s{s} d{d} variant {variant}: {us:8.1f} us  {4 * b * h * s * s * d / us * 1e-6:7.1f} TFLOP/s  "
              f"rel-L2 vs fp32 SDPA {err:.2e}", flush=True)
