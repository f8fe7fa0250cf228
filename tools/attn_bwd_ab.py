#!/usr/bin/env python
"""Attention backward (b200_attention_bwd) at the fine-tuning step's shapes (config 4: UNet batch 32, 256x16 latents),
captured in a CUDA graph of 10 launches (the host call costs more than the small shapes).  Round 2 used it to A/B a
software-pipelined form of the backward kernel (two score buffers of half as many columns, as in the forward kernel): parity
green but SLOWER (b32 s1024 d32: 487 vs 429 us; b8 s3000 d64: 1417 vs 1362 us) -- the per-block statistics staging and
barriers double with half-size blocks -- so it was not kept."""
import sys
from pathlib import Path
import torch
import torch.nn.functional as F
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops  # noqa: E402

for b, s, h, d in [(32, 1024, 8, 32), (32, 256, 8, 48), (32, 64, 8, 80), (8, 3000, 8, 64), (2, 333, 8, 64), (2, 333, 8, 96)]:
    torch.manual_seed(0)
    C = h * d
    qkv = torch.randn(b, s, 3 * C, device="cuda").to(torch.bfloat16)
    dout = torch.randn(b, s, C, device="cuda").to(torch.bfloat16)
    out = torch.empty(b, s, C, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(b, h, s, dtype=torch.float32, device="cuda")
    delta = torch.empty(b, h, s, dtype=torch.float32, device="cuda")
    dqkv = torch.full((b, s, 3 * C), float("nan"), dtype=torch.bfloat16, device="cuda")
    ops.attention_lse(qkv, out, lse, b, s, h, d)
    ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, b, s, h, d)
    torch.cuda.synchronize()
    err = float("nan")
    if b * s * s * h <= 32 * 1024 * 1024 * 8:
        qr = qkv[:2].float().requires_grad_(True)
        q, k, v = [t.view(2, s, h, d).transpose(1, 2) for t in qr.chunk(3, -1)]
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(2, s, C)
        o.backward(dout[:2].float())
        err = ((dqkv[:2].float() - qr.grad).norm() / qr.grad.norm()).item()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, b, s, h, d)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                ops.attention_bwd(qkv, out, dout, lse, delta, dqkv, b, s, h, d)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    print(f"attention_bwd b{b} s{s} d{d}: {us:8.1f} us  {14 * b * h * s * s * d / us * 1e-6:7.1f} TFLOP/s  rel-L2 vs fp32 autograd {err:.2e}",
          flush=True)
