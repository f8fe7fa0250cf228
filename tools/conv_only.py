#!/usr/bin/env python
"""One level-0 3x3 convolution of the bench workload (128 -> 128 channels @ 250x16, UNet batch 16: M = 64000, N = 128,
K = 1152) in a loop, for ncu --set full captures.   python tools/conv_only.py [pair=0|1] [bn=128]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops, packing  # noqa: E402

pair = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 128
nb, h, w, ci, co = 16, 250, 16, 128, 128
x = torch.randn(nb, h, w, ci, device="cuda").to(torch.bfloat16)
wt = torch.randn(co, ci, 3, 3) * (9 * ci) ** -0.5
pw = packing.pack([packing.conv3x3_to_k(wt)], torch.zeros(co), bn, 9, ci, device="cuda")
out = torch.empty(nb * h * w, co, dtype=torch.bfloat16, device="cuda")
for _ in range(5):
    ops.conv_gemm(pw, x, nb, h, w, out, cta_pair=pair)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv_gemm(pw, x, nb, h, w, out, cta_pair=pair)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"conv L0 128->128 pair={int(pair)} bn={bn}: {us:.1f} us  {2.0 * nb * h * w * co * 9 * ci / us / 1e6:.0f} TFLOP/s")
