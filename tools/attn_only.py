#!/usr/bin/env python
"""Just the level-1 attention launch of the bench workload (b16 s1000 d32), for ncu source-level captures."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops  # noqa: E402
b, s, h, d = 16, 1000, 8, 32
qkv = (torch.randn(b, s, 3 * h * d, device="cuda") * 1.0).to(torch.bfloat16)
out = torch.empty(b, s, h * d, dtype=torch.bfloat16, device="cuda")
import torch.nn.functional as F
q, k, v = [t.view(b, s, h, d).transpose(1, 2).float() for t in qkv.chunk(3, -1)]
ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s, h * d)
for variant in [int(a) for a in sys.argv[1:]] or [0]:
    for _ in range(5):
        ops.attention(qkv, out, b, s, h, d, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.attention(qkv, out, b, s, h, d, variant=variant)
    e1.record(); torch.cuda.synchronize()
    err = ((out.float() - ref).norm() / ref.norm()).item()
    print(f"attention b{b} s{s} d{d} variant {variant}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us   rel-L2 vs fp32 SDPA {err:.2e}")
