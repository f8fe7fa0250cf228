#!/usr/bin/env python
"""CTA-(0,0) cycle timeline of the attention kernel at the step's three shapes (needs the -DB200_ATTN_PROFILE=1 build:
tools/build_variant.sh attnprof -DB200_ATTN_PROFILE=1;  B200LDM_LIB=.../variants/libb200ldm_attnprof.so python tools/attn_timeline.py)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import ops  # noqa: E402
for b, s, h, d in [(16, 64, 8, 80), (16, 252, 8, 48), (16, 1000, 8, 32)]:
    qkv = torch.randn(b, s, 3 * h * d, device="cuda").to(torch.bfloat16)
    out = torch.empty(b, s, h * d, dtype=torch.bfloat16, device="cuda")
    print(f"--- b{b} s{s} d{d}", flush=True)
    for _ in range(3):
        ops.attention(qkv, out, b, s, h, d)
        torch.cuda.synchronize()
