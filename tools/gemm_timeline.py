#!/usr/bin/env python
"""CTA-0 cycle timeline of one b200_conv_gemm launch: fixed points of the kernel, the MMA issuer's wait / issue split and
the time stamp of every k-block's TMA issues and MMA issue in the first tile (profiles/r01_gemm_mainloop_timeline.md).

Needs a profile build of the library (the product build carries no instrumentation):

    cd audioldm_with_lora_b200 && mkdir -p variants && for f in conv_gemm attention attention_bwd norm sampler train; do
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
           -DB200_GEMM_PROFILE=1 -c csrc/$f.cu -o variants/$f.o; done
    nvcc -shared -o variants/libb200ldm_prof.so variants/*.o -lcudart -gencode arch=compute_100a,code=sm_100a

    B200LDM_LIB=audioldm_with_lora_b200/variants/libb200ldm_prof.so B200_GEMM_DEBUG=12 python tools/gemm_timeline.py [ncases]

B200_GEMM_DEBUG bits: 2 = issue no MMAs (garbage results, timing only), 4 = timeline, 8 = per-k-block stamps.
MAXCTAS=n caps the grid (is a limit per SM or chip-wide?), PAIR=1 runs the 2-CTA mode, B200_GEMM_STAGES=n caps the ring.
"""
import ctypes
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from audioldm_with_lora_b200 import _lib, ops, packing  # noqa: E402

NAMES = {0: "entry", 1: "prologue done", 2: "pdl_wait done", 3: "producer: 1st TMA issued", 4: "producer: stages filled",
         5: "producer: done", 6: "mma: 1st full", 7: "mma: 2nd full", 8: "mma: kb 17", 9: "mma: last full", 10: "mma: tfull commit",
         16: "epi c0: math done", 17: "epi c0: 1st tmem_ld back", 18: "epi c0: smem written", 19: "epi c0: store issued",
         20: "epi c2: math done", 21: "epi c2: staging free", 22: "epi c2: smem written", 23: "epi c2: store issued",
         12: "epi: tfull seen", 13: "epi: tile done", 14: "epi: stores complete", 15: "exit"}
g = torch.Generator().manual_seed(0)
cases = [("conv L0 128->128", 16, 250, 16, 128, 128, 9), ("conv L1 256->256", 16, 125, 8, 256, 256, 9),
         ("lin L3 640->640", 1, 1024, 1, 640, 640, 1)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for label, nb, hh, ww, ci, co, taps in cases:
    x = torch.randn(nb * hh * ww, ci, generator=g).to("cuda", torch.bfloat16)
    wt = torch.randn(co, taps * ci, generator=g) * (taps * ci) ** -0.5
    out = torch.empty(nb * hh * ww, co, dtype=torch.bfloat16, device="cuda")
    bn = ops.choose_block_n(co, ops.num_m_tiles(nb, hh, ww))
    pw = packing.pack([wt], torch.zeros(co), bn, taps, ci, device="cuda")
    for _ in range(3):
        ops.conv_gemm(pw, x, nb, hh, ww, out, max_ctas=int(os.environ.get('MAXCTAS', '0')), cta_pair=bool(int(os.environ.get('PAIR', '0'))))
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 160)()
    _lib.check(_lib.load().b200_debug_timeline(ctypes.cast(buf, ctypes.c_void_p), 160), "timeline")
    t0 = buf[0]
    print(f"== {label} bn={bn}")
    for i in sorted(NAMES):
        if buf[i]:
            print(f"   {NAMES[i]:<28} +{buf[i] - t0:>7} cyc")
    if buf[31]:
        n = buf[31]
        print(f"   MMA thread over {n} k-blocks of the first tile: wait for operands {buf[25] / n:.0f}  issue+commit+advance {buf[28] / n:.0f} cyc/kb")
    if buf[32]:
        n = buf[31]
        base = buf[2]
        print("   kb: A issued / B issued / operands seen by MMA thread / MMAs issued+committed   (cycles after pdl_wait)")
        for k in range(min(n, 32)):
            print(f"   {k:>3}: {int(buf[64 + k]) - int(base):>7} {int(buf[96 + k]) - int(base):>7} {int(buf[32 + k]) - int(base):>7} {int(buf[128 + k]) - int(base):>7}")
